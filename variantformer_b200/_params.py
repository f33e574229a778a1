"""Parameter containers: nn.Module trees whose `state_dict()` keys and shapes equal the reference's, so reference
checkpoints load with `load_state_dict` unchanged.  They hold weights only — every forward in this
package goes through the CUDA engine, never through these modules."""
import torch
import torch.nn as nn


class Affine(nn.Module):
    """weight [out, in] + bias [out] (nn.Linear layout) or weight/bias [d] (LayerNorm layout); uninitialised."""

    def __init__(self, *shape, bias_shape=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(*(bias_shape or shape[:1])), requires_grad=False)


class Table(nn.Module):
    """nn.Embedding-shaped parameter: weight [rows, dim]."""

    def __init__(self, rows, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(rows, dim), requires_grad=False)


class MHAParams(nn.Module):
    """flash_attn.modules.mha.MHA parameter layout: self (Wqkv "(three h d)") or cross (Wq, Wkv "(two h d)")."""

    def __init__(self, d, cross=False):
        super().__init__()
        if cross:
            self.Wq = Affine(d, d); self.Wkv = Affine(2 * d, d)
        else:
            self.Wqkv = Affine(3 * d, d)
        self.out_proj = Affine(d, d)


class AttnBlock(nn.Module):
    """`FlashAttLayer` container: `.MHA`."""

    def __init__(self, d, cross=False):
        super().__init__()
        self.MHA = MHAParams(d, cross)


def load_into(module: nn.Module, state_dict, strict=True):
    return module.load_state_dict(state_dict, strict=strict)
