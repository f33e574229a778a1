"""seq2gene layer containers (parameter layout of the reference's seq2gene/modules/layers.py:47-86, :502-524,
:1012-1111).  Execution: Engine._layer / Engine.run."""
import torch
import torch.nn as nn

from ..._params import Affine, AttnBlock, Table
from ...utils.alibi import alibi_slopes

FFN_HIDDEN = 2048


class ContextFlashAttentionEncoderLayer(nn.Module):
    """LN1 -> self-MHA(+ALiBi) -> +src -> LN2 -> cross-MHA(context) -> + -> LN3 -> GeGLU FFN -> +src(layer input)."""

    def __init__(self, d_model, nhead, hidden_dim=FFN_HIDDEN, dropout=0.1, batch_first=True, use_alibi=False,
                 make_data_kv=False, mlp_dout=0.0, cross_alibi=False, flash_attn_3=False):
        super().__init__()
        if make_data_kv or cross_alibi or flash_attn_3:
            raise NotImplementedError("make_data_kv / cross_alibi / flash_attn_3 variants are not on the hot path")
        self.mixer = AttnBlock(d_model)
        self.crossMHA = AttnBlock(d_model, cross=True)
        self.norm1 = Affine(d_model); self.norm2 = Affine(d_model); self.norm3 = Affine(d_model)
        self.linear_geglu_1 = Affine(hidden_dim, d_model)
        self.linear_geglu_2 = Affine(d_model, hidden_dim // 2)
        self.use_alibi, self.num_heads = use_alibi, nhead
        if use_alibi:
            self.register_buffer("m", alibi_slopes(nhead))          # persistent, like the reference (unused in forward)


class MultiRegistry(nn.Module):
    """One learned registry token per tissue, prepended to the gene stream (layers.py:502-524)."""

    def __init__(self, num_tissues, emb_dim):
        super().__init__()
        self.num_registry_tokens = num_tissues
        self.registry_tokens = Table(num_tissues, emb_dim)

    def get_registry_tokens(self):
        return self.registry_tokens.weight


class TissueExpressionHeads(nn.Module):
    """Shared 'bigger' MLP head + Softplus: Sequential indices 0 Linear, 1 LayerNorm, 4 Linear, 6 Linear(emb,1)."""

    def __init__(self, emb_dim, num_tissues, use_bigger_head=False, multi_head=True, mlp_dout=0.1, loss_fn="poisson",
                 head_type="mlp"):
        super().__init__()
        if not (use_bigger_head and not multi_head and head_type == "mlp" and loss_fn == "poisson"):
            raise NotImplementedError("only the vf_model.yaml head (use_bigger_head, shared, mlp, poisson) is implemented")
        self.multi_head = multi_head
        self.tissue_expressions = nn.ModuleDict({"0": Affine(emb_dim, emb_dim), "1": Affine(emb_dim),
                                                 "4": Affine(emb_dim, emb_dim), "6": Affine(1, emb_dim)})
