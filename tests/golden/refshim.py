"""Harness that imports the *reference's own Python* in the build container.

Used ONLY by tests/golden/make_*_golden.py (fixture generation) and by the
optional `-m "not gpu"` cross-checks that are skipped when /root/reference is
absent (it never exists on the GPU box).  Follows SURVEY.md Appendix C:
  * stub `lightning.pytorch.LightningModule` and `pybedtools` (not installed),
  * replace flash_attn's inner attention `forward`s (they assert is_cuda) with
    a per-sequence fp32 math attention,
  * undo the process-global `set_float32_matmul_precision("medium")` that
    importing seq2reg/model.py performs (seq2reg/model.py:12).
Nothing here is product code and nothing is copied from the reference.
"""
import inspect
import math
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("VF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "seq2gene"))


class _HP(dict):
    __getattr__ = dict.get


class _LightningModule(nn.Module):
    def save_hyperparameters(self, ignore=None):
        loc = dict(inspect.currentframe().f_back.f_locals)
        loc.pop("self", None); loc.pop("__class__", None)
        loc.update(loc.pop("kwargs", {}))
        for k in (ignore or []):
            loc.pop(k, None)
        self.hparams = _HP(loc)

    def log(self, *a, **k):
        pass


def _varlen_attention(q, k, v, cu_q, cu_k, slopes, scale):
    out = torch.empty_like(q)
    scale = scale or 1.0 / math.sqrt(q.shape[-1])
    for b in range(len(cu_q) - 1):
        qs, qe, ks, ke = cu_q[b], cu_q[b + 1], cu_k[b], cu_k[b + 1]
        s = torch.einsum("thd,shd->hts", q[qs:qe].float(), k[ks:ke].float()) * scale
        if slopes is not None:
            sq, sk = qe - qs, ke - ks
            i = torch.arange(sq)[:, None]; j = torch.arange(sk)[None, :]
            s = s - slopes.float()[:, None, None] * (i + sk - sq - j).abs()[None]
        out[qs:qe] = torch.einsum("hts,shd->thd", s.softmax(-1), v[ks:ke].float()).to(q.dtype)
    return out


_installed = False


def install():
    """Register stubs + patches and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    lp = types.ModuleType("lightning.pytorch"); lp.LightningModule = _LightningModule
    l = types.ModuleType("lightning"); l.pytorch = lp
    sys.modules.setdefault("lightning", l)
    sys.modules.setdefault("lightning.pytorch", lp)
    sys.modules.setdefault("pybedtools", types.ModuleType("pybedtools"))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import flash_attn.modules.mha as mha

    def _self(self, qkv, causal=None, cu_seqlens=None, max_seqlen=None):
        if cu_seqlens is None:
            B, S = qkv.shape[:2]
            cu = list(range(0, (B + 1) * S, S)); x = qkv.reshape(B * S, *qkv.shape[2:])
            return _varlen_attention(x[:, 0], x[:, 1], x[:, 2], cu, cu, self.alibi_slopes,
                                     self.softmax_scale).reshape(B, S, *qkv.shape[3:])
        cu = cu_seqlens.tolist()
        return _varlen_attention(qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu, self.alibi_slopes, self.softmax_scale)

    def _cross(self, q, kv, causal=None, cu_seqlens=None, max_seqlen=None, cu_seqlens_k=None, max_seqlen_k=None):
        return _varlen_attention(q, kv[:, 0], kv[:, 1], cu_seqlens.tolist(), cu_seqlens_k.tolist(),
                                 self.alibi_slopes, self.softmax_scale)

    mha.FlashSelfAttention.forward = _self
    mha.FlashCrossAttention.forward = _cross
    import seq2reg.model  # noqa: F401  (flips matmul precision to "medium")
    import seq2gene.model_combined_modulator  # noqa: F401
    torch.set_float32_matmul_precision("highest")
    _installed = True


def build_reference_model(cfg: dict, seq2reg_hp: dict, seed: int = 0):
    """Random-init reference Seq2GenePredictorCombinedModulator (+2 Seq2RegPredictor) in fp32 eval mode."""
    install()
    from seq2reg.model import Seq2RegPredictor
    from seq2gene.model_combined_modulator import Seq2GenePredictorCombinedModulator
    torch.manual_seed(seed)
    cre_tok = Seq2RegPredictor(**seq2reg_hp)
    gene_tok = Seq2RegPredictor(**seq2reg_hp)
    model = Seq2GenePredictorCombinedModulator(cre_tokenizer=cre_tok, gene_tokenizer=gene_tok, **cfg)
    model.eval()
    model.vep = False
    model.trainer = types.SimpleNamespace(precision="bf16-mixed")  # -> internal precision=None, no fp16 casts
    return model
