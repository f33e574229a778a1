"""Single-module forwards through the C ABI: what `forward()` of the mirror classes runs when a caller uses the
reference's layer-level API instead of the batched engine schedule.

Reference signatures served here (file:line under the reference checkout):
  seq2reg/modules.py:149                        FlashTransformerLayer.forward(src, src_key_padding_mask, precision)
  seq2gene/modules/layers.py:88-98              ContextFlashAttentionEncoderLayer.forward(src, context, ...)
  seq2gene/modules/layers.py:1113               TissueExpressionHeads.forward(g_exp, tissue_vector)
  seq2gene/model_combined_modulator.py:137-148  CombinedModulator.forward(cre_x, gene_x, context, masks, ...)

Conventions shared with the reference: padding masks are True = padding; `unpad_info` dicts carry `cu_seqlens`
(int32 prefix sums of the valid lengths) exactly as flash_attn.bert_padding.unpad_input returns them; a layer called
with `unpad_info` takes and returns unpadded [rows, d] tensors.  Rows of a padded output that correspond to padding
are zero (the reference leaves values there that no consumer reads).

Every function here computes with the same kernels and the same numerics policy as the engine (bf16 operands, fp32
accumulation, fp32 row-centred residual stream, LayerNorm folded into the consuming GEMM).
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import ops
from ._lib import EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GEGLU_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_RESID_F32
from .engine import (AttnPlan, Engine, Workspace, _Linear, _Norm, context_layer_weights, cross_layer_weights,
                     seq2reg_layer_weights)


def _require_cuda(t, what):
    if t.device.type != "cuda":
        raise RuntimeError(f"{what} runs on a B200 only: move the module and its inputs to CUDA first (no CPU path)")


class _Cache:
    """Device-side folded weights of one module, rebuilt when its parameters move or are reloaded."""

    def __init__(self):
        self.key, self.val = None, None

    def get(self, module, build):
        p = next(module.parameters())
        key = (p.device, p.data_ptr(), p._version)
        if self.key != key:
            _require_cuda(p, type(module).__name__)
            self.val, self.key = build({k: v for k, v in module.state_dict().items()}, p.device), key
        return self.val


def _lens_from(mask, unpad_info, batch, seqlen):
    """-> (valid lengths int64 [B], keep mask bool [B, S] or None when the input is already unpadded)."""
    if unpad_info is not None:
        cu = unpad_info["cu_seqlens"]
        cu = cu.detach().cpu().numpy() if torch.is_tensor(cu) else np.asarray(cu)
        return np.diff(cu).astype(np.int64), None
    if mask is not None:
        keep = ~mask.bool()
        return keep.sum(1).cpu().numpy().astype(np.int64), keep
    return np.full(batch, seqlen, np.int64), None


def _unpad(x, keep):
    """Valid rows as a FRESH fp32 matrix (the kernels centre it in place: never alias the caller's tensor)."""
    if keep is not None:
        return x[keep].float().contiguous()
    return x.reshape(-1, x.shape[-1]).to(torch.float32, copy=True).contiguous()


def _repad(rows, keep, like):
    if like.dim() == 2:
        return rows
    if keep is None:
        return rows.view(like.shape[0], like.shape[1], -1)
    out = torch.zeros(like.shape[0], like.shape[1], rows.shape[1], dtype=rows.dtype, device=rows.device)
    out[keep] = rows
    return out


def _runner(ws, D):
    eng = Engine.__new__(Engine)
    eng.ws, eng.w = ws, SimpleNamespace(D=D)
    return eng


# ------------------------------------------------------------------------------------------------------------------
def seq2reg_layer_forward(L, ws, nhead, slopes, src, src_key_padding_mask=None, unpad_info=None):
    """FlashTransformerLayer.forward / seq2gene FlashAttentionEncoderLayer.forward: x = MHA(LN1(src)) + src;
    out = FFN(LN2(x)) + src (padded [B, S, d] with a mask, or unpadded [rows, d] with unpad_info)."""
    _require_cuda(src, "FlashTransformerLayer")
    B, S = (src.shape[0], src.shape[1]) if src.dim() == 3 else (0, 0)
    d = src.shape[-1]
    lens, keep = _lens_from(src_key_padding_mask, unpad_info, B, S)
    x = _unpad(src, keep)
    n = x.shape[0]
    dev = x.device
    xb = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    piv, st0 = ops.center_rows(x, out_bf16=xb)
    plan = AttnPlan(lens, dev, d // nhead)
    qkv = ops.gemm(xb, L["qkv"].w, EPI_BIAS_BF16, bias=L["qkv"].b, ln=L["qkv"].ln(st0))
    a = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    plan.run(qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:], nhead, d // nhead, slopes, a)
    s1 = torch.empty((n, ops.stats_parts(d), 2), dtype=torch.float32, device=dev)
    ops.gemm(a, L["out"].w, EPI_BIAS_RESID_F32, bias=L["out"].b, resid=x, out2=xb, stats_out=s1, mirror_only=True)
    f = ops.gemm(xb, L["g1"].w, EPI_BIAS_GEGLU_BF16, bias=L["g1"].b, ln=L["g1"].ln(s1))
    ops.gemm(f, L["g2"].w, EPI_BIAS_RESID_F32, bias=L["g2"].b, resid=x, out=x)
    out, _ = ops.uncenter_rows(x, piv)
    return _repad(out, keep, src).to(src.dtype)


def context_layer_forward(L, ws, D, nhead, slopes, src, context, src_key_padding_mask=None, context_padding_mask=None,
                          unpad_info=None, context_unpad_info=None, gene_unpad_info=None, cross_slopes=None):
    """ContextFlashAttentionEncoderLayer.forward (layers.py:88-165) on padded [B, S, D] or unpadded [rows, D] input.
    cross_slopes: ALiBi slopes of the cross-attention (cross_alibi=True: bias -slope |i + Sk - Sq - j|) or None."""
    _require_cuda(src, "ContextFlashAttentionEncoderLayer")
    src_info = gene_unpad_info if gene_unpad_info is not None else unpad_info
    if context_padding_mask is None and src_key_padding_mask is not None and context_unpad_info is None:
        context_padding_mask = src_key_padding_mask                      # layers.py:104-105
    B, S = (src.shape[0], src.shape[1]) if src.dim() == 3 else (0, 0)
    q_lens, keep = _lens_from(src_key_padding_mask, src_info, B, S)
    Bc, Sc = (context.shape[0], context.shape[1]) if context.dim() == 3 else (0, 0)
    k_lens, ckeep = _lens_from(context_padding_mask, context_unpad_info, Bc, Sc)
    assert len(q_lens) == len(k_lens), "src and context must hold the same number of sequences"
    x = _unpad(src, keep)
    ctx = ops.cast_bf16(_unpad(context, ckeep))
    n, dev, hd = x.shape[0], x.device, D // nhead
    xb = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
    piv, st0 = ops.center_rows(x, out_bf16=xb)
    kv = ops.gemm(ctx, L["kv"].w, EPI_BIAS_BF16, bias=L["kv"].b)
    plan_self, plan_cross = AttnPlan(q_lens, dev, hd), AttnPlan(q_lens, dev, hd, k_lens=k_lens)

    def self_attn(qkv, out):
        plan_self.run(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], nhead, hd, slopes, out)

    def cross_attn(q, out):
        plan_cross.run(q, kv[:, :D], kv[:, D:], nhead, hd, cross_slopes, out)
    _runner(ws, D)._layer(L, x, xb, st0, n, self_attn, cross_attn, "m")
    out, _ = ops.uncenter_rows(x, piv)
    return _repad(out, keep, src).to(src.dtype)


def cross_layer_forward(L, ws, D, nhead, cross_slopes, src, context, src_key_padding_mask=None, context_padding_mask=None,
                        gene_unpad_info=None, context_unpad_info=None):
    """ContextFlashCrossAttentionEncoderLayer.forward (layers.py:268-325): x = MHAcross(LN1(src), context) + src;
    out = FFN(LN2(x)) + src.  cross_slopes: ALiBi slopes of the cross-attention (cross_alibi=True) or None."""
    _require_cuda(src, "ContextFlashCrossAttentionEncoderLayer")
    B, S = (src.shape[0], src.shape[1]) if src.dim() == 3 else (0, 0)
    q_lens, keep = _lens_from(src_key_padding_mask, gene_unpad_info, B, S)
    Bc, Sc = (context.shape[0], context.shape[1]) if context.dim() == 3 else (0, 0)
    k_lens, ckeep = _lens_from(context_padding_mask, context_unpad_info, Bc, Sc)
    assert len(q_lens) == len(k_lens), "src and context must hold the same number of sequences"
    x = _unpad(src, keep)
    ctx = ops.cast_bf16(_unpad(context, ckeep))
    n, dev, hd = x.shape[0], x.device, D // nhead
    xb = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
    piv, st0 = ops.center_rows(x, out_bf16=xb)
    kv = ops.gemm(ctx, L["kv"].w, EPI_BIAS_BF16, bias=L["kv"].b)
    q = ops.gemm(xb, L["q"].w, EPI_BIAS_BF16, bias=L["q"].b, ln=L["q"].ln(st0))
    a = torch.empty((n, D), dtype=torch.bfloat16, device=dev)
    AttnPlan(q_lens, dev, hd, k_lens=k_lens).run(q, kv[:, :D], kv[:, D:], nhead, hd, cross_slopes, a)
    s1 = torch.empty((n, ops.stats_parts(D), 2), dtype=torch.float32, device=dev)
    ops.gemm(a, L["out2"].w, EPI_BIAS_RESID_F32, bias=L["out2"].b, resid=x, out2=xb, stats_out=s1, mirror_only=True)
    f = ops.gemm(xb, L["g1"].w, EPI_BIAS_GEGLU_BF16, bias=L["g1"].b, ln=L["g1"].ln(s1))
    ops.gemm(f, L["g2"].w, EPI_BIAS_RESID_F32, bias=L["g2"].b, resid=x, out=x)
    out, _ = ops.uncenter_rows(x, piv)
    return _repad(out, keep, src).to(src.dtype)


def head_weights(sd, device, prefix="tissue_expressions."):
    return dict(h0=_Linear(sd, prefix + "0", device), hn=_Norm(sd, prefix + "1", device), h4=_Linear(sd, prefix + "4", device),
                h6_w=sd[prefix + "6.weight"].to(device=device, dtype=torch.float32).reshape(-1).contiguous(),
                h6_b=sd[prefix + "6.bias"].to(device=device, dtype=torch.float32).contiguous())


def head_forward(W, g_exp):
    """TissueExpressionHeads.forward for the shared 'bigger' head (layers.py:1078-1087): every row goes through the same
    MLP (tissue enters through the registry token upstream), Softplus output.  -> [batch, 1]."""
    _require_cuda(g_exp, "TissueExpressionHeads")
    e = ops.cast_bf16(g_exp.float().contiguous())
    h1 = ops.gemm(e, W["h0"].w, EPI_BIAS_F32, bias=W["h0"].b)
    h1n = ops.layernorm(h1, W["hn"].g, W["hn"].b, gelu=True)
    h2 = ops.gemm(h1n, W["h4"].w, EPI_BIAS_GELU_BF16, bias=W["h4"].b)
    return ops.head_out(h2, W["h6_w"], W["h6_b"], softplus=True).unsqueeze(1)


def combined_modulator_forward(mod, cre_x, gene_x, context=None, cre_padding_mask=None, gene_padding_mask=None,
                               context_padding_mask=None, cre_token_position=None, gene_token_position=None):
    """CombinedModulator.forward (model_combined_modulator.py:137-328): gene_0(g, cre); for i: cre_i(cre[, label
    context]); gene_{i+1}(g, cre) [+ gene_res when use_res] — on whatever batch the caller assembled (one CRE stream per
    batch row, no tissue de-duplication: that is the engine's job).  Serves every layer variant the reference can
    build: use_context (CRE layers with label cross-attention) or not, only_cross_attention gene layers, use_res,
    cross_alibi.  -> (gene_out [B, Sg, D], gene_token_embedding, cre_token_embedding)."""
    B = cre_x.shape[0]
    ctx_emb = None
    if mod.use_context and context is not None:
        ctx_emb = mod.second_level_context_embedding.weight[context.long()]       # [B, Sc, D]
    cmask = context_padding_mask if context_padding_mask is not None else cre_padding_mask
    g, c = gene_x, cre_x
    gene_res = gene_x.clone() if mod.use_res else None

    def gene_layer(i, g, c):
        g = mod.gene_layers[i](g, c, src_key_padding_mask=gene_padding_mask, context_padding_mask=cre_padding_mask)
        return g + gene_res if gene_res is not None else g
    g = gene_layer(0, g, c)
    for i in range(mod.num_layers - 1):
        if mod.use_context:
            c = mod.cre_layers[i](c, ctx_emb, src_key_padding_mask=cre_padding_mask, context_padding_mask=cmask)
        else:
            c = mod.cre_layers[i](c, src_key_padding_mask=cre_padding_mask)
        g = gene_layer(i + 1, g, c)
    ar = torch.arange(B, device=g.device)
    zeros = torch.zeros(B, g.shape[2], device=g.device, dtype=g.dtype)
    gtok = g[ar, gene_token_position.long().reshape(-1)] if gene_token_position is not None else zeros
    ctok = c[ar, cre_token_position.long().reshape(-1)] if cre_token_position is not None else zeros.clone()
    return g, gtok, ctok


__all__ = ["seq2reg_layer_forward", "context_layer_forward", "cross_layer_forward", "head_forward",
           "combined_modulator_forward", "seq2reg_layer_weights", "context_layer_weights", "cross_layer_weights",
           "head_weights", "Workspace", "_Cache"]
