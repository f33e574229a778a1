"""CPU ORACLE (test infrastructure, NOT product code) — stages 2-4 in plain torch fp32.

A functional restatement of the reference's PyTorch forward that works directly
on a reference `state_dict` (same key names), with no Lightning, no flash_attn
and no autocast.  Floating-point path => the oracle is a torch fp32 reference
(the tier rules keep torch fp32 for floating-point kernels).  It is pinned by
tests/golden/model_golden.npz, produced by tests/golden/make_model_golden.py
from the reference's own classes (imported in the build container).

Reference sites restated (paths relative to the reference checkout):
  seq2reg/model.py:193-279            Seq2RegPredictor.forward(only_embed=True)
  seq2reg/modules.py:149-191          FlashTransformerLayer.forward
  seq2gene/modules/layers.py:88-165   ContextFlashAttentionEncoderLayer.forward
  seq2gene/model_combined_modulator.py:137-328, 540-720, 722-829, 857-907
  seq2gene/modules/layers.py:508-521  MultiRegistry
  seq2gene/modules/layers.py:1078-1087,1113-1144  TissueExpressionHeads (shared bigger head)
  flash_attn.modules.mha.MHA          packed Wqkv "(three h d)", Wkv "(two h d)", ALiBi -slope*|i+Sk-Sq-j|

`schedule="reference"` repeats the CRE stream once per tissue exactly like
model_combined_modulator.py:622-649; `schedule="dedup"` computes the
tissue-independent CRE stream once per gene (the product's schedule).  A test
asserts both give identical results.

`emulate_bf16=True` rounds every GEMM / attention operand to bf16 (fp32
accumulate, fp32 residual stream): a numerical model of the CUDA path used to
choose tolerances, never a substitute for it.
"""
import math

import torch
import torch.nn.functional as F


def alibi_slopes(n: int) -> torch.Tensor:
    """seq2gene/modules/layers.py:15-37 / seq2reg/modules.py:13-33."""
    def pow2(n):
        start = 2 ** (-(2 ** -(math.log2(n) - 3)))
        return [start * start ** i for i in range(n)]
    if math.log2(n).is_integer():
        return torch.tensor(pow2(n), dtype=torch.float32)
    c = 2 ** math.floor(math.log2(n))
    return torch.tensor(pow2(c) + alibi_slopes(2 * c)[0::2][: n - c].tolist(), dtype=torch.float32)


def sinusoidal_pe(d_model: int, length: int) -> torch.Tensor:
    """seq2reg/model.py:15-37."""
    pe = torch.zeros(length, d_model)
    position = torch.arange(0, length).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float) * -(math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div)
    pe[:, 1::2] = torch.cos(position * div)
    return pe


class _Num:
    """Numerics policy: plain fp32, or bf16-rounded GEMM/attention operands."""

    def __init__(self, emulate_bf16=False, stream_bf16=False):
        self.bf16 = emulate_bf16
        self.stream_bf16 = stream_bf16        # design probe: residual stream rounded to bf16 after every add

    def s(self, x):
        return x.to(torch.bfloat16).float() if self.stream_bf16 else x

    def r(self, x):
        return x.to(torch.bfloat16).float() if self.bf16 else x

    def linear(self, x, w, b):
        return F.linear(self.r(x), self.r(w), b)


def _attention(num, q, k, v, cu_q, cu_k, slopes):
    """Varlen non-causal attention; q [Tq,H,D], k/v [Tk,H,D]; softmax scale 1/sqrt(D)."""
    out = torch.empty_like(q)
    scale = 1.0 / math.sqrt(q.shape[-1])
    q, k, v = num.r(q), num.r(k), num.r(v)
    for b in range(len(cu_q) - 1):
        qs, qe, ks, ke = cu_q[b], cu_q[b + 1], cu_k[b], cu_k[b + 1]
        s = torch.einsum("thd,shd->hts", q[qs:qe], k[ks:ke]) * scale
        if slopes is not None:
            sq, sk = qe - qs, ke - ks
            i = torch.arange(sq)[:, None]; j = torch.arange(sk)[None, :]
            s = s - slopes[:, None, None] * (i + sk - sq - j).abs()[None]
        p = s.softmax(-1)
        out[qs:qe] = torch.einsum("hts,shd->thd", num.r(p) if num.bf16 else p, v[ks:ke])
    return out


def _mha_self(num, sd, pre, x, cu, H, slopes):
    n, d = x.shape
    qkv = num.linear(x, sd[pre + "Wqkv.weight"], sd[pre + "Wqkv.bias"]).view(n, 3, H, d // H)
    o = _attention(num, qkv[:, 0], qkv[:, 1], qkv[:, 2], cu, cu, slopes).reshape(n, d)
    return num.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def _mha_cross(num, sd, pre, x, ctx, cu_q, cu_k, H):
    n, d = x.shape
    q = num.linear(x, sd[pre + "Wq.weight"], sd[pre + "Wq.bias"]).view(n, H, d // H)
    kv = num.linear(ctx, sd[pre + "Wkv.weight"], sd[pre + "Wkv.bias"]).view(ctx.shape[0], 2, H, d // H)
    o = _attention(num, q, kv[:, 0], kv[:, 1], cu_q, cu_k, None).reshape(n, d)
    return num.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])


def _ln(sd, pre, x):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + "weight"], sd[pre + "bias"], 1e-5)


def _geglu_ffn(num, sd, pre, x):
    u, gate = num.linear(x, sd[pre + "linear_geglu_1.weight"], sd[pre + "linear_geglu_1.bias"]).chunk(2, dim=-1)
    return num.linear(u * F.gelu(gate), sd[pre + "linear_geglu_2.weight"], sd[pre + "linear_geglu_2.bias"])


def seq2reg_embed(sd, pre, hp, tokens, pad_mask, num=None):
    """Seq2RegPredictor.forward(only_embed=True) on valid tokens only.

    tokens int64 [n, L], pad_mask bool [n, L] (True = padding) -> fp32 [n, d].
    Padded positions never influence valid ones (attention is varlen over valid
    tokens, everything else is per-token) and are excluded from the mean
    (seq2reg/model.py:263-267), so they are simply not computed.
    """
    num = num or _Num()
    H, L = hp["num_heads"], hp["num_layers"]
    n, T = tokens.shape
    keep = ~pad_mask
    lens = keep.sum(1)
    cu = [0] + torch.cumsum(lens, 0).tolist()
    pos = torch.arange(T).expand(n, T)[keep]
    x = sd[pre + "token_embedding.weight"][tokens[keep]]
    if hp.get("positional_encoding", "sinusoidal") == "sinusoidal":
        x = x + sinusoidal_pe(x.shape[1], T)[pos]
        slopes = None
    else:
        slopes = alibi_slopes(H)
    for l in range(L):
        p = f"{pre}transformer_encoder.{l}."
        src = x
        a = _mha_self(num, sd, p + "MHA.", _ln(sd, p + "norm1.", src), cu, H, slopes)
        x = num.s(a + src)
        x = num.s(_geglu_ffn(num, sd, p, _ln(sd, p + "norm2.", x)) + src)   # res_long = layer input
    seg = torch.repeat_interleave(torch.arange(n), lens)
    out = torch.zeros(n, x.shape[1]).index_add_(0, seg, x)
    return out / lens[:, None]                                              # 0/0 = NaN for an all-pad window, as upstream


def _context_layer(num, sd, p, src, cu_src, context, cu_ctx, H, slopes):
    """ContextFlashAttentionEncoderLayer.forward on unpadded streams (layers.py:88-165)."""
    a = _mha_self(num, sd, p + "mixer.MHA.", _ln(sd, p + "norm1.", src), cu_src, H, slopes)
    x = num.s(a + src)
    c = _mha_cross(num, sd, p + "crossMHA.MHA.", _ln(sd, p + "norm2.", x), context, cu_src, cu_ctx, H)
    x = num.s(c + x)
    return num.s(_geglu_ffn(num, sd, p, _ln(sd, p + "norm3.", x)) + src)


def head(num, sd, e):
    """layers.py:1078-1087 shared 'bigger' head + Softplus."""
    p = "tissue_heads.tissue_expressions."
    h = num.linear(e, sd[p + "0.weight"], sd[p + "0.bias"])
    h = F.gelu(F.layer_norm(h, (h.shape[-1],), sd[p + "1.weight"], sd[p + "1.bias"], 1e-5))
    h = F.gelu(num.linear(h, sd[p + "4.weight"], sd[p + "4.bias"]))
    return F.softplus(num.linear(h, sd[p + "6.weight"], sd[p + "6.bias"]))


@torch.no_grad()
def predict_step(sd, cfg, seq2reg_hp, batch, schedule="reference", emulate_bf16=False,
                 return_streams=False, stream_bf16=False):
    """Restates Seq2GenePredictorCombinedModulator.predict_step for the vf_model.yaml variant
    (use_context, multi_registry, not only_cross_attention, shared bigger head).

    batch: dict with the reference's collate keys (vcfdataset.py:53-63).
    -> {"pred_gene_exp": [np (T_i,1)], "embeddings": [np (T_i,emb)]}
    """
    num = _Num(emulate_bf16, stream_bf16)
    H, NL, D = cfg["num_heads"], cfg["num_layers"], cfg["emb_dim"]
    slopes = alibi_slopes(H) if cfg.get("use_alibi", True) else None
    B = len(batch["cre_sequences"])
    preds, embs, streams = [], [], []
    for g in range(B):
        cre_tok = batch["cre_sequences"][g][:, 0, :]; cre_mask = batch["cre_attention_masks"][g][:, 0, :]
        gene_tok = batch["gene_embeddings"][g][:, 0, :]; gene_mask = batch["gene_attention_masks"][g][:, 0, :]
        tissues = batch["tissue_context"][g].long()
        labels = batch["ref_cre_labels"][g].long()
        T, Cn, Gn = len(tissues), cre_tok.shape[0], gene_tok.shape[0]
        cre = seq2reg_embed(sd, "cre_tokenizer.", seq2reg_hp, cre_tok, cre_mask, num)
        gene = seq2reg_embed(sd, "gene_tokenizer.", seq2reg_hp, gene_tok, gene_mask, num)
        if "cre_map.weight" in sd:
            cre = num.linear(cre, sd["cre_map.weight"], sd["cre_map.bias"])
        gene = num.linear(gene, sd["gene_map.weight"], sd["gene_map.bias"])
        ctx = sd["combined_modulator.second_level_context_embedding.weight"][labels]
        reg = sd["start_tkn.registry_tokens.weight"][tissues]                      # [T, D]
        gx = torch.cat([reg[:, None, :], gene[None].expand(T, Gn, D)], 1).reshape(T * (Gn + 1), D)
        cu_gene = [i * (Gn + 1) for i in range(T + 1)]
        R = 1 if schedule == "dedup" else T                                       # copies of the CRE stream
        cx = cre.repeat(R, 1); cctx = ctx.repeat(R, 1)
        cu_cre = [i * Cn for i in range(R + 1)]
        cu_cre_for_gene = cu_cre if R == T else None

        def gene_layer(i, gx, cx):
            p = f"combined_modulator.gene_layers.{i}."
            if cu_cre_for_gene is not None:
                return _context_layer(num, sd, p, gx, cu_gene, cx, cu_cre_for_gene, H, slopes)
            # dedup: every tissue copy attends to the single shared CRE stream
            a = _mha_self(num, sd, p + "mixer.MHA.", _ln(sd, p + "norm1.", gx), cu_gene, H, slopes)
            x = num.s(a + gx)
            c = _mha_cross(num, sd, p + "crossMHA.MHA.", _ln(sd, p + "norm2.", x), cx,
                           [0, gx.shape[0]], [0, Cn], H)
            x = num.s(c + x)
            return num.s(_geglu_ffn(num, sd, p, _ln(sd, p + "norm3.", x)) + gx)

        gx = gene_layer(0, gx, cx)
        for i in range(NL - 1):
            cx = _context_layer(num, sd, f"combined_modulator.cre_layers.{i}.", cx, cu_cre, cctx, cu_cre, H, slopes)
            gx = gene_layer(i + 1, gx, cx)
        emb = gx.view(T, Gn + 1, D)[:, 0, :]
        pred = head(num, sd, emb)
        preds.append(pred.numpy().copy()); embs.append(emb.numpy().copy())
        if return_streams:
            streams.append({"cre_in": cre.numpy().copy(), "gene_in": gene.numpy().copy(),
                            "cre_out": cx[:Cn].numpy().copy(), "gene_out": gx.numpy().copy()})
    out = {"pred_gene_exp": preds, "embeddings": embs}
    if return_streams:
        out["streams"] = streams
    return out
