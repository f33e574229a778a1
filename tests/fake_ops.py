"""CPU stand-in for variantformer_b200.ops used ONLY by tests/test_engine_host_logic.py.

It mimics each C-ABI kernel's contract (layouts, strides, in-place outputs, tile-interleaved GeGLU) with plain
torch on the CPU so that the engine's *host-side bookkeeping* (unpadding, tile maps, gather indices, tissue
stacking, label counts) can be checked against the oracle without a GPU.  It is not a fallback: nothing in
variantformer_b200/ imports it, and the product raises without libvf_b200.so + a B200.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from variantformer_b200._lib import (EPI_BIAS_BF16, EPI_BIAS_F32, EPI_BIAS_GEGLU_BF16, EPI_BIAS_GELU_BF16,
                                     EPI_BIAS_RESID_F32)

LAUNCHES = 0


def _store(out, val):
    out.copy_(val.to(out.dtype))
    return out


def gemm(a, w, epilogue, bias=None, resid=None, out=None, out2=None, ln=None, stats_out=None, mirror_only=False):
    y = a.float() @ w.float().t()
    if ln is not None:                       # LayerNorm fold: rstd * (acc - mean * colsum)
        stats, colsum, dim, eps = ln
        st = stats.sum(1)
        mean = st[:, 0] / dim
        rstd = torch.rsqrt((st[:, 1] / dim - mean * mean).clamp_min(0) + eps)
        y = rstd[:, None] * (y - mean[:, None] * colsum[None, :])
    if bias is not None:
        y = y + bias
    if epilogue == EPI_BIAS_GEGLU_BF16:
        N = w.shape[0]
        t = y.view(y.shape[0], N // 256, 2, 128)
        y = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(y.shape[0], N // 2)
        dt = torch.bfloat16
    elif epilogue == EPI_BIAS_GELU_BF16:
        y = F.gelu(y); dt = torch.bfloat16
    elif epilogue == EPI_BIAS_BF16:
        dt = torch.bfloat16
    else:
        if epilogue == EPI_BIAS_RESID_F32 and resid is not None:
            y = y + resid.float()
        dt = torch.float32
    if out2 is not None:
        out2.copy_(y.to(torch.bfloat16))
    if stats_out is not None:
        stats_out.zero_()
        stats_out[:, 0].copy_(torch.stack([y.sum(1), (y * y).sum(1)], 1))
    if mirror_only:
        return out2
    if out is None:
        return y.to(dt)
    assert out.dtype == dt
    return _store(out, y)


class TileMap:
    def __init__(self, q_lens, block_m, device, k_lens=None):
        self.block_m = block_m
        self.lens = np.asarray(q_lens)


def cu_seqlens(lens, device):
    cu = np.zeros(len(lens) + 1, np.int32)
    np.cumsum(np.asarray(lens, np.int64), out=cu[1:])
    return torch.from_numpy(cu)


def attention(q, k, v, cu_q, cu_k, tiles, heads, head_dim, slopes=None, out=None):
    assert (np.diff(cu_q.numpy()) == tiles.lens).all(), "tile map does not describe the query sequences"
    res = torch.empty(q.shape[0], heads * head_dim)
    cq, ck = cu_q.tolist(), cu_k.tolist()
    for b in range(len(cq) - 1):
        qq = q[cq[b]:cq[b + 1]].float().reshape(-1, heads, head_dim)
        kk = k[ck[b]:ck[b + 1]].float().reshape(-1, heads, head_dim)
        vv = v[ck[b]:ck[b + 1]].float().reshape(-1, heads, head_dim)
        s = torch.einsum("thd,shd->hts", qq, kk) / math.sqrt(head_dim)
        if slopes is not None:
            sq, sk = qq.shape[0], kk.shape[0]
            i = torch.arange(sq)[:, None]; j = torch.arange(sk)[None, :]
            s = s - slopes[:, None, None] * (i + sk - sq - j).abs()[None]
        res[cq[b]:cq[b + 1]] = torch.einsum("hts,shd->thd", s.softmax(-1), vv).reshape(-1, heads * head_dim)
    return _store(out, res) if out is not None else res.bfloat16()


class SlotMap:
    def __init__(self, q_lens, device, k_lens=None):
        self.q_lens = np.asarray(q_lens)
        self.k_lens = self.q_lens if k_lens is None else np.asarray(k_lens)
        self.units = None

    @classmethod
    def from_units(cls, units, device):
        self = cls.__new__(cls)
        self.units = np.asarray(units, np.int64).reshape(-1, 5)
        return self


def attention_mc(q, k, v, slots, heads, head_dim, slopes=None, out=None):
    if slots.units is None:
        return attention(q, k, v, cu_seqlens(slots.q_lens, None), cu_seqlens(slots.k_lens, None),
                         TileMap(slots.q_lens, 128, None), heads, head_dim, slopes, out)
    res = torch.zeros(q.shape[0], heads * head_dim)
    for qrow, nrows, krow, Sk, qpos0 in slots.units.tolist():        # explicit slot records (include/vf_b200.h)
        qq = q[qrow:qrow + nrows].float().reshape(-1, heads, head_dim)
        kk = k[krow:krow + Sk].float().reshape(-1, heads, head_dim)
        vv = v[krow:krow + Sk].float().reshape(-1, heads, head_dim)
        sc = torch.einsum("thd,shd->hts", qq, kk) / math.sqrt(head_dim)
        if slopes is not None:
            i = qpos0 + torch.arange(nrows)[:, None]; j = torch.arange(Sk)[None, :]
            sc = sc - slopes[:, None, None] * (i - j).abs()[None]
        res[qrow:qrow + nrows] = torch.einsum("hts,shd->thd", sc.softmax(-1), vv).reshape(-1, heads * head_dim)
    return _store(out, res) if out is not None else res.bfloat16()


def label_attention(q, kv9, logc, row_seq, heads, head_dim, out=None):
    D = heads * head_dim
    qq = q.float().reshape(-1, heads, head_dim)
    k9 = kv9[:, :D].reshape(9, heads, head_dim); v9 = kv9[:, D:].reshape(9, heads, head_dim)
    s = torch.einsum("rhd,chd->rhc", qq, k9) / math.sqrt(head_dim) + logc[row_seq.long()][:, None, :]
    res = torch.einsum("rhc,chd->rhd", s.softmax(-1), v9).reshape(-1, D)
    return _store(out, res) if out is not None else res.bfloat16()


def layernorm(x, gamma, beta, eps=1e-5, gelu=False, out=None):
    y = F.layer_norm(x, (x.shape[1],), gamma, beta, eps)
    if gelu:
        y = F.gelu(y)
    return _store(out, y) if out is not None else y.bfloat16()


def stats_parts(n):
    return 2 * ((n + 255) // 256)


def rowstats(x, stats=None, out_bf16=None):
    st = torch.stack([x.sum(1), (x * x).sum(1)], 1)[:, None, :]
    if out_bf16 is not None:
        out_bf16.copy_(x.bfloat16())
    return _store(stats, st) if stats is not None else st


def window_lengths(pad_mask_u8):
    return (pad_mask_u8 == 0).sum(1).int()


def compact_tokens(tokens_i32, pad_mask_u8, cu, n_tok):
    keep = pad_mask_u8 == 0
    n, L = tokens_i32.shape
    ids = tokens_i32[keep]; pos = torch.arange(L, dtype=torch.int32).expand(n, L)[keep]
    assert ids.numel() == n_tok
    return ids, pos


def embed_tokens(ids, pos, emb, pe):
    x = emb[ids.long()]
    return x + pe[pos.long()] if pe is not None else x


def center_rows(x, pivot=None, stats=None, out_bf16=None):
    mean = x.mean(1)
    x.sub_(mean[:, None])
    st = torch.stack([x.sum(1), (x * x).sum(1)], 1)[:, None, :]
    if out_bf16 is not None:
        out_bf16.copy_(x.bfloat16())
    pivot = _store(pivot, mean) if pivot is not None else mean
    return pivot, (_store(stats, st) if stats is not None else st)


def uncenter_rows(x, pivot, idx=None, want_f32=True, want_bf16=False, out_bf16=None):
    y = x + (pivot if idx is None else pivot[idx.long()])[:, None]
    ob = _store(out_bf16, y) if out_bf16 is not None else (y.bfloat16() if want_bf16 else None)
    return (y if want_f32 else None), ob


def masked_meanpool(x, cu, n_win, want_f32=False, pivot=None):
    lens = torch.from_numpy(np.diff(cu.numpy())).long()
    seg = torch.repeat_interleave(torch.arange(n_win), lens)
    if pivot is not None:
        x = x + pivot[:, None]
    p = torch.zeros(n_win, x.shape[1]).index_add_(0, seg, x) / lens[:, None]
    return (p.bfloat16(), p) if want_f32 else p.bfloat16()


def gather_rows(table_a, table_b, idx, want_f32=True, want_bf16=False):
    rows = [table_a[i] if i >= 0 else table_b[-i - 1] for i in idx.tolist()]
    y = torch.stack(rows)
    return (y if want_f32 else None), (y.bfloat16() if want_bf16 else None)


def head_out(h_bf16, w, b, softplus=True):
    y = h_bf16.float() @ w + b
    return F.softplus(y) if softplus else y


def cast_bf16(x):
    return x.bfloat16()
