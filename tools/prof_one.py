"""Launch ONE kernel configuration a few times (for `ncu --launch-skip 2 --launch-count 1`).  GPU box only.

    python tools/prof_one.py gemm M N K epi [full]        full = engine configuration (mirror + stats / LN fold)
    python tools/prof_one.py attn {gself|gcross|cself|rcre|rgene}
"""
import sys

import torch

sys.path.insert(0, ".")
from variantformer_b200 import ops  # noqa: E402
from variantformer_b200._lib import EPI_BIAS_GEGLU_BF16, EPI_BIAS_RESID_F32  # noqa: E402

DEV = "cuda"


def gemm(M, N, K, epi, full):
    a = torch.randn(M, K, device=DEV).bfloat16(); w = torch.randn(N, K, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    n_out = N // 2 if epi == EPI_BIAS_GEGLU_BF16 else N
    resid = torch.randn(M, N, device=DEV) if epi == EPI_BIAS_RESID_F32 else None
    out = torch.empty(M, n_out, device=DEV, dtype=torch.float32 if epi in (2, 3) else torch.bfloat16)
    kw = {}
    if full and epi in (2, 3):
        kw = dict(out2=torch.empty(M, N, device=DEV, dtype=torch.bfloat16),
                  stats_out=torch.empty(M, ops.stats_parts(N), 2, device=DEV))
    elif full:
        st = torch.rand(M, 12, 2, device=DEV) / 12 + 1.0 / 12
        st[:, :, 1] += K / 12
        kw = dict(ln=(st, bias, K, 1e-5))
    for _ in range(3):
        ops.gemm(a, w, epi, bias=bias, resid=resid, out=out, **kw)
    torch.cuda.synchronize()


ATTN = {  # name: (lens_q, lens_k, H, hd, alibi)
    "rcre": ([97] * 8192, None, 8, 64, False), "rgene": ([200] * 1600, None, 8, 64, False),
    "cself": ([1024] * 8, None, 32, 48, True), "gself": ([201] * 504, None, 32, 48, True),
    "gcross": ([12663] * 8, [1024] * 8, 32, 48, False),
}


def attn(name, kb):
    lens_q, lens_k, H, hd, alibi = ATTN[name]
    d = H * hd
    nq = sum(lens_q); lk = lens_k or lens_q; nk = sum(lk)
    q = torch.randn(nq, d, device=DEV).bfloat16(); k = torch.randn(nk, d, device=DEV).bfloat16()
    v = torch.randn(nk, d, device=DEV).bfloat16(); o = torch.empty(nq, d, device=DEV, dtype=torch.bfloat16)
    slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV) if alibi else None
    cq, ck = ops.cu_seqlens(lens_q, DEV), ops.cu_seqlens(lk, DEV)
    slots = ops.SlotMap(lens_q, DEV, k_lens=lens_k)
    for _ in range(3):
        ops.attention_mc(q, k, v, slots, H, hd, slopes, out=o)
    torch.cuda.synchronize()


if __name__ == "__main__":
    if sys.argv[1] == "gemm":
        M, N, K, epi = map(int, sys.argv[2:6])
        gemm(M, N, K, epi, len(sys.argv) > 6)
    else:
        attn(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 64)
