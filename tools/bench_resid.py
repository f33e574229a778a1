"""Residual-epilogue GEMM variants at the gene out_proj shape (GPU box only): fp32 vs bf16 residual, in place or not."""
import json
import sys

import torch

sys.path.insert(0, ".")
from variantformer_b200 import ops  # noqa: E402
from variantformer_b200._lib import EPI_BIAS_RESID_F32  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    for (M, N, K) in [(101304, 1536, 1536), (1114624, 512, 512)]:
        a = torch.randn(M, K, device=DEV).bfloat16(); w = torch.randn(N, K, device=DEV).bfloat16()
        bias = torch.randn(N, device=DEV)
        x32 = torch.randn(M, N, device=DEV); x16 = x32.bfloat16(); hb = torch.empty_like(x16)
        st = torch.empty(M, ops.stats_parts(N), 2, device=DEV)
        out = torch.empty(M, N, device=DEV)
        cases = {
            "resid32 -> mirror+stats": lambda: ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=x32, out2=hb, stats_out=st, mirror_only=True),
            "resid16 -> mirror+stats": lambda: ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=x16, out2=hb, stats_out=st, mirror_only=True),
            "resid16 in place -> mirror+stats": lambda: ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=x16, out2=x16, stats_out=st, mirror_only=True),
            "resid32 in place -> out32+mirror+stats": lambda: ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=x32, out=x32, out2=hb, stats_out=st),
            "resid32 -> out32": lambda: ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, resid=x32, out=out),
            "no resid -> mirror": lambda: ops.gemm(a, w, EPI_BIAS_RESID_F32, bias=bias, out2=hb, mirror_only=True),
        }
        for name, fn in cases.items():
            ms = timeit(fn)
            print(json.dumps(dict(M=M, N=N, K=K, case=name, ms=round(ms, 4), tflops=round(2.0 * M * N * K / ms / 1e9, 1))))


if __name__ == "__main__":
    main()
