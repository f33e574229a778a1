"""Micro-benchmarks of the individual kernels at the canonical shapes.  GPU box only.
Prints one line per case: achieved TFLOP/s (or GB/s) from CUDA-event timing, L2 flushed between iterations."""
import json
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import os  # noqa: E402
from variantformer_b200 import _lib  # noqa: E402
if os.environ.get("VF_BENCH_LIB"):                  # A/B experiments: a differently compiled build of the same library
    _lib.LIB_PATH = os.path.abspath(os.environ["VF_BENCH_LIB"])
from variantformer_b200 import ops  # noqa: E402
from variantformer_b200._lib import EPI_BIAS_BF16, EPI_BIAS_GEGLU_BF16, EPI_BIAS_RESID_F32  # noqa: E402

DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    res = []
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    for (M, N, K, epi, name) in [] if only == "attention" else [
        (101304, 4608, 1536, EPI_BIAS_BF16, "gene Wqkv x8 genes"),
        (101304, 1536, 1536, EPI_BIAS_RESID_F32, "gene out_proj x8"),
        (101304, 2048, 1536, EPI_BIAS_GEGLU_BF16, "gene geglu1 x8"),
        (101304, 1536, 1024, EPI_BIAS_RESID_F32, "gene geglu2 x8"),
        (12663, 4608, 1536, EPI_BIAS_BF16, "gene Wqkv x1"),
        (8192, 4608, 1536, EPI_BIAS_BF16, "cre Wqkv x8"),
        (8192, 3072, 1536, EPI_BIAS_BF16, "gene Wkv x8"),
        (1114624, 1536, 512, EPI_BIAS_BF16, "seq2reg Wqkv x8"),
        (1114624, 512, 512, EPI_BIAS_RESID_F32, "seq2reg out_proj x8"),
        (1114624, 2048, 512, EPI_BIAS_GEGLU_BF16, "seq2reg geglu1 x8"),
        (1114624, 512, 1024, EPI_BIAS_RESID_F32, "seq2reg geglu2 x8"),
    ]:
        a = torch.randn(M, K, device=DEV).bfloat16(); w = torch.randn(N, K, device=DEV).bfloat16()
        bias = torch.randn(N, device=DEV)
        n_out = N // 2 if epi == EPI_BIAS_GEGLU_BF16 else N
        resid = torch.randn(M, N, device=DEV) if epi == EPI_BIAS_RESID_F32 else None
        out = torch.empty(M, n_out, device=DEV, dtype=torch.float32 if epi == EPI_BIAS_RESID_F32 else torch.bfloat16)
        ms = timeit(lambda: ops.gemm(a, w, epi, bias=bias, resid=resid, out=out))
        # the configuration the engine runs: fp32 epilogues also write the bf16 mirror and the row statistics,
        # bf16 epilogues apply the folded LayerNorm
        if epi == EPI_BIAS_RESID_F32:
            st = torch.empty(M, ops.stats_parts(N), 2, device=DEV)
            o2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
            # (in place on the fp32 stream, as the engine calls it)
            ms_full = timeit(lambda: ops.gemm(a, w, epi, bias=bias, resid=resid, out=resid, out2=o2, stats_out=st))
            # the out_proj configuration: bf16 residual (the stream's mirror), in place, mirror + statistics only
            ms_r16 = timeit(lambda: ops.gemm(a, w, epi, bias=bias, resid=o2, out2=o2, stats_out=st, mirror_only=True))
            del o2
        else:
            st = torch.rand(M, 12, 2, device=DEV) / 12 + 1.0 / 12
            st[:, :, 1] += K / 12
            ms_full = timeit(lambda: ops.gemm(a, w, epi, bias=bias, out=out, ln=(st, bias, K, 1e-5)))
        if epi != EPI_BIAS_RESID_F32:
            ms_r16 = None
        ms_t = timeit(lambda: torch.matmul(a, w.t()))
        tf = 2.0 * M * N * K / ms / 1e9
        res.append(dict(kernel="gemm", name=name, M=M, N=N, K=K, epi=epi, ms=ms, tflops=tf, ms_engine_cfg=ms_full,
                        tflops_engine_cfg=2.0 * M * N * K / ms_full / 1e9, ms_resid16_mirror_only=ms_r16, cublas_ms=ms_t,
                        cublas_tflops=2.0 * M * N * K / ms_t / 1e9))
        print(json.dumps(res[-1])); sys.stdout.flush()
        del a, w, out, resid
    # attention
    for (name, lens_q, lens_k, H, hd, alibi, bm) in [] if only == "gemm" else [
        ("seq2reg self cre x8", [97] * 8192, None, 8, 64, False, 64),
        ("seq2reg self gene x8", [200] * 1600, None, 8, 64, False, 64),
        ("cre self x8", [1024] * 8, None, 32, 48, True, 128),
        ("gene self x8", [201] * 504, None, 32, 48, True, 64),
        ("gene cross x8", [12663] * 8, [1024] * 8, 32, 48, False, 128),
    ]:
        d = H * hd
        nq = sum(lens_q); lk = lens_k or lens_q; nk = sum(lk)
        q = torch.randn(nq, d, device=DEV).bfloat16(); k = torch.randn(nk, d, device=DEV).bfloat16()
        v = torch.randn(nk, d, device=DEV).bfloat16(); o = torch.empty(nq, d, device=DEV, dtype=torch.bfloat16)
        slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV) if alibi else None
        cq, ck = ops.cu_seqlens(lens_q, DEV), ops.cu_seqlens(lk, DEV)
        fl = sum(4.0 * a * b * d for a, b in zip(lens_q, lk))
        try:
            slots = ops.SlotMap(lens_q, DEV, k_lens=lens_k)
            ms = timeit(lambda: ops.attention_mc(q, k, v, slots, H, hd, slopes, out=o))
            res.append(dict(kernel="attention_mc", name=name, ms=ms, tflops=fl / ms / 1e9))
            print(json.dumps(res[-1])); sys.stdout.flush()
        except Exception as e:
            print("attention_mc failed:", e)
    # layernorm bandwidth
    x = torch.randn(101304, 1536, device=DEV); g = torch.ones(1536, device=DEV); b = torch.zeros(1536, device=DEV)
    o = torch.empty(101304, 1536, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.layernorm(x, g, b, out=o))
    res.append(dict(kernel="layernorm", ms=ms, gbs=x.numel() * 6 / ms / 1e6)); print(json.dumps(res[-1]))
    json.dump(res, open("gpurun_out/bench_kernels.json", "w"), indent=1)


if __name__ == "__main__":
    main()
