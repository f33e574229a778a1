"""The survey's kernel-to-beat, timed on the same box: flash_attn 2.8.x (library) on the five attention shapes of the
hot path (seq2gene/modules/layers.py:344-351, 372-467; seq2reg/modules.py:159-171) next to vf_attention_mc_varlen, and
cuBLASLt (torch.matmul) next to vf_gemm_bf16_ln on the GEMM shapes.  GPU box only; L2 flushed between iterations.

    python tools/bench_flash_attn.py [out.json]

flash_attn is NOT a dependency of the product; this script is the only place that imports it.
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from variantformer_b200 import ops  # noqa: E402

DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


SHAPES = [  # name, q lens, k lens (None = self), heads, head_dim, alibi
    ("seq2reg self, 8192 CRE windows x 97 tok", [97] * 8192, None, 8, 64, False),
    ("seq2reg self, 1600 gene chunks x 200 tok", [200] * 1600, None, 8, 64, False),
    ("CRE self + ALiBi, 8 x 1024", [1024] * 8, None, 32, 48, True),
    ("gene self + ALiBi, 504 x 201", [201] * 504, None, 32, 48, True),
    ("gene->CRE cross, 8 x (63*201) x 1024", [12663] * 8, [1024] * 8, 32, 48, False),
]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/flash_attn_vs_mc.json"
    res = []
    try:
        from flash_attn import flash_attn_varlen_func
        import flash_attn
        fa_version = flash_attn.__version__
    except Exception as e:                                     # noqa: BLE001
        flash_attn_varlen_func, fa_version = None, f"unavailable: {e}"
    print("flash_attn", fa_version, flush=True)
    for name, lens_q, lens_k, H, hd, alibi in SHAPES:
        d = H * hd
        lk = lens_k or lens_q
        nq, nk = sum(lens_q), sum(lk)
        g = torch.Generator(device=DEV).manual_seed(1)
        q = torch.randn(nq, d, device=DEV, generator=g).bfloat16()
        k = torch.randn(nk, d, device=DEV, generator=g).bfloat16()
        v = torch.randn(nk, d, device=DEV, generator=g).bfloat16()
        o = torch.empty(nq, d, device=DEV, dtype=torch.bfloat16)
        slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV) if alibi else None
        fl = sum(4.0 * a * b * d for a, b in zip(lens_q, lk))
        slots = ops.SlotMap(lens_q, DEV, k_lens=lens_k)
        ms = timeit(lambda: ops.attention_mc(q, k, v, slots, H, hd, slopes, out=o))
        row = dict(shape=name, heads=H, head_dim=hd, alibi=alibi, mc_ms=ms, mc_tflops=fl / ms / 1e9)
        if flash_attn_varlen_func is not None:
            cq, ck = ops.cu_seqlens(lens_q, DEV), ops.cu_seqlens(lk, DEV)
            q3, k3, v3 = q.view(nq, H, hd), k.view(nk, H, hd), v.view(nk, H, hd)
            try:
                fn = lambda: flash_attn_varlen_func(q3, k3, v3, cq, ck, max(lens_q), max(lk), causal=False,  # noqa: E731
                                                    alibi_slopes=slopes)
                ref = fn().reshape(nq, d)
                ms_fa = timeit(fn)
                ops.attention_mc(q, k, v, slots, H, hd, slopes, out=o)
                torch.cuda.synchronize()
                row.update(flash_attn_ms=ms_fa, flash_attn_tflops=fl / ms_fa / 1e9, speedup=ms_fa / ms,
                           max_abs_diff_vs_flash_attn=float((o.float() - ref.float()).abs().max()))
            except Exception as e:                             # noqa: BLE001
                row["flash_attn_error"] = str(e)[:200]
        res.append(row)
        print(json.dumps(row), flush=True)
        del q, k, v, o
    json.dump({"flash_attn_version": fa_version, "gpu": torch.cuda.get_device_name(0), "rows": res},
              open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
