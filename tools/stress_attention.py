"""Determinism stress of the attention kernel (seq_len <= 128: one-tile sequences — with ops.PAIR_UNRELATED_TILES two
of them share a work item and stream their own keys; longer: tiles of one sequence share the K/V stream):
the same launch N times, every output compared bit for bit with the first; a second stream keeps unrelated kernels
running to perturb the timing.  GPU box only.   python tools/stress_attention.py [launches] [seq_len | cross | ragged]
"ragged": 8192 sequences of 1..200 rows (split items whose slots have different numbers of key blocks).
"""
import os
import sys

import torch

sys.path.insert(0, ".")
from variantformer_b200 import _lib  # noqa: E402

if os.environ.get("VF_BENCH_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["VF_BENCH_LIB"])
from variantformer_b200 import ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    mode = sys.argv[2] if len(sys.argv) > 2 else "97"
    L = int(mode) if mode.isdigit() else 97
    cross = mode == "cross"                                   # stacked gene->CRE cross-attention: pairs + left-over tiles
    H, hd = (32, 48) if cross else (8, 64)
    lens = [12663] * 8 if cross else [L] * (8192 if L <= 128 else 1600)
    if mode == "ragged":
        import numpy as np
        lens = np.random.default_rng(0).integers(1, 201, 8192).tolist()
    k_lens = [1024] * 8 if cross else None
    d = H * hd
    tot = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(tot, d, device="cuda", generator=g).bfloat16()
    k, v = (torch.randn(sum(k_lens or lens), d, device="cuda", generator=g).bfloat16() for _ in range(2))
    slots = ops.SlotMap(lens, "cuda", k_lens=k_lens)
    ref = ops.attention_mc(q, k, v, slots, H, hd, None).clone()
    out = torch.empty_like(ref)
    bad = torch.zeros((), dtype=torch.int64, device="cuda")
    side = torch.cuda.Stream()
    a = torch.randn(4096, 4096, device="cuda"); b = torch.randn(1 << 26, device="cuda")
    for i in range(n):
        if i % 7 == 0:
            with torch.cuda.stream(side):               # unrelated traffic: a GEMM and a streaming copy
                (a @ a).sum(); b.add_(1.0)
        ops.attention_mc(q, k, v, slots, H, hd, None, out=out)
        bad += (out.view(torch.int16) != ref.view(torch.int16)).any().to(torch.int64)
    torch.cuda.synchronize()
    print(f"launches {n} {mode if not mode.isdigit() else 'seq_len ' + str(L)} pairing={ops.PAIR_UNRELATED_TILES} "
          f"items={slots.n_items}: {int(bad.item())} differ from the first", flush=True)


if __name__ == "__main__":
    main()
