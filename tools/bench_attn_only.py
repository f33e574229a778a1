"""Attention-only micro-benchmark (legacy vs tcgen05) for profiling under ncu."""
import sys

import torch

sys.path.insert(0, ".")
from variantformer_b200 import ops  # noqa: E402

DEV = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "cross"
cfgs = {"cross": ([12663] * 8, [1024] * 8, 32, 48, False), "gself": ([201] * 504, None, 32, 48, True),
        "cself": ([1024] * 8, None, 32, 48, True)}
lens_q, lens_k, H, hd, alibi = cfgs[which]
d = H * hd
lk = lens_k or lens_q
q = torch.randn(sum(lens_q), d, device=DEV).bfloat16(); k = torch.randn(sum(lk), d, device=DEV).bfloat16()
v = torch.randn(sum(lk), d, device=DEV).bfloat16(); o = torch.empty(sum(lens_q), d, device=DEV, dtype=torch.bfloat16)
slopes = torch.tensor([2 ** (-8 * (h + 1) / H) for h in range(H)], device=DEV) if alibi else None
cq, ck = ops.cu_seqlens(lens_q, DEV), ops.cu_seqlens(lk, DEV)
items = ops.TileMap(lens_q, ops.TC_BLOCK_M, DEV, k_lens=lens_k)
for _ in range(3):
    ops.attention_tc(q, k, v, cq, ck, items, H, hd, slopes, out=o, key_block=int(sys.argv[2]) if len(sys.argv) > 2 else 64)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention_tc(q, k, v, cq, ck, items, H, hd, slopes, out=o, key_block=int(sys.argv[2]) if len(sys.argv) > 2 else 64); e1.record(); torch.cuda.synchronize()
print(which, "tc ms", e0.elapsed_time(e1))
