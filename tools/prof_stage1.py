"""Stage 1 of one benchmark slab (8 genes x 1024 CRE windows + 8 gene windows), a few times: for `ncu -k regex:bpe|encode`.
GPU box only.    python tools/prof_stage1.py [repeats]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from variantformer_b200.stage1 import Genome, SampleVariants, WindowTokenizer  # noqa: E402
from variantformer_b200.pipeline import HotPath  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    chroms, var, sets = bench.make_workload(1234, 1, 8, 1024, 63)
    dev = torch.device("cuda")
    hot = HotPath.__new__(HotPath)
    hot.engine = type("E", (), {"device": dev})()
    hot.genome = Genome.from_arrays(chroms, dev)
    hot.tok = WindowTokenizer(dev)
    hot.nb, hot.up, hot.down, hot.max_length, hot.max_chunks = 50, 1000, 300000, 200, 200
    variants = SampleVariants(var, dev)
    for _ in range(n):
        t = hot.tokenize(sets[0], variants)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); t = hot.tokenize(sets[0], variants); ev[1].record(); torch.cuda.synchronize()
    print("stage 1 of one slab:", round(ev[0].elapsed_time(ev[1]), 3), "ms (incl. host bookkeeping gaps); tokens", int(t["lens"][0].sum()), "+", int(t["lens"][1].sum()))


if __name__ == "__main__":
    main()
