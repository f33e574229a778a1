"""bench.py — gene x tissue predictions/s of the VariantFormer batched-inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: the oracle port on the host cores)

One step = one pass of the FULL hot path over one slab of synthetic input per GPU:
  stage 1  genotype application + reverse complement + BPE-500 for B x C CRE windows and B gene windows
           (genome + the sample's variants resident in HBM),
  stage 2  seq2reg encoders over every window (valid tokens only),
  stage 3  seq2gene CombinedModulator, 24 CRE + 25 gene layers, T tissues stacked per gene,
  stage 4  expression head,
at the configuration BASELINE.json's metric is quoted on (config 3: full hierarchical model, C=1024 CRE windows,
G=200 gene chunks, T=63 tissues per gene; B genes per slab).  Weak scaling: every rank owns its own slab; the
only collective is the final gather of expression + embeddings (inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gene_x_tissue_predictions_per_s"

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
# communicator), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _RESULT_OUT.write(json.dumps(obj) + "\n")
    _RESULT_OUT.flush()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(p, source="measured")
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------------------------
# synthetic workload: seeded genome, one sample's variants, genes with C CREs and a 301 kb window
# ----------------------------------------------------------------------------------------------------------------
def make_workload(seed, n_sets, B, C, T, chrom_len=24_000_000):
    from variantformer_b200.pipeline import GeneSpec
    from variantformer_b200.utils import synth
    rng = np.random.default_rng(seed)
    chrom = synth.make_chromosome(rng, chrom_len)
    var = synth.make_variants(rng, chrom)
    sets = []
    for _ in range(n_sets):
        genes = []
        for _ in range(B):
            lay = synth.make_gene_layout(rng, chrom_len, C, body_len=int(rng.integers(320_000, 900_000)))
            genes.append(GeneSpec("chr1", lay["start"], lay["end"], lay["strand"], lay["cre_start"], lay["cre_end"],
                                  lay["labels"], list(range(T))))
        sets.append(genes)
    return {"chr1": chrom}, {"chr1": var}, sets


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (oracle/ is test infrastructure; this is one of the two places allowed to execute it)
# ----------------------------------------------------------------------------------------------------------------
def cpu_sample(sd_cpu, cfg, hp, chroms, var, gene, C, n_tissues):
    """Stage 1 (C oracle) + stages 2-4 (fp32 torch oracle, REFERENCE schedule) for ONE gene x n_tissues on the host."""
    import torch
    from oracle import model_fp32, stage1 as O
    from variantformer_b200.stage1 import cre_window, gene_window
    t0 = time.perf_counter()
    bpe = O.OracleBPE()
    chrom = chroms[gene.chrom]; v = var[gene.chrom]
    lens = np.array([len(a) for a in v["alt"]], np.int32); pool = np.frombuffer(b"".join(v["alt"]), np.uint8)
    aoff = np.cumsum(lens) - lens
    minus = gene.strand == "-"

    def window(a0, a1):
        lo, hi = np.searchsorted(v["pos"], a0), np.searchsorted(v["pos"], a1)
        s = O.apply_variants(chrom, a0, a1, v["pos"][lo:hi], v["ref_len"][lo:hi], aoff[lo:hi], lens[lo:hi],
                             v["gt"][lo:hi], pool)
        return O.reverse_complement(s) if minus else s
    order = np.argsort(gene.cre_start, kind="stable")
    order = order[::-1] if minus else order
    toks, masks = [], []
    for i in order:
        a0, a1 = cre_window(gene.cre_start[i], gene.cre_end[i], 50)
        o, m = O.adjust_length(bpe.encode(window(a0, a1)), 200)
        toks.append(o); masks.append(m)
    a0, a1 = gene_window(gene.start, gene.end, gene.strand, 1000, 300000)
    gt, gm = O.chunkify(bpe.encode(window(a0, a1)), 200, 200)
    t_stage1 = time.perf_counter() - t0
    batch = {"cre_sequences": [torch.from_numpy(np.stack(toks)).long().unsqueeze(1)],
             "cre_attention_masks": [torch.from_numpy(np.stack(masks)).unsqueeze(1)],
             "gene_embeddings": [torch.from_numpy(gt).long().unsqueeze(1)],
             "gene_attention_masks": [torch.from_numpy(gm).unsqueeze(1)],
             "tissue_context": [torch.arange(n_tissues)],
             "ref_cre_labels": [torch.from_numpy(np.asarray(gene.cre_labels)[order].copy())]}
    t0 = time.perf_counter()
    model_fp32.predict_step(sd_cpu, cfg, hp, batch, schedule="reference")
    return t_stage1, time.perf_counter() - t0


def cpu_baseline(sd_cpu, cfg, hp, chroms, var, gene, C, T):
    """Bounded sample: one gene timed at 1 and 2 tissues.  The reference schedule repeats the CRE and gene streams
    once per tissue (model_combined_modulator.py:622-649), so cost(T) = base + T*marginal exactly; the T-tissue
    rate of the configured workload is T / (base + T*marginal)."""
    import torch
    torch.set_num_threads(os.cpu_count())
    torch.set_float32_matmul_precision("highest")
    s1, t1 = cpu_sample(sd_cpu, cfg, hp, chroms, var, gene, C, 1)
    _, t2 = cpu_sample(sd_cpu, cfg, hp, chroms, var, gene, C, 2)
    marginal = max(t2 - t1, 1e-9); base = max(t1 - marginal, 0.0) + s1
    value = T / (base + T * marginal)
    return {"value": value, "unit": "predictions/s", "cores": os.cpu_count(), "kind": "port",
            "sample": (f"oracle port (C stage 1 + torch fp32 reference schedule), 1 gene C={C} G=200 timed at T=1 "
                       f"({t1:.1f}s) and T=2 ({t2:.1f}s), stage1 {s1:.2f}s; value = {T}/(base+{T}*marginal), "
                       f"base={base:.1f}s marginal={marginal:.2f}s; bcftools/samtools subprocess cost not timeable (absent)"),
            "t_T1_s": t1, "t_T2_s": t2, "t_stage1_s": s1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from variantformer_b200.utils import random_init
    cfg, hp = dict(random_init.V4_PCG_MODEL), dict(random_init.SEQ2REG_HP)
    sd = random_init.make_state_dict(cfg, hp, seed=0)
    chroms, var, sets = make_workload(1234, 1, 1, args.cre, args.tissues, chrom_len=8_000_000)
    vals = []
    warm = min(args.warmup, 1)          # a CPU has no clocks or caches worth more than one warm-up pass of ~30 s
    t_all = time.perf_counter()
    for i in range(warm + args.steps):
        r = cpu_baseline(sd, cfg, hp, chroms, var, sets[0][0], args.cre, args.tissues)
        if i >= warm:
            vals.append(r)
    total = time.perf_counter() - t_all
    v = float(np.mean([x["value"] for x in vals]))
    ms = 1e3 * float(np.mean([x["t_T1_s"] + x["t_T2_s"] for x in vals]))
    last = dict(vals[-1], value=v)
    emit(({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "predictions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "CPU arm: rank 0 only, host cores"},
        "cpu_baseline": last, "e2e": {"value": v, "unit": "predictions/s", "h2d_bytes_per_step": 0,
                                      "d2h_bytes_per_step": 0}, "wall_s": total}))


def workload_name(args):
    return (f"config3-slab: full hierarchical seq2reg(6L,512d)+seq2gene(25L,1536d,32h) + stage-1 tokenisation, "
            f"{args.genes_per_step} genes/GPU/step x C={args.cre} CRE windows x G=200 gene chunks x T={args.tissues} tissues")


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genes-per-step", type=int, default=8)
    ap.add_argument("--cre", type=int, default=1024)
    ap.add_argument("--tissues", type=int, default=63)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", type=int, default=25, help="debug only: seq2gene depth (25 = the benchmark)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from variantformer_b200 import ops, parallel
    from variantformer_b200.engine import Engine
    from variantformer_b200.pipeline import HotPath
    from variantformer_b200.stage1 import Genome, SampleVariants
    from variantformer_b200.utils import random_init

    rank, world, local = parallel.init_from_env()
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cfg = dict(random_init.V4_PCG_MODEL, num_layers=args.layers); hp = dict(random_init.SEQ2REG_HP)
    sd = random_init.make_state_dict(cfg, hp, seed=0, device=dev)
    engine = Engine(sd, cfg, hp, device=dev)
    B, C, T = args.genes_per_step, args.cre, args.tissues
    n_sets = 3
    chroms, var, sets = make_workload(1234 + rank, n_sets, B, C, T)
    hot = HotPath(engine, Genome.from_arrays(chroms, dev))
    variants = SampleVariants(var, dev)
    preds_per_step = B * T
    counts = [preds_per_step] * world

    def run_steps(first, n, to_host, gather=True):
        """n consecutive steps through the pipelined public API (stage 1 of slab i+1 overlaps the model of slab i)."""
        last = None
        for last in hot.predict_pipelined((sets[(first + i) % n_sets] for i in range(n)), variants, to_host=to_host):
            if world > 1 and not to_host and gather:    # the one collective: final gather of expression + embeddings
                parallel.gather_rows(last[0], counts, world, rank); parallel.gather_rows(last[1], counts, world, rank)
        return last

    run_steps(0, args.warmup, False)
    torch.cuda.synchronize()

    # ---- timed region: device-resident leg (`value`) ----
    sampler = ClockSampler(local); sampler.start()
    launches0 = ops.LAUNCHES
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pred, emb, err = run_steps(0, args.steps, False)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = ops.LAUNCHES - launches0
    assert int(err.item()) == 0, "stage-1 kernel flagged an error"
    assert bool(torch.isfinite(pred).all()) and bool(torch.isfinite(emb).all()), "non-finite outputs"
    t = torch.tensor([elapsed_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * preds_per_step * args.steps / (elapsed_ms / 1e3)

    # ---- end-to-end leg: host window tables in, numpy results out, every step ----
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    p_np, e_np = run_steps(0, args.steps, True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * preds_per_step * args.steps / float(t.item())
    h2d = int(B * C * (8 + 8 + 4 + 4 + 4 + 1 + 8) + B * 40 + B * T * 8)      # window tables, labels, tissue ids
    d2h = int(p_np.nbytes + e_np.nbytes + 4 * (B * C + B))                   # results + token counts

    out = {
        "metric": METRIC, "value": value, "unit": "predictions/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "weights": "random-init vf_model.yaml v4_pcg (seed 0)",
                   "l2": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; 3 gene sets rotate",
                   "parallelism": f"dp{world} (gene x sample sharding, final all_gather only)",
                   "pipelining": "stage 1 + host bookkeeping of slab i+1 on a side stream while slab i's model runs; "
                                 "CRE stack on its own stream next to the gene stack",
                   "algorithmic_tflop_per_step": None},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "predictions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
    }
    # ---- per-kernel leg (rank 0): the same steps once more with a CUDA-event pair around every launch, on ONE stream
    #      (the CRE stack otherwise overlaps the gene stack on a second stream and a kernel's event-to-event time would
    #      include its neighbour): numerator / denominator of the roofline and the kernel breakdown ----
    if rank == 0:
        from variantformer_b200 import engine as engine_mod
        prof = ops.EventProfiler(); ops.PROFILER = prof
        two = engine_mod.CRE_STREAM
        engine_mod.CRE_STREAM = False
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        run_steps(0, args.steps, False, gather=False)   # (rank 0 alone: no collective in this pass)
        p1.record()
        torch.cuda.synchronize()
        prof_ms = p0.elapsed_time(p1)
        engine_mod.CRE_STREAM = two
        ops.PROFILER = None
        peaks = load_peaks()
        summ = prof.summarize()
        g = summ.get("gemm", {"ms": 0.0, "flops": 0.0, "n": 0})
        achieved = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        out["roofline"] = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all epilogues, 1-CTA and CTA-pair)", "achieved": achieved,
                           "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                           "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                           "launches": g["n"], "share_of_step": g["ms"] / prof_ms if prof_ms else None,
                           "measured_in": "single-stream instrumented pass after the timed region "
                                          f"({prof_ms / args.steps:.1f} ms/step vs {elapsed_ms / args.steps:.1f} timed)"}
        out["kernel_breakdown_ms_per_step"] = {k: v["ms"] / args.steps for k, v in summ.items()}
        det = prof.summarize_detail()
        out["kernel_detail"] = {k: {"ms_per_step": round(v["ms"] / args.steps, 3), "launches_per_step": v["n"] / args.steps,
                                    "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["ms"] > 0 else 0.0}
                                for k, v in sorted(det.items(), key=lambda kv: -kv[1]["ms"])[:24]}
        total_flops = sum(v["flops"] for v in summ.values())
        out["config"]["algorithmic_tflop_per_step"] = total_flops / args.steps / 1e12
        out["achieved_tflops_all_kernels"] = total_flops / (elapsed_ms / 1e3) / 1e12
        if world == 1 and not args.no_cpu_baseline:
            sd_cpu = {k: v.float().cpu() for k, v in sd.items()}
            out["cpu_baseline"] = cpu_baseline(sd_cpu, cfg, hp, chroms, var, sets[0][0], C, T)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
