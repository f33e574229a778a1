"""Summarise an ncu CSV log of `bench.py` (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum)
into profiles/r02_gemm_traffic.json (DRAM bytes per GEMM launch: bench.py's roofline.traffic) and a per-kernel launch
list summary (share of the step per kernel family).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
    python tools/ncu_traffic.py gpurun_out/r02_launches.csv profiles/r02_gemm_traffic.json profiles/r02_launches_summary.csv
"""
import collections
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def main():
    src, out_json, out_csv = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    per = collections.defaultdict(lambda: {"n": 0, "ms": 0.0, "rd": 0.0, "wr": 0.0})
    ids = collections.defaultdict(dict)
    for r in rd:
        name, metric = r["Kernel Name"], r["Metric Name"]
        val = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        ids[(r["ID"], name)][metric] = val
    for (_, name), m in ids.items():
        fam = ("gemm_tcgen05_kernel" if "gemm_tcgen05" in name else "attention_mc_kernel" if "attention_mc" in name else
               "bpe_tokenize(_cluster)_kernel" if "bpe_tokenize" in name else name.split("(")[0].split("<")[0].replace("void ", "").replace("vf::", ""))
        p = per[fam]
        p["n"] += 1; p["ms"] += m.get("gpu__time_duration.sum", 0.0)
        p["rd"] += m.get("dram__bytes_read.sum", 0.0); p["wr"] += m.get("dram__bytes_write.sum", 0.0)
    tot = sum(p["ms"] for p in per.values())
    with open(out_csv, "w") as f:
        f.write("kernel,launches,total_ms,share_of_kernel_time,dram_read_GB,dram_write_GB\n")
        for fam, p in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
            f.write(f"{fam},{p['n']},{p['ms']:.3f},{p['ms'] / tot:.4f},{p['rd'] / 1e9:.3f},{p['wr'] / 1e9:.3f}\n")
    g = per["gemm_tcgen05_kernel"]
    cfg = dict(a.split("=") for a in sys.argv[4:])
    json.dump({"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over `python bench.py "
                         f"--steps 2 --warmup 3` ({src}); all gemm_tcgen05_kernel launches of the run averaged",
               "gemm_launches": g["n"], "gemm_dram_bytes_per_launch": (g["rd"] + g["wr"]) / max(g["n"], 1),
               "gemm_dram_read_bytes_per_launch": g["rd"] / max(g["n"], 1),
               "gemm_dram_write_bytes_per_launch": g["wr"] / max(g["n"], 1),
               "gemm_share_of_kernel_time_under_ncu": g["ms"] / tot,
               "config_id": int(cfg.get("config", 3)), "genes_per_step": int(cfg.get("genes", 8)), "cre": int(cfg.get("cre", 1024)),
               "tissues": int(cfg.get("tissues", 63))}, open(out_json, "w"), indent=1)
    print(open(out_csv).read())


if __name__ == "__main__":
    main()
