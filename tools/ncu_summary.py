"""Markdown summary of an `ncu --set full --import-source on` report: key raw metrics per kernel + the instructions
with the most stall samples.   python tools/ncu_summary.py report.ncu-rep out.md ["note"]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none: `{rep}`", "", note, ""]
    for r in rows[2:]:
        lines.append(f"## {r[hdr.index('Kernel Name')][:110]}")
        lines.append("```")
        for k in KEYS:
            if k in hdr:
                lines.append(f"{k} [{units[hdr.index(k)]}] = {r[hdr.index(k)]}")
        stall = {h: float(r[i].replace(',', '')) for i, h in enumerate(hdr)
                 if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]}
        top = sorted(stall.items(), key=lambda kv: -kv[1])[:8]
        lines.append("stalled warps per issue (top): " + ", ".join(
            f"{k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for k, v in top))
        lines.append("```")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    starts = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    for n, st in enumerate(starts):
        h = srows[st]
        end = starts[n + 1] - 1 if n + 1 < len(starts) else len(srows)
        data = [r for r in srows[st + 1:end] if len(r) == len(h)]
        ix = {x: i for i, x in enumerate(h)}
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
        inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
        lines += [f"### kernel {n}: {tot} PC samples, {inst} warp instructions executed; hottest instructions", "```"]
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
            lines.append(f"{int(r[ix['# Samples']]) / tot * 100:5.2f}%  x{int(r[ix['Instructions Executed']]):>9}  "
                         f"{r[ix['Source']].strip()[:90]}")
        lines.append("```")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
