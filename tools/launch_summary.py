"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`) per kernel: launches, total
time, share.  Usage: python tools/launch_summary.py launches.csv out.csv "header comment"
"""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.Counter(); cnt = collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"<.*", "", r[ik].replace("vf::", "").replace("void ", "")).split("(")[0]
        v = float(r[iv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        tot[name] += v * scale; cnt[name] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# {note}\n# {sum(cnt.values())} launches, {total:.1f} ms under ncu (cold-cache, serialised: compare SHARES "
                "with bench.py kernel_breakdown_ms_per_step)\n# kernel, launches, total_ms, share\n")
        for k, v in tot.most_common():
            f.write(f"{k},{cnt[k]},{v:.3f},{v / total:.4f}\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
