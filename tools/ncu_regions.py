"""Attribute the PC samples of an ncu report (--set full --import-source on) to the warp roles of attention_mc_kernel.

    python tools/ncu_regions.py report.ncu-rep

The role regions are found in the SASS itself: USETMAXREG.DEALLOC opens the producer / MMA-issuer warpgroup, the first
UTCHMMA after it the issuer code, USETMAXREG.ALLOC the softmax warpgroups.  Per region: share of samples, instructions
executed, samples per executed warp-instruction (= average stall cycles per instruction, in sampling units) and the
top stall reasons.
"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    data = rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    src = [r[ix["Source"]].strip() for r in data]
    dealloc = next(i for i, s in enumerate(src) if "USETMAXREG.DEALLOC" in s)
    alloc = next(i for i, s in enumerate(src) if "USETMAXREG" in s and "ALLOC" in s and "DEALLOC" not in s)
    first_mma = next(i for i, s in enumerate(src) if "UTCHMMA" in s and i > dealloc)
    # the issuer code starts at the branch target before the first UTCHMMA; approximate by the last UTMALDG + 1
    last_tma = max(i for i, s in enumerate(src) if "UTMALDG" in s)
    regions = [("prologue", 0, dealloc), ("producer", dealloc, last_tma + 40), ("mma issuer", last_tma + 40, alloc),
               ("softmax+epilogue", alloc, len(src))]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "(Not" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"total samples {tot}; dealloc@{dealloc} last_tma@{last_tma} first_mma@{first_mma} alloc@{alloc} n={len(src)}")
    for name, a, b in regions:
        smp = sum(int(r[ix["# Samples"]]) for r in data[a:b])
        ex = sum(int(r[ix["Instructions Executed"]]) for r in data[a:b])
        st = collections.Counter()
        for r in data[a:b]:
            for k in stall_cols:
                if r[ix[k]].isdigit():
                    st[k.replace("stall_", "")] += int(r[ix[k]])
        top = ", ".join(f"{k} {100 * v / max(smp, 1):.0f}%" for k, v in st.most_common(6))
        print(f"{name:18s} samples {100 * smp / tot:5.1f}%  executed {ex / 1e6:8.2f} M  samples/kinst {1e3 * smp / max(ex, 1):7.2f}  [{top}]")
    if len(sys.argv) > 2:
        a, b = regions[int(sys.argv[2])][1:]
        top = sorted(range(a, b), key=lambda i: -int(data[i][ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]
        for i in sorted(top):
            r = data[i]
            st = {k.replace("stall_", ""): int(r[ix[k]]) for k in stall_cols if r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
            print(f"{i:5d} {100 * int(r[ix['# Samples']]) / tot:5.2f}% ex={int(r[ix['Instructions Executed']]):8d} {src[i][:60]:60s} "
                  f"{sorted(st.items(), key=lambda kv: -kv[1])[:3]}")


if __name__ == "__main__":
    main()
