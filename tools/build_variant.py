"""Build a variant of libvf_b200.so into build/ (git-ignored) for same-box A/B timing.

    python tools/build_variant.py NAME [-DFLAG[=V] ...] [--src path/to/alternative/vf_attention_mc.cu]

Only vf_attention_mc.cu / vf_gemm.cu are recompiled with the extra flags; the other objects are reused from the in-tree
build.  Load with VF_BENCH_LIB=build/libvf_NAME.so tools/bench_kernels.py.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from variantformer_b200.csrc import build as B  # noqa: E402


def main():
    name = sys.argv[1]
    flags = [a for a in sys.argv[2:] if a.startswith("-D")]
    alt = {}
    args = sys.argv[2:]
    for i, a in enumerate(args):
        if a == "--src":
            alt[os.path.basename(args[i + 1])] = os.path.abspath(args[i + 1])
    B.build()
    out = os.path.join(ROOT, "build")
    os.makedirs(out, exist_ok=True)
    objs = []
    for src in B.SOURCES:
        o = os.path.join(B.HERE, src.replace(".cu", ".o"))
        if src in ("vf_attention_mc.cu", "vf_gemm.cu", "vf_encode.cu") and (flags or src in alt):
            o = os.path.join(out, f"{name}_{src.replace('.cu', '.o')}")
            s = alt.get(src, os.path.join(B.HERE, src))
            r = subprocess.run([B.NVCC] + B.FLAGS + flags + ["-I", B.HERE, "-c", s, "-o", o], capture_output=True, text=True)
            if r.returncode != 0:
                print(r.stdout + r.stderr)
                raise SystemExit(1)
            for line in (r.stdout + r.stderr).splitlines():
                if "spill" in line and "0 bytes spill stores" not in line:
                    print(src, line.strip())
        objs.append(o)
    lib = os.path.join(out, f"libvf_{name}.so")
    subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-lcudart"], check=True)
    print(lib)


if __name__ == "__main__":
    main()
