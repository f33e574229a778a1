"""Print gpurun_out/bench_kernels.json (tools/bench_kernels.py) as a compact table."""
import json
import sys

for d in json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench_kernels.json")):
    if d["kernel"] == "gemm":
        print(f"{d['name']:22s} plain {d['ms']:.3f} ms {d['tflops']:7.1f} TF/s | engine cfg {d['ms_engine_cfg']:.3f} ms "
              f"{d['tflops_engine_cfg']:7.1f} TF/s | cuBLAS {d['cublas_ms']:.3f} ms {d['cublas_tflops']:7.1f} TF/s"
              + (f" | bf16 resid, mirror only {d['ms_resid16_mirror_only']:.3f} ms" if d.get('ms_resid16_mirror_only') else ""))
    elif "tflops" in d:
        print(f"{d['kernel']:14s} {d['name']:22s} {d['ms']:.3f} ms {d['tflops']:7.1f} TF/s")
