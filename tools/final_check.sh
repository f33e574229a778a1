#!/bin/bash
# Round-end self check on a GPU box: the driver's three steps (GPU tests, smoke, both bench arms) in one call.
#   bash tools/final_check.sh [tag]        (add NCU=1 for the launch list + DRAM traffic pass, ~17 min)
tag=${1:-final}
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
tail -c 300 gpurun_out/${tag}_bench_default.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench_default.json")); r = json.load(open("gpurun_out/${tag}_bench_reference.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "gemm frac", d["roofline"]["frac"],
      "parity", d["parity"]["ok"], d["parity"]["rel_err_max_norm"], "cpu", d["cpu_baseline"]["value"], "ref arm", r["value"],
      "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"])
PY
if [ -n "$NCU" ]; then
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
      > gpurun_out/${tag}_under_ncu.json 2> gpurun_out/${tag}_under_ncu.err
fi
