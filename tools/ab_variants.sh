#!/bin/bash
# Same-box A/B of library builds (tools/build_variant.py) on the five attention shapes: bash tools/ab_variants.sh base NAME...
# Prints vf_attention_mc_varlen ms per shape (97-tok windows, 200-tok chunks, CRE self, gene self, gene->CRE cross).
for v in "$@"; do
  if [ "$v" = base ]; then unset VF_LIB; else export VF_LIB=build/libvf_$v.so; fi
  python tools/bench_flash_attn.py gpurun_out/fa_$v.json > gpurun_out/fa_$v.log 2>&1
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/fa_$v.log") if l.startswith("{")]
print("$v", " ".join(f"{x['mc_ms']:.3f}" for x in d), " diff", " ".join(f"{x['max_abs_diff_vs_flash_attn']:.4f}" for x in d))
PY
done
