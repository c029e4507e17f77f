#!/bin/bash
# Same-box A/B of an experiment switch over the whole step: profiles/ab_bench.sh VAR "0 1 0 1"   (needs the experiments build)
# prints value (volumes/s), ms per step, the attention / qkv / fc1 categories and the SM clock for each setting, interleaved.
VAR=$1; shift
export MST_LIB_PATH=$PWD/new-vit_b200/libmst_b200_exp.so
for s in $1; do
  env $VAR=$s timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_${VAR}_$s.json 2>/dev/null
  python - "$VAR" "$s" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}_{sys.argv[2]}.json"))
k = d["kernels"]
print(sys.argv[1], sys.argv[2], "vol/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "attn", round(k["attention"]["ms_per_step"], 3),
      "qkv", round(k["gemm_qkv"]["ms_per_step"], 3), "fc1", round(k["gemm_fc1"]["ms_per_step"], 3), "MHz", d["clocks"]["sm_mhz"])
PY
done
