#!/bin/bash
# every non-default BASELINE config through bench.py (short runs): bash tools/run_configs.sh TAG [steps]
TAG=$1; STEPS=${2:-5}
mkdir -p gpurun_out
for c in so_single breakdown_65536 brandimarte_1m large_m20; do
  timeout 420 python bench.py --config $c --steps $STEPS --warmup 3 --cpu-seconds 4 > gpurun_out/cfg_${TAG}_$c.json 2> gpurun_out/cfg_${TAG}_$c.err || tail -5 gpurun_out/cfg_${TAG}_$c.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/cfg_${TAG}_$c.json"))
    print("$c", "value %.3f M e2e %.3f M" % (d["value"]/1e6, d["e2e"]["value"]/1e6), "parity", d["parity_sample"]["ok"], "errors", d["env_errors"], "cpu %.3f M" % (d["cpu_baseline"]["value"]/1e6), "policy", [(p.get("envs_per_gpu"), round(p.get("cuda_graph",{}).get("value",0)/1e6,2), round(p.get("eager_launches",{}).get("value",0)/1e6,2), p.get("error")) for p in d["policy_in_loop"]])
except Exception as e:
    print("$c failed", e)
PY
done
