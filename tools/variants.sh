#!/bin/bash
# quick A/B of run-time knobs on the bench workload (under gpurun): bash tools/variants.sh TAG "ENV1=.. ENV2=.." "ENV3=.." ...
# BENCH_ARGS adds bench.py flags to every run
TAG=$1; shift
mkdir -p gpurun_out
i=0
for V in "$@"; do
  i=$((i+1))
  env $V timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-policy --large-envs 0 --parity-envs 0 ${BENCH_ARGS} > gpurun_out/var_${TAG}_$i.json 2> gpurun_out/var_${TAG}_$i.err || tail -3 gpurun_out/var_${TAG}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_${TAG}_$i.json"))
    print("[$V] value %.1f M, e2e %.1f M, sweep %s, errors %d, steps ms %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, [(s["env_steps_per_launch"], round(s["value"]/1e6,2), s.get("ms_min_median_max")) for s in d.get("rollout_sweep", [])], d["env_errors"], d.get("step_ms_all")))
except Exception as e:
    print("[$V] failed", e)
PY
done
