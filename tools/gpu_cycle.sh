#!/bin/bash
# one GPU cycle: parity tests, a short bench, the per-CTA trace; "full" adds the ncu launch list and one
# full capture of a steady step-kernel launch
# usage (under gpurun): bash tools/gpu_cycle.sh TAG [full]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q -k "random_batch or single_steps" 2>&1 | tail -3; timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
timeout 240 python bench.py --steps 20 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value M", d["value"]/1e6, "ms", d["ms_per_step"], "e2e M", d["e2e"]["value"]/1e6, "large M", (d.get("large_batch") or {}).get("value",0)/1e6, "errors", d["env_errors"])
print("sweep", [(s["env_steps_per_launch"], round(s["value"]/1e6,2)) for s in d.get("rollout_sweep", [])])
PY
if [ -f deep_reinforcement_learning_for_fjsp_b200/libfjsp_b200_trace.so ]; then
FJSP_B200_LIB=deep_reinforcement_learning_for_fjsp_b200/libfjsp_b200_trace.so timeout 300 python tools/cta_trace.py --launches 4 > gpurun_out/trace_$TAG.log 2>&1
tail -22 gpurun_out/trace_$TAG.log
fi
if [ "$2" = "full" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --large-envs 0 --no-cpu-baseline --no-sweep > gpurun_out/ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fjsp_step_kernel --launch-skip 68 --launch-count 1 -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --large-envs 0 --no-cpu-baseline --no-sweep > gpurun_out/ncu_full_$TAG.log 2>&1
fi
