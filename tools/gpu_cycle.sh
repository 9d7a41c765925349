#!/bin/bash
# one GPU cycle: parity tests, a short bench, the ncu launch list and one full capture of a steady step-kernel launch
# usage (under gpurun): bash tools/gpu_cycle.sh TAG [full]
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
python bench.py --steps 20 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print(d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, (d.get("large_batch") or {}).get("value",0)/1e6, d["env_errors"])
PY
if [ "$2" = "full" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --large-envs 0 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fjsp_step_kernel --launch-skip 68 --launch-count 1 -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --large-envs 0 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
fi
