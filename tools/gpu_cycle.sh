#!/bin/bash
# one GPU cycle: parity tests, the default bench line, the per-CTA trace; "full" adds the reference arm, the
# other four BASELINE configs, the ncu launch list and full captures of a steady step-kernel launch at 4096 and
# 65 536 copies (each ncu run only after the same command has exited 0 without ncu)
# usage (under gpurun): bash tools/gpu_cycle.sh TAG [full]
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tests_$TAG.log
cat gpurun_out/tests_$TAG.log
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-seconds 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value M", d["value"]/1e6, "ms", d["ms_per_step"], "e2e M", d["e2e"]["value"]/1e6, "large M", (d.get("large_batch") or {}).get("value",0)/1e6, "errors", d["env_errors"], "parity", d["parity_sample"]["ok"])
print("sweep", [(s["env_steps_per_launch"], round(s["value"]/1e6,2)) for s in d.get("rollout_sweep", [])], "policy", [(p.get("envs_per_gpu"), round(p.get("cuda_graph",{}).get("value",0)/1e6,2)) for p in d.get("policy_in_loop", [])])
PY
if [ -f deep_reinforcement_learning_for_fjsp_b200/libfjsp_b200_trace.so ]; then
FJSP_B200_LIB=deep_reinforcement_learning_for_fjsp_b200/libfjsp_b200_trace.so timeout 300 python tools/cta_trace.py --launches 4 > gpurun_out/trace_$TAG.log 2>&1
grep -v "^CTA \|slot map" gpurun_out/trace_$TAG.log | tail -12
fi
if [ "$2" = "full" ]; then
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
bash tools/run_configs.sh $TAG 5
ARGS="--steps 3 --warmup 3 --large-envs 0 --no-cpu-baseline --no-sweep --no-policy --parity-envs 0"
timeout 300 python bench.py $ARGS > /dev/null 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $ARGS > gpurun_out/ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fjsp_step_kernel --launch-skip 68 --launch-count 1 -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 300 python bench.py --envs 65536 $ARGS > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:fjsp_step_kernel --launch-skip 68 --launch-count 1 -f -o gpurun_out/prof_${TAG}_65536 python bench.py --envs 65536 $ARGS > gpurun_out/ncu_full_${TAG}_65536.log 2>&1
fi
