#!/bin/bash
# job AC: ncu --set full of K1 (both variants) and of the T-step rollout kernel on the final round-2 tree
mkdir -p gpurun_out
BENCH="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-socket --no-policy --min-time-ms 5 --e2e-steps 3"
$BENCH > /dev/null 2> gpurun_out/r2ac_plain.err; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 400 -c 1 -o gpurun_out/r2ac_step_obs -f $BENCH > gpurun_out/r2ac_ncu1.log 2>&1; echo "ncu step rc=$?"
python profiles/rollout_bench.py > gpurun_out/r2ac_rollout_plain.json 2>&1; echo "rollout plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -o gpurun_out/r2ac_rollout -f python profiles/rollout_bench.py > gpurun_out/r2ac_ncu2.log 2>&1; echo "ncu rollout rc=$?"
