#!/bin/bash
# job I: ncu evidence for round 2 (each capture after the same command has exited 0 without ncu)
mkdir -p gpurun_out
BENCH="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-socket --min-time-ms 5 --e2e-steps 3"
$BENCH > gpurun_out/r2i_bench_plain.json 2> gpurun_out/r2i_bench_plain.err; echo "plain bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2i_launches.csv $BENCH > gpurun_out/r2i_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python profiles/k5_bench.py --reps 2 > gpurun_out/r2i_k5_plain.json 2>&1; echo "k5 plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:policy_rollout -c 1 -o gpurun_out/r2i_k5 -f python profiles/k5_bench.py --reps 1 > gpurun_out/r2i_ncu_k5.log 2>&1; echo "ncu k5 rc=$?"
python profiles/gae_bench.py --reps 3 > gpurun_out/r2i_gae_plain.json 2>&1; echo "gae plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gae_kernel|moments_kernel|normalize_kernel|returns_kernel" -s 20 -c 8 -o gpurun_out/r2i_tail -f python profiles/gae_bench.py --reps 3 > gpurun_out/r2i_ncu_tail.log 2>&1; echo "ncu tail rc=$?"
ls -la gpurun_out | tail -12
