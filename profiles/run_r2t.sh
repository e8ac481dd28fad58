#!/bin/bash
# job T: ncu evidence for the final round-2 tree (each capture after the same command has exited 0 without ncu)
mkdir -p gpurun_out
BENCH="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-socket --min-time-ms 5 --e2e-steps 3"
$BENCH > gpurun_out/r2t_bench_plain.json 2> gpurun_out/r2t_bench_plain.err; echo "plain bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2t_launches.csv $BENCH > gpurun_out/r2t_ncu_bench.log 2>&1; echo "ncu launch list rc=$?"
python profiles/k5_bench.py --reps 2 > gpurun_out/r2t_k5_plain.json 2>&1; echo "k5 plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:policy_rollout -c 1 -o gpurun_out/r2t_k5 -f python profiles/k5_bench.py --reps 1 > gpurun_out/r2t_ncu_k5.log 2>&1; echo "ncu k5 rc=$?"
K5_ENVS=32768 python profiles/k5_bench.py --reps 2 > gpurun_out/r2t_k5_32k_plain.json 2>&1; echo "k5 32k plain rc=$?"
K5_ENVS=32768 ncu --set full --clock-control none --import-source on -k regex:policy_rollout -c 1 -o gpurun_out/r2t_k5_32k -f python profiles/k5_bench.py --reps 1 > gpurun_out/r2t_ncu_k5_32k.log 2>&1; echo "ncu k5 32k rc=$?"
ls -la gpurun_out | tail -8
