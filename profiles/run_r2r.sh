#!/bin/bash
# job R: SM clock / power while K5 runs back to back (is the fused rollout power-limited?)
mkdir -p gpurun_out
LIB=${1:-base}
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_throttle_reasons.active --format=csv -lms 50 > gpurun_out/r2r_smi_$LIB.csv &
SMI=$!
sleep 1
DRONE_B200_LIB=$PWD/build_variants/libdd_$LIB.so python profiles/k5_bench.py --reps 1500 > gpurun_out/r2r_k5_$LIB.json 2>> gpurun_out/r2r_err.log
sleep 0.5
kill $SMI
python - "$LIB" <<'P'
import sys, json
rows = [l.strip().split(", ") for l in open(f"gpurun_out/r2r_smi_{sys.argv[1]}.csv")][1:]
busy = [r for r in rows if float(r[2].split()[0]) > 400]
print("samples", len(rows), "busy", len(busy))
if busy:
    clk = sorted(int(r[0].split()[0]) for r in busy); pw = sorted(float(r[2].split()[0]) for r in busy)
    print("sm MHz median", clk[len(clk)//2], "min", clk[0], "max", clk[-1], "| power W median", pw[len(pw)//2], "max", pw[-1], "limit", busy[0][3], "| reasons", sorted(set(r[5] for r in busy)))
print(open(f"gpurun_out/r2r_k5_{sys.argv[1]}.json").read()[:300])
P
