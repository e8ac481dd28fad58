#!/bin/bash
# job Y: steady-state DRAM traffic of K1 on the final round-2 tree (no cache flush between the profiled launches; they rotate
# over 6 shards of ~116 MiB, so every launch misses the 126 MB L2 anyway)
mkdir -p gpurun_out
BENCH="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-socket --no-policy --min-time-ms 5 --e2e-steps 3"
$BENCH > gpurun_out/r2y_bench_plain.json 2> gpurun_out/r2y_bench_plain.err; echo "plain bench rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:step_kernel -s 300 -c 192 --csv --log-file gpurun_out/r2y_traffic.csv $BENCH > gpurun_out/r2y_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/r2y_traffic.csv | cut -c1-300
