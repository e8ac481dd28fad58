#!/bin/bash
# K5 ablation timing: which part of the serial chain carries the time? (variants give WRONG results, timing only)
mkdir -p gpurun_out; : > gpurun_out/r2j_k5.jsonl
python profiles/k5_bench.py --reps 20 >> gpurun_out/r2j_k5.jsonl 2>> gpurun_out/r2j_k5.err
for f in build_variants/libdd_k5_*.so; do DRONE_B200_LIB=$PWD/$f python profiles/k5_bench.py --reps 20 >> gpurun_out/r2j_k5.jsonl 2>> gpurun_out/r2j_k5.err; done
python profiles/k5_bench.py --reps 20 >> gpurun_out/r2j_k5.jsonl 2>> gpurun_out/r2j_k5.err
python - <<'PY'
import json
for l in open('gpurun_out/r2j_k5.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1].ljust(28), round(d['ms_per_launch'],4))
PY
