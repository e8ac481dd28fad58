#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/r2p_k5.jsonl
for i in 1 2; do
python profiles/k5_bench.py --reps 30 >> gpurun_out/r2p_k5.jsonl 2>> gpurun_out/r2p_k5.err
for f in build_variants/libdd_k5_*.so; do DRONE_B200_LIB=$PWD/$f python profiles/k5_bench.py --reps 30 >> gpurun_out/r2p_k5.jsonl 2>> gpurun_out/r2p_k5.err; done
done
python - <<'PY'
import json
for l in open('gpurun_out/r2p_k5.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1].ljust(28), d['envs'], round(d['ms_per_launch'],4))
PY
