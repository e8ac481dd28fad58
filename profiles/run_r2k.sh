#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_policy.py tests/test_gpu_guardbands.py tests/test_gpu_shaping.py -q --timeout 900 -x > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -n 15 gpurun_out/r2k_pytest.log | cut -c1-300
: > gpurun_out/r2k_k5.jsonl
for i in 1 2; do
python profiles/k5_bench.py --reps 30 --values >> gpurun_out/r2k_k5.jsonl 2>> gpurun_out/r2k_k5.err
DRONE_B200_LIB=$PWD/build_variants/libdd_k5_ss.so python profiles/k5_bench.py --reps 30 --values >> gpurun_out/r2k_k5.jsonl 2>> gpurun_out/r2k_k5.err
done
python profiles/k5_bench.py --reps 30 --envs 75776 >> gpurun_out/r2k_k5.jsonl 2>> gpurun_out/r2k_k5.err
python - <<'PY'
import json
for l in open('gpurun_out/r2k_k5.jsonl'):
    d=json.loads(l); print(d['lib'].split('/')[-1].ljust(20), d['envs'], round(d['ms_per_launch'],4), round(d['mlp_tflops'],1), d.get('values_ms'))
PY
