#!/bin/bash
# single-GPU job H: full tests (incl. guard bands), K5 wait-variant A/B, parity report, final ncu captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -n 5 gpurun_out/r2h_pytest.log
: > gpurun_out/r2h_k5.jsonl
for i in 1 2; do
python profiles/k5_bench.py --reps 30 >> gpurun_out/r2h_k5.jsonl 2>> gpurun_out/r2h_k5.err
for v in hint allwait allwait_hint; do DRONE_B200_LIB=$PWD/build_variants/libdd_k5_$v.so python profiles/k5_bench.py --reps 30 >> gpurun_out/r2h_k5.jsonl 2>> gpurun_out/r2h_k5.err; done
done
cut -c1-250 gpurun_out/r2h_k5.jsonl
python tests/parity_report.py > gpurun_out/r2h_parity_report.json 2> gpurun_out/r2h_parity_report.err; echo "parity report rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench_k20.json 2> gpurun_out/r2h_bench_k20.err; echo "bench rc=$?"
