#!/bin/bash
# final single-GPU validation of the round-2 tree: tests, smoke, bench (driver arguments), parity report, K5 ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -n 4 gpurun_out/r2o_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2o_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2o_ref.json 2> gpurun_out/r2o_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_k20.json 2> gpurun_out/r2o_bench_k20.err; echo "bench rc=$?"
python bench.py > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err; echo "bench default rc=$?"
python tests/parity_report.py > gpurun_out/r2o_parity_report.json 2> gpurun_out/r2o_parity_report.err; echo "parity rc=$?"
python profiles/k5_bench.py --reps 2 > gpurun_out/r2o_k5_plain.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_rollout -c 1 -o gpurun_out/r2o_k5 -f python profiles/k5_bench.py --reps 1 > gpurun_out/r2o_ncu_k5.log 2>&1; echo "ncu k5 rc=$?"
