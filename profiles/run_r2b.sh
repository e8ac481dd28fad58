#!/bin/bash
# round-2 GPU job B: parity tests, bench at the driver's arguments, K5 A/B (fp16 / bf16 operands), GAE variants, ncu of K5 + tail kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_k20.json 2> gpurun_out/r2b_bench_k20.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --eager --no-cpu-baseline --no-socket --no-policy --no-curriculum --no-f64 > gpurun_out/r2b_bench_k20_eager.json 2> gpurun_out/r2b_bench_k20_eager.err; echo "bench eager rc=$?"
for op in fp16 bf16; do python profiles/k5_bench.py --reps 20 --operands $op --values >> gpurun_out/r2b_k5.jsonl 2>> gpurun_out/r2b_k5.err; done
python profiles/k5_bench.py --reps 20 --envs 75776 >> gpurun_out/r2b_k5.jsonl 2>> gpurun_out/r2b_k5.err
cat gpurun_out/r2b_k5.jsonl | cut -c1-600
python profiles/gae_bench.py --label default > gpurun_out/r2b_gae.jsonl 2> gpurun_out/r2b_gae.err
for f in build_variants/libdd_gae_*.so; do DRONE_B200_LIB=$PWD/$f python profiles/gae_bench.py --label $(basename $f) >> gpurun_out/r2b_gae.jsonl 2>> gpurun_out/r2b_gae.err; done
ncu --set full --clock-control none --import-source on -k regex:policy_rollout -c 1 -o gpurun_out/r2b_k5 -f python profiles/k5_bench.py --reps 1 > gpurun_out/r2b_ncu_k5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"gae_kernel|moments_kernel|normalize_kernel" -s 12 -c 6 -o gpurun_out/r2b_tail -f python profiles/gae_bench.py --reps 3 > gpurun_out/r2b_ncu_tail.log 2>&1
ls -la gpurun_out/
