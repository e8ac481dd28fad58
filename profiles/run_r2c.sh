#!/bin/bash
# round-2 GPU job C: tests, bench at several K (graph piece length), GAE ring variants, compute-sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench_k20.json 2> gpurun_out/r2c_bench_k20.err; echo "bench rc=$?"
for K in 96 1000; do python bench.py --steps $K --warmup 5 --no-cpu-baseline --no-socket --no-policy --no-curriculum --no-f64 > gpurun_out/r2c_bench_k$K.json 2> gpurun_out/r2c_bench_k$K.err; done
python bench.py --steps 20 --warmup 5 --chains 3 --no-cpu-baseline --no-socket --no-policy --no-curriculum --no-f64 > gpurun_out/r2c_bench_k20_c3.json 2> gpurun_out/r2c_bench_k20_c3.err
python bench.py --steps 20 --warmup 5 --chains 1 --no-cpu-baseline --no-socket --no-policy --no-curriculum --no-f64 > gpurun_out/r2c_bench_k20_c1.json 2> gpurun_out/r2c_bench_k20_c1.err
python profiles/gae_bench.py --label default > gpurun_out/r2c_gae.jsonl 2> gpurun_out/r2c_gae.err
for f in build_variants/libdd_ring_*.so; do DRONE_B200_LIB=$PWD/$f python profiles/gae_bench.py --label $(basename $f) >> gpurun_out/r2c_gae.jsonl 2>> gpurun_out/r2c_gae.err; done
cut -c1-330 gpurun_out/r2c_gae.jsonl
timeout 600 compute-sanitizer --tool memcheck python profiles/sanitize.py > gpurun_out/r2c_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck python profiles/sanitize.py > gpurun_out/r2c_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r2c_sanitizer_memcheck.log gpurun_out/r2c_sanitizer_racecheck.log
