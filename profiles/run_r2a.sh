#!/bin/bash
# round-2 GPU job A: parity tests, the bench at the driver's arguments, the PCIe ceiling, GAE kernel variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_k20.json 2> gpurun_out/r2a_bench_k20.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 5 --eager --no-cpu-baseline --no-socket --no-policy --no-curriculum --no-f64 > gpurun_out/r2a_bench_k20_eager.json 2> gpurun_out/r2a_bench_k20_eager.err; echo "bench eager rc=$?"
python profiles/pcie_ceiling.py > gpurun_out/r2a_pcie_n1.json 2>&1
python profiles/gae_bench.py --label default > gpurun_out/r2a_gae.jsonl 2> gpurun_out/r2a_gae.err
for f in build_variants/libdd_gae_*.so; do DRONE_B200_LIB=$PWD/$f python profiles/gae_bench.py --label $(basename $f) >> gpurun_out/r2a_gae.jsonl 2>> gpurun_out/r2a_gae.err; done
cat gpurun_out/r2a_gae.jsonl | cut -c1-400
