#!/bin/bash
# 2-GPU job: the full GPU test suite (incl. the non-current-device test), the bench and the PCIe ceiling at N = 1 and 2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -n 6 gpurun_out/r2f_pytest.log
python profiles/pcie_ceiling.py > gpurun_out/r2f_pcie_n1.json 2> gpurun_out/r2f_pcie_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/pcie_ceiling.py > gpurun_out/r2f_pcie_n2.json 2> gpurun_out/r2f_pcie_n2.err
cat gpurun_out/r2f_pcie_n1.json gpurun_out/r2f_pcie_n2.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench n1 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_ref_n1.json 2> gpurun_out/r2f_ref_n1.err; echo "ref rc=$?"
tail -c 600 gpurun_out/r2f_bench_n2.err
