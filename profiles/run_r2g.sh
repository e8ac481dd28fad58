#!/bin/bash
# 8-GPU job: PCIe ceiling and bench at N = 4 and 8 (one box)
mkdir -p gpurun_out
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N profiles/pcie_ceiling.py 2> gpurun_out/r2g_pcie_n$N.err | grep '^{' > gpurun_out/r2g_pcie_n$N.json
cat gpurun_out/r2g_pcie_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; echo "bench n$N rc=$?"
done
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
lscpu | head -20 > gpurun_out/r2g_lscpu.txt 2>&1
head -c 400 gpurun_out/r2g_bench_n8.json
