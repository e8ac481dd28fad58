#!/bin/bash
# stride / size experiment for the GAE scan: is the [T][n] layout with n = 65536 (256 KB row stride) the limiter?
mkdir -p gpurun_out; : > gpurun_out/r2d_gae.jsonl
for envs in 65536 65600 69632 131072 262144; do
  python profiles/gae_bench.py --label ring_default --envs $envs --reps 60 >> gpurun_out/r2d_gae.jsonl 2>> gpurun_out/r2d_gae.err
  for f in build_variants/libdd_reg_*.so; do DRONE_B200_LIB=$PWD/$f python profiles/gae_bench.py --label $(basename $f) --envs $envs --reps 60 >> gpurun_out/r2d_gae.jsonl 2>> gpurun_out/r2d_gae.err; done
done
cut -c1-250 gpurun_out/r2d_gae.jsonl
