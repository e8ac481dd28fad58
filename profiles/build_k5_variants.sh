#!/bin/bash
# Build A/B variants of libdrone_b200.so for K5 experiments into build_variants/ (git-ignored; travels to the GPU box).
#   profiles/build_k5_variants.sh name1:"-DDD_K5_X=1 -DDD_K5_Y=2" name2:"..."
# The other two translation units are compiled once and linked into every variant.
set -e
cd "$(dirname "$0")/.."
SRC=reinforcement-learning-101_b200/csrc
OUT=build_variants
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -ffp-contract=off -Iinclude"
for f in drone_kernels ppo_kernels; do
  if [ ! -f $OUT/$f.o ] || [ $SRC/$f.cu -nt $OUT/$f.o ] || [ $SRC/drone_core.cuh -nt $OUT/$f.o ] || [ $SRC/drone_device.cuh -nt $OUT/$f.o ] || [ include/drone_b200.h -nt $OUT/$f.o ]; then
    nvcc $FLAGS -c $SRC/$f.cu -o $OUT/$f.o &
  fi
done
wait
pids=()
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  ( nvcc $FLAGS $defs -Xptxas -v -c $SRC/policy_rollout.cu -o $OUT/pr_$name.o 2> $OUT/pr_$name.ptxas.txt &&
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libdd_$name.so $OUT/drone_kernels.o $OUT/ppo_kernels.o $OUT/pr_$name.o &&
    echo "built $OUT/libdd_$name.so [$defs]" ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
exit $rc
