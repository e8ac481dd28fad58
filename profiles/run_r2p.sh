#!/bin/bash
# job P: A/B timing of K5 build variants (build_variants/libdd_<name>.so, made by profiles/build_k5_variants.sh)
mkdir -p gpurun_out
OUT=gpurun_out/r2p_k5_variants.jsonl
: > $OUT
for v in base "$@" base; do
  DRONE_B200_LIB=$PWD/build_variants/libdd_$v.so timeout 300 python profiles/k5_bench.py --reps 20 --checksum >> $OUT 2>> gpurun_out/r2p_err.log || echo "{\"lib\": \"$v\", \"failed\": $?}" >> $OUT
done
python - <<'P'
import json
for l in open("gpurun_out/r2p_k5_variants.jsonl"):
    d = json.loads(l)
    print(d.get("lib", "").split("libdd_")[-1], d.get("ms_per_launch"), d.get("checksum", d))
P
