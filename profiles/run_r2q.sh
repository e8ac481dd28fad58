#!/bin/bash
# job Q: A/B timing of K5 builds x launch knobs.  Each argument: <variant>[,ENV=VALUE]...   (build_variants/libdd_<variant>.so)
mkdir -p gpurun_out
OUT=gpurun_out/r2q_k5.jsonl
: > $OUT
EXTRA=${K5_BENCH_ARGS:---reps 20 --checksum}
for spec in "$@"; do
  IFS=',' read -ra parts <<< "$spec"
  v=${parts[0]}
  envs=("${parts[@]:1}")
  env DRONE_B200_LIB=$PWD/build_variants/libdd_$v.so "${envs[@]}" timeout 180 python profiles/k5_bench.py $EXTRA > gpurun_out/r2q_one.json 2>> gpurun_out/r2q_err.log
  rc=$?
  if [ $rc -eq 0 ]; then python - "$spec" <<'P' >> $OUT
import json, sys
d = json.load(open("gpurun_out/r2q_one.json")); d["spec"] = sys.argv[1]; print(json.dumps(d))
P
  else echo "{\"spec\": \"$spec\", \"failed\": $rc}" >> $OUT; tail -n 5 gpurun_out/r2q_err.log; fi
done
python - <<'P'
import json
for l in open("gpurun_out/r2q_k5.jsonl"):
    d = json.loads(l)
    print(d["spec"], d.get("ms_per_launch"), d.get("checksum", d.get("failed")))
P
