#!/bin/bash
# job S: validation of the final tree on one GPU: tests, smoke, reference arm, bench (driver arguments and defaults), K5 stand-alone
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -n 4 gpurun_out/r2s_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r2s_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_k20.json 2> gpurun_out/r2s_bench_k20.err; echo "bench rc=$?"
python profiles/k5_bench.py --reps 20 --values > gpurun_out/r2s_k5_plain.json 2>&1
python - <<'P'
import json
d = json.load(open("gpurun_out/r2s_bench_k20.json"))
print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["roofline"]["frac"], d["e2e"]["value"])
v = d["variants"]["fused_policy_rollout"]
print("k5", v["ms_per_launch"], v["frac_of_sustained_bf16"], v["all_148_sms"]["ms_per_launch"])
print(open("gpurun_out/r2s_k5_plain.json").read()[:400])
P
