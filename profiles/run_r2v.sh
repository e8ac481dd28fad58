#!/bin/bash
# job V: final single-GPU validation + evidence: tests, smoke, reference arm, bench (driver arguments), parity report, K5 plain + ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -n 4 gpurun_out/r2v_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2v_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2v_ref.json 2> gpurun_out/r2v_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2v_bench_k20.json 2> gpurun_out/r2v_bench_k20.err; echo "bench rc=$?"
python tests/parity_report.py > gpurun_out/r2v_parity_report.json 2> gpurun_out/r2v_parity_report.err; echo "parity rc=$?"
python profiles/k5_bench.py --reps 20 --values --checksum > gpurun_out/r2v_k5_plain.json 2>&1; echo "k5 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:policy_rollout -c 1 -o gpurun_out/r2v_k5 -f python profiles/k5_bench.py --reps 1 > gpurun_out/r2v_ncu_k5.log 2>&1; echo "ncu k5 rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/r2v_bench_k20.json"))
print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["roofline"]["frac"], d["e2e"]["value"])
v = d["variants"]["fused_policy_rollout"]
print("k5 in bench", v["ms_per_launch"], v["frac_of_sustained_bf16"], v["all_148_sms"]["ms_per_launch"], v.get("ppo_data_path", {}).get("ms"), v.get("ppo_iteration_e2e", {}).get("ms"))
print(open("gpurun_out/r2v_k5_plain.json").read()[:260])
P
