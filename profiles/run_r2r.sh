#!/bin/bash
# 2-GPU validation of the final tree: full GPU tests (incl. non-current-device), bench N=1 and N=2, stand-alone GAE on the same box
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -n 4 gpurun_out/r2r_pytest.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; echo "bench n1 rc=$?"
python profiles/gae_bench.py --label same_box_as_bench > gpurun_out/r2r_gae.json 2> gpurun_out/r2r_gae.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; echo "bench n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2r_ref_n2.json 2> gpurun_out/r2r_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
for f in ("n1","n2"):
    t=open(f'gpurun_out/r2r_bench_{f}.json').read(); print(f, "stdout is one JSON line:", t.startswith('{') and t.strip().count('\n')==0)
    d=json.loads(t); o=d['roofline']['others']
    print("  value %.4g frac %.4f e2e %.4g k5 %.4f gae %.3f host_us %.2f eager %.3f mgc %s" % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['variants']['fused_policy_rollout']['ms_per_launch'], o['gae_kernel']['frac'], d['variants']['gym_step_eager_api']['host_us_per_step_raw_call'], d['variants']['gym_step_eager_api']['frac'], (d.get('multi_gpu_check') or {}).get('status')))
g=json.load(open('gpurun_out/r2r_gae.json')); print("stand-alone gae on this box:", {k:(round(v['ms']*1e3,1), round(v['frac'],3)) for k,v in g.items() if isinstance(v,dict)})
print(open('gpurun_out/r2r_ref_n2.json').read()[:300])
PY
