import importlib, sys, os, json, numpy as np, torch, time
sys.path.insert(0, "/root/repo")
dd = importlib.import_module("reinforcement-learning-101_b200")
d = np.load("/root/repo/tests/golden/policy_v1.npz"); sd = {k: torch.from_numpy(d[k]) for k in d.files if k.startswith("network")}
c = np.load("/root/repo/tests/golden/critic_v1.npz"); sdc = {k: torch.from_numpy(c[k]) for k in c.files if k.startswith("network")}
blob = dd.PolicyBlob(sd, device="cuda:0"); vblob = dd.ValueBlob(sdc, device="cuda:0")
t0 = time.time()
for n, T, reps in ((65536, 250, 300), (75776, 1000, 20), (1000, 3000, 3), (1 << 20, 20, 5), (129, 5000, 2)):
    e = dd.BatchedDroneEnv(n, device="cuda:0", seed=n, randomize_drone=True, randomize_platform=True, max_steps=250, auto_reset=True)
    e.reset()
    buf = dd.policy_rollout(e, blob, T, sample=True, want="arldo" if n * T < 3e7 else "ar")
    for j in range(reps):
        dd.policy_rollout(e, blob, T, sample=True, t0=(j + 1) * T, want="arldo" if n * T < 3e7 else "ar", out=buf)
    torch.cuda.synchronize()
    s = e.stats()
    assert s["env_steps"] == n * T * (reps + 1), s
    print(n, T, reps, "ok", round(time.time() - t0, 1), s["landing_rate"])
x = torch.randn(50_000_003, 15, device="cuda:0") * 0.3
v = dd.value_forward(vblob, x); p = dd.policy_forward(blob, x)
torch.cuda.synchronize()
assert torch.isfinite(v).all() and torch.isfinite(p).all()
print("forward 50M rows ok", round(time.time() - t0, 1))
