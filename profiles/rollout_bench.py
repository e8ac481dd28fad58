import importlib, sys, json, torch
sys.path.insert(0, "/root/repo")
dd = importlib.import_module("reinforcement-learning-101_b200")
n, T = 1 << 20, 50
res = {}
for pol in ("random", "bangbang"):
    e = dd.BatchedDroneEnv(n, device="cuda:0", seed=0, randomize_drone=True, randomize_platform=True, max_steps=250, auto_reset=True)
    e.reset(); e.rollout(T, pol); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for j in range(20): e.rollout(T, pol, t0=(j + 1) * T)
    b.record(); torch.cuda.synchronize()
    res[pol] = n * T * 20 / (a.elapsed_time(b) * 1e-3)
print(json.dumps(res))
