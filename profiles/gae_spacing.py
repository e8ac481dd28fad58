import importlib, json, os, sys, torch
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/bench.py") else os.getcwd())
dd = importlib.import_module("reinforcement-learning-101_b200")
dev = torch.device("cuda:0")
T, n = 250, 65536
lib = dd.native.lib(); st = torch.cuda.current_stream().cuda_stream
arena = torch.zeros(1 << 30, dtype=torch.uint8, device=dev)
base = (arena.data_ptr() + (1 << 26) - 1) // (1 << 26) * (1 << 26) - arena.data_ptr()     # 64 MiB aligned start
def timed(f, reps=300):
    f(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(reps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
res = {}
for name, skew in (("64MiB exact", 0), ("+256B", 256), ("+4KiB", 4096), ("+36KiB", 36864), ("+260KiB (1 row + 4 KiB)", 266240), ("+1MiB", 1 << 20), ("+1.3MiB", 1363148 // 256 * 256)):
    sp = (1 << 26) + skew
    offs = [base + k * sp for k in range(5)]
    r = arena[offs[0]:offs[0] + T * n * 4].view(torch.float32).view(T, n); r.normal_()
    v = arena[offs[1]:offs[1] + (T + 1) * n * 4].view(torch.float32).view(T + 1, n); v.normal_()
    a = arena[offs[2]:offs[2] + T * n * 4].view(torch.float32).view(T, n)
    d = arena[offs[3]:offs[3] + T * n]; d.copy_((torch.rand(T * n, device=dev) < 0.01).to(torch.uint8))
    m = torch.zeros(3, dtype=torch.float64, device=dev)
    def f(k):
        for _ in range(k):
            lib.dd_gae_moments(r.data_ptr(), v.data_ptr(), d.data_ptr(), a.data_ptr(), None, m.data_ptr(), 0.99, 0.95, T, n, st)
    res[name] = round(timed(f), 2)
print(json.dumps(res))
