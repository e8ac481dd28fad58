"""ctypes front-end of ``libdrone_b200_host.so`` -- the HOST instantiation of the kernels' per-environment source
(csrc/drone_core.cuh via csrc/host_twin.cpp, include/drone_b200_host.h).  TEST INFRASTRUCTURE: lives under tests/,
the product package never loads that library.

``HostBatch`` mirrors ``BatchedDroneEnv``'s buffers as numpy arrays in the DDState layout and calls the twins with the
same DDState / DDParams / DDEnvConfig structs the device entry points take."""
import ctypes as C
import importlib

import numpy as np

dd = importlib.import_module("reinforcement-learning-101_b200")
nv = dd.native

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(nv.build_host_twin())
        vp, i32, i64, u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32
        PS, PP, PC = C.POINTER(nv.DDState), C.POINTER(nv.DDParams), C.POINTER(nv.DDEnvConfig)
        L.dd_host_abi_version.restype = C.c_int
        L.dd_reset_host.restype = C.c_int
        L.dd_reset_host.argtypes = [PS, PP, PC, vp, vp, i32, i64]
        L.dd_step_host.restype = C.c_int
        L.dd_step_host.argtypes = [PS, PP, PC, vp, vp, i32, vp, vp, vp, vp, i64]
        L.dd_rollout_host.restype = C.c_int
        L.dd_rollout_host.argtypes = [PS, PP, PC, i32, vp, u32, i32, vp, vp, vp, i32, vp, vp, i64]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


class HostBatch:
    def __init__(self, n, dtype=np.float64, seed=0, randomize_drone=False, randomize_platform=True, max_steps=0,
                 auto_reset=False, env_id_base=0, shaping="ppo", params=None):
        self.n, self.dtype = int(n), np.dtype(dtype)
        R = self.dtype
        self.pos_vel = np.zeros((n, 4), R); self.att_fuel = np.zeros((n, 4), R); self.platform = np.zeros((n, 2), R)
        self.steps = np.zeros(n, np.int32); self.episode = np.zeros(n, np.uint32); self.flags = np.zeros(n, np.uint8)
        self.prev_dist = np.full(n, np.nan, R)
        self.stats = np.zeros(nv.STATS_WORDS, np.uint64)
        self.params = params if params is not None else nv.default_params()
        self.state = nv.DDState(_p(self.pos_vel), _p(self.att_fuel), _p(self.platform), _p(self.steps), _p(self.episode),
                                _p(self.flags), nv.F32 if R == np.float32 else nv.F64, 0, _p(self.prev_dist))
        self.cfg = nv.DDEnvConfig(int(seed), int(env_id_base), int(max_steps or 0), int(bool(auto_reset)),
                                  int(bool(randomize_drone)), int(bool(randomize_platform)), 0,
                                  nv.SHAPING_PG if shaping == "pg" else nv.SHAPING_PPO)

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what} failed with code {rc}")

    def reset(self, mask=None, stride=15):
        obs = np.zeros((self.n, stride), self.dtype)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._check(lib().dd_reset_host(C.byref(self.state), C.byref(self.params), C.byref(self.cfg), _p(m), _p(obs), stride, self.n),
                    "dd_reset_host")
        return obs

    def inject(self, x, y, px, py):
        self.pos_vel[:] = 0; self.att_fuel[:] = 0
        self.pos_vel[:, 0] = x; self.pos_vel[:, 1] = y; self.platform[:, 0] = px; self.platform[:, 1] = py
        self.att_fuel[:, 2] = self.params.max_fuel
        self.steps[:] = 0; self.flags[:] = 0; self.episode += 1
        self.prev_dist[:] = np.nan

    def step(self, actions, want_obs=True, want_final=False, stride=15):
        a = np.ascontiguousarray(actions, np.uint8)
        obs = np.zeros((self.n, stride), self.dtype) if want_obs else None
        rew = np.zeros(self.n, self.dtype); fl = np.zeros(self.n, np.uint8)
        fin = np.zeros((self.n, stride), self.dtype) if want_final else None
        self._check(lib().dd_step_host(C.byref(self.state), C.byref(self.params), C.byref(self.cfg), _p(a), _p(obs), stride,
                                       _p(rew), _p(fl), _p(fin), _p(self.stats), self.n), "dd_step_host")
        return (obs, rew, fl, fin) if want_final else (obs, rew, fl)

    def rollout(self, T, policy, actions=None, t0=0, want=("reward", "done"), stride=15):
        out = {}
        if "reward" in want: out["reward"] = np.zeros((T, self.n), self.dtype)
        if "done" in want: out["done"] = np.zeros((T, self.n), np.uint8)
        if "obs" in want: out["obs"] = np.zeros((T, self.n, stride), self.dtype)
        if "shaped" in want: out["shaped"] = np.zeros((T, self.n), self.dtype)
        a = None if actions is None else np.ascontiguousarray(actions, np.uint8)
        self._check(lib().dd_rollout_host(C.byref(self.state), C.byref(self.params), C.byref(self.cfg), int(policy), _p(a), int(t0),
                                          int(T), _p(out.get("reward")), _p(out.get("done")), _p(out.get("obs")), stride,
                                          _p(out.get("shaped")), _p(self.stats), self.n), "dd_rollout_host")
        return out
