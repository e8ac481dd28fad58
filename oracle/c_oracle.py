"""ctypes front-end of ``drone_oracle.c`` (ORACLE, test-only).

``OracleBatch`` holds n environments as float64 numpy structure-of-arrays and
steps them with the C restatement.  Used by tests/ (checker for the CUDA path),
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_drone.so")

DONE, LANDED, CRASHED, TRUNCATED = 1, 2, 4, 8
CAUSE_MASK, CAUSE_GROUND, CAUSE_FUEL, CAUSE_OOB = 0x30, 0x10, 0x20, 0x30
ACT_MAIN, ACT_LEFT, ACT_RIGHT, ACT_SKIP = 1, 2, 4, 0x80
POL_TRACE, POL_RANDOM, POL_BANGBANG = 0, 1, 2

_F64 = ("x", "y", "vx", "vy", "angle", "angvel", "fuel", "px", "py", "ep_return")


class _State(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_double)) for k in _F64] + [
        ("steps", C.POINTER(C.c_int32)),
        ("episode", C.POINTER(C.c_uint32)),
        ("flags", C.POINTER(C.c_uint8)),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("episodes", C.c_uint64), ("landed", C.c_uint64), ("crashed", C.c_uint64),
        ("truncated", C.c_uint64), ("sum_return", C.c_double), ("sum_length", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    """Compile the C oracle in place (gcc, seconds)."""
    src = os.path.join(_HERE, "drone_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_drone.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u8p, f64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_double)
        L.oracle_reset.argtypes = [C.POINTER(_State), u8p, f64p, C.c_int, C.c_int, C.c_int,
                                   C.c_uint64, C.c_uint64, C.c_int64]
        L.oracle_reset.restype = None
        L.oracle_step.argtypes = [C.POINTER(_State), u8p, f64p, C.c_int, f64p, u8p, f64p, C.POINTER(Stats),
                                  C.c_int32, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int64]
        L.oracle_step.restype = None
        L.oracle_rollout.argtypes = [C.POINTER(_State), C.c_int, u8p, f64p, u8p, C.POINTER(Stats),
                                     C.c_int32, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                     C.c_uint32, C.c_int64, C.c_int64]
        L.oracle_rollout.restype = None
        L.oracle_spawn.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, f64p]
        L.oracle_spawn.restype = None
        L.oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.oracle_philox4x32_10.restype = None
        L.oracle_random_action.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.oracle_random_action.restype = C.c_uint
        L.oracle_fill_random_actions.argtypes = [u8p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64]
        L.oracle_fill_random_actions.restype = None
        L.oracle_moments.argtypes = [C.POINTER(C.c_float), C.c_int64, f64p]
        L.oracle_moments.restype = None
        L.oracle_gae.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), u8p, C.POINTER(C.c_float),
                                 C.c_double, C.c_double, C.c_int64, C.c_int64]
        L.oracle_gae.restype = None
        del vp
        _lib = L
    return _lib


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)


def spawn(seed, env_id, episode_index, randomize_drone=True, randomize_platform=True):
    o = (C.c_double * 4)()
    lib().oracle_spawn(seed, env_id, episode_index, int(randomize_drone), int(randomize_platform), o)
    return tuple(float(v) for v in o)


def random_actions(seed, env_id_base, t0, T, n):
    a = np.empty((T, n), np.uint8)
    lib().oracle_fill_random_actions(_p(a, C.c_uint8), seed, env_id_base, t0, T, n)
    return a


def moments(x):
    x = np.ascontiguousarray(x, np.float32).ravel()
    o = (C.c_double * 3)()
    lib().oracle_moments(_p(x, C.c_float), x.size, o)
    return tuple(float(v) for v in o)


def gae(rewards, values, dones, gamma=0.99, lam=0.95):
    rewards = np.ascontiguousarray(rewards, np.float32)
    values = np.ascontiguousarray(values, np.float32)
    dones = np.ascontiguousarray(dones, np.uint8)
    T, n = rewards.shape
    assert values.shape == (T + 1, n) and dones.shape == (T, n)
    adv = np.empty((T, n), np.float32)
    lib().oracle_gae(_p(rewards, C.c_float), _p(values, C.c_float), _p(dones, C.c_uint8),
                     _p(adv, C.c_float), gamma, lam, T, n)
    return adv


class OracleBatch:
    """n float64 environments, C-stepped."""

    def __init__(self, n, seed=0, randomize_drone=False, randomize_platform=True,
                 max_steps=0, auto_reset=False, env_id_base=0):
        self.n = int(n)
        self.seed, self.env_id_base = int(seed), int(env_id_base)
        self.randomize_drone, self.randomize_platform = bool(randomize_drone), bool(randomize_platform)
        self.max_steps, self.auto_reset = int(max_steps or 0), bool(auto_reset)
        for k in _F64:
            setattr(self, k, np.zeros(self.n, np.float64))
        self.steps = np.zeros(self.n, np.int32)
        self.episode = np.zeros(self.n, np.uint32)
        self.flags = np.zeros(self.n, np.uint8)
        self.stats = Stats()

    def _state(self, lo=0):
        s = _State()
        for k in _F64:
            setattr(s, k, _p(getattr(self, k)[lo:], C.c_double))
        s.steps = _p(self.steps[lo:], C.c_int32)
        s.episode = _p(self.episode[lo:], C.c_uint32)
        s.flags = _p(self.flags[lo:], C.c_uint8)
        return s

    def reset(self, mask=None):
        obs = np.zeros((self.n, 15), np.float64)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            self.write_obs_into(obs)
        st = self._state()
        lib().oracle_reset(C.byref(st), _p(mask, C.c_uint8), _p(obs, C.c_double), 15,
                           int(self.randomize_drone), int(self.randomize_platform),
                           self.seed, self.env_id_base, self.n)
        return obs

    def write_obs_into(self, obs):
        """obs of the current state without stepping (all-skip step)."""
        a = np.full(self.n, ACT_SKIP, np.uint8)
        st = self._state()
        lib().oracle_step(C.byref(st), _p(a, C.c_uint8), _p(obs, C.c_double), 15, None, None, None, None,
                          0, 0, 0, 0, 0, 0, self.n)
        return obs

    def inject(self, x, y, px, py):
        """``g.reset(); g.drone.reset(x, y); g.platform.reset(px, py)`` for every env."""
        self.x[:] = x; self.y[:] = y; self.px[:] = px; self.py[:] = py
        for k in ("vx", "vy", "angle", "angvel", "ep_return"):
            getattr(self, k)[:] = 0.0
        self.fuel[:] = 1000.0
        self.steps[:] = 0; self.flags[:] = 0; self.episode += 1

    def step(self, actions, want_final=False):
        actions = np.ascontiguousarray(actions, np.uint8)
        assert actions.shape == (self.n,)
        obs = np.empty((self.n, 15), np.float64)
        reward = np.empty(self.n, np.float64)
        done = np.empty(self.n, np.uint8)
        final = np.zeros((self.n, 15), np.float64) if want_final else None
        st = self._state()
        lib().oracle_step(C.byref(st), _p(actions, C.c_uint8), _p(obs, C.c_double), 15, _p(reward, C.c_double),
                          _p(done, C.c_uint8), _p(final, C.c_double), C.byref(self.stats),
                          self.max_steps, int(self.auto_reset), int(self.randomize_drone),
                          int(self.randomize_platform), self.seed, self.env_id_base, self.n)
        return (obs, reward, done, final) if want_final else (obs, reward, done)

    def rollout(self, T, policy=POL_RANDOM, actions=None, t0=0, record=False, lo=0, hi=None):
        """T steps (slice [lo, hi) of the envs -- lets several threads share one batch)."""
        hi = self.n if hi is None else hi
        m = hi - lo
        rew = np.empty((T, m), np.float64) if record else None
        don = np.empty((T, m), np.uint8) if record else None
        if actions is not None:
            actions = np.ascontiguousarray(actions, np.uint8)
            assert actions.shape == (T, m)
        st = self._state(lo)
        stats = Stats()
        lib().oracle_rollout(C.byref(st), int(policy), _p(actions, C.c_uint8), _p(rew, C.c_double),
                             _p(don, C.c_uint8), C.byref(stats), self.max_steps, int(self.auto_reset),
                             int(self.randomize_drone), int(self.randomize_platform), self.seed,
                             self.env_id_base + lo, t0, T, m)
        return rew, don, stats
