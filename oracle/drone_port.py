"""Per-instance float64 restatement of the reference ``DroneGame`` (ORACLE, test-only).

One ``PortDroneGame`` object == one reference ``DroneGame(render_mode=None)``.
It follows the reference statement order exactly and uses the same numpy scalar
calls (``np.radians/cos/sin/sqrt``, ``np.random.randint``) so that on the same
numpy build every number it produces is bit-identical to the reference's.

Reference anchors (paths under /root/reference/delivery_drone/game/):
  constants            config.py:4-5,18-20,23-29,32-36,39-40,45,54-58,61-68
  rotate / wrap / dist physics.py:6-23,26-39,42-44
  thrust + integrate   drone.py:44-76,78-103
  bottom centre etc.   drone.py:130-153
  platform bbox        platform.py:51-74
  reset / step / obs   game_engine.py:59-93,95-138,140-177
  reward + terminals   game_engine.py:179-279
  info dict            game_engine.py:281-298

It is deliberately *also* the CPU baseline that ``bench.py`` times: it does the
same per-step work as the reference (python scalars, dict construction, three
``sqrt`` calls in get_state/_get_info), so its speed stands in for the
reference's on a box where /root/reference does not exist.

Parity status: pinned against the live reference by
``tests/golden/make_golden.py`` (KAT1..KAT7 + randomised corpus) -- see
``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import numpy as np


class K:
    """The constants of config.py that the headless path reads."""

    WIDTH = 800                 # config.py:4
    HEIGHT = 600                # config.py:5
    GRAVITY = 0.3               # config.py:18
    DRAG = 0.99                 # config.py:19
    ANGULAR_DRAG = 0.95         # config.py:20
    DRONE_H = 20                # config.py:24
    MAIN_THRUST = 0.6           # config.py:25
    SIDE_THRUST = 0.3           # config.py:26
    MAX_FUEL = 1000.0           # config.py:27
    FUEL_MAIN = 2.0             # config.py:28
    FUEL_SIDE = 1.0             # config.py:29
    PLAT_W = 100                # config.py:32
    PLAT_H = 20                 # config.py:33
    PLAT_Y = 500                # config.py:34  (HEIGHT - 100)
    PLAT_Y_MIN = 100            # config.py:35
    PLAT_Y_MAX = 550            # config.py:36
    LAND_SPEED = 3.0            # config.py:39
    LAND_ANGLE = 20.0           # config.py:40
    OOB_MARGIN = 50             # config.py:45
    R_LAND = 100.0              # config.py:54
    R_CRASH = -100.0            # config.py:55
    R_FUEL = -50.0              # config.py:56
    R_OOB = -50.0               # config.py:57
    R_STEP = -0.1               # config.py:58
    START_X = 400               # config.py:61
    START_Y = 100               # config.py:62
    SPAWN_X = (100, 700)        # config.py:65-66 (inclusive)
    SPAWN_Y = (50, 250)         # config.py:67-68 (inclusive)


def _rot(x, y, deg):
    # physics.py:16-23
    rad = np.radians(deg)
    c = np.cos(rad)
    s = np.sin(rad)
    return x * c - y * s, x * s + y * c


def _dist(x1, y1, x2, y2):
    # physics.py:42-44
    return np.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)


class PortDroneGame:
    """Headless stand-in for ``DroneGame`` with the same public surface."""

    def __init__(self, render_mode=None, randomize_drone=False, randomize_platform=True):
        if render_mode is not None:
            raise ValueError("the oracle port is headless only")
        self.randomize_drone = randomize_drone
        self.randomize_platform = randomize_platform
        # drone (drone.py:12-42)
        self.x = K.START_X
        self.y = K.START_Y
        self.vx = 0.0
        self.vy = 0.0
        self.angle = 0.0
        self.angvel = 0.0
        self.fuel = K.MAX_FUEL
        self.crashed = False
        self.landed = False
        # platform (game_engine.py:41-46)
        self.px = K.WIDTH // 2
        self.py = K.PLAT_Y
        # counters (game_engine.py:49-53)
        self.steps = 0
        self.total_reward = 0
        self.episode = 0
        self.done = False

    # -- reset: game_engine.py:59-93, drone.py:221-238, platform.py:104-114 --
    def reset(self):
        if self.randomize_drone:
            sx = np.random.randint(K.SPAWN_X[0], K.SPAWN_X[1] + 1)
            sy = np.random.randint(K.SPAWN_Y[0], K.SPAWN_Y[1] + 1)
        else:
            sx, sy = K.START_X, K.START_Y
        self.inject(sx, sy)
        if self.randomize_platform:
            self.px = np.random.randint(K.PLAT_W // 2 + 50, K.WIDTH - K.PLAT_W // 2 - 50)
            self.py = np.random.randint(K.PLAT_Y_MIN, K.PLAT_Y_MAX)
        else:
            self.px, self.py = K.WIDTH // 2, K.PLAT_Y
        self.steps = 0
        self.total_reward = 0
        self.done = False
        self.episode += 1
        return self.get_state()

    def inject(self, x, y, px=None, py=None):
        """Parity helper == ``g.drone.reset(x, y); g.platform.reset(px, py)``."""
        self.x, self.y = x, y
        self.vx = self.vy = 0.0
        self.angle = self.angvel = 0.0
        self.fuel = K.MAX_FUEL
        self.crashed = self.landed = False
        if px is not None:
            self.px, self.py = px, py

    # -- step: game_engine.py:95-138 --
    def step(self, action):
        if self.done:                                   # game_engine.py:107-111
            info = self.info()
            info["needs_reset"] = True
            return self.get_state(), 0, True, info

        main = bool(action.get("main_thrust", 0))
        left = bool(action.get("left_thrust", 0))
        right = bool(action.get("right_thrust", 0))

        # drone.py:58-76 -- fuel is re-tested before every thruster
        if main and self.fuel > 0:
            tx, ty = _rot(0, -K.MAIN_THRUST, self.angle)
            self.vx += tx
            self.vy += ty
            self.fuel -= K.FUEL_MAIN
        if left and self.fuel > 0:
            self.angvel -= K.SIDE_THRUST
            self.fuel -= K.FUEL_SIDE
        if right and self.fuel > 0:
            self.angvel += K.SIDE_THRUST
            self.fuel -= K.FUEL_SIDE
        self.fuel = max(0, self.fuel)

        # drone.py:88-103 (dt == 1.0)
        self.vy += K.GRAVITY * 1.0
        self.vx *= K.DRAG
        self.vy *= K.DRAG
        self.x += self.vx * 1.0
        self.y += self.vy * 1.0
        self.angle += self.angvel * 1.0
        self.angvel *= K.ANGULAR_DRAG
        a = self.angle                                   # physics.py:35-39
        while a > 180:
            a -= 360
        while a < -180:
            a += 360
        self.angle = a

        r = self._reward()
        self.total_reward += r
        self.steps += 1
        return self.get_state(), r, self.done, self.info()

    # -- helpers used by the terminal tests --
    def _bottom_on_platform(self):
        ox, oy = _rot(0, K.DRONE_H / 2, self.angle)      # drone.py:136-137
        bx, by = self.x + ox, self.y + oy
        left = self.px - K.PLAT_W / 2                     # platform.py:57-60
        right = self.px + K.PLAT_W / 2
        top = self.py - K.PLAT_H / 2
        bottom = self.py + K.PLAT_H / 2
        return (left <= bx <= right) and (top <= by <= bottom)   # platform.py:74

    def _speed(self):
        return np.sqrt(self.vx ** 2 + self.vy ** 2)      # drone.py:145

    def _upright(self):
        return abs(self.angle) <= K.LAND_ANGLE           # drone.py:153

    def _landing(self):
        # game_engine.py:224-242 (short-circuit order preserved)
        if self.crashed or self.landed:
            return False
        if not self._bottom_on_platform():
            return False
        if self._speed() > K.LAND_SPEED:
            return False
        if not self._upright():
            return False
        return True

    def _crash(self):
        # game_engine.py:250-267
        if self.crashed:
            return True
        if self.y > K.HEIGHT - 50:
            if not self._bottom_on_platform():
                return True
            if self._speed() > K.LAND_SPEED:
                return True
            if not self._upright():
                return True
        return False

    def _oob(self):
        # game_engine.py:275-279
        m = K.OOB_MARGIN
        return (self.x < -m or self.x > K.WIDTH + m or self.y < -m or self.y > K.HEIGHT + m)

    def _reward(self):
        # game_engine.py:185-216 -- priority: land, crash, fuel, oob, shaping
        r = K.R_STEP
        if self._landing():
            self.landed = True
            self.done = True
            return r + K.R_LAND
        if self._crash():
            self.crashed = True
            self.done = True
            return r + K.R_CRASH
        if self.fuel <= 0:
            self.crashed = True
            self.done = True
            return r + K.R_FUEL
        if self._oob():
            self.crashed = True
            self.done = True
            return r + K.R_OOB
        d = _dist(self.x, self.y, self.px, self.py)
        return r + (500 - d) / 5000

    # -- observation: game_engine.py:146-177 --
    def get_state(self):
        dx = self.px - self.x
        dy = self.py - self.y
        d = _dist(self.x, self.y, self.px, self.py)
        return {
            "drone_x": self.x / K.WIDTH,
            "drone_y": self.y / K.HEIGHT,
            "drone_vx": self.vx / 10.0,
            "drone_vy": self.vy / 10.0,
            "drone_angle": self.angle / 180.0,
            "drone_angular_vel": self.angvel / 10.0,
            "drone_fuel": self.fuel / K.MAX_FUEL,
            "platform_x": self.px / K.WIDTH,
            "platform_y": self.py / K.HEIGHT,
            "distance_to_platform": d / K.WIDTH,
            "dx_to_platform": dx / K.WIDTH,
            "dy_to_platform": dy / K.HEIGHT,
            "speed": self._speed() / 10.0,
            "landed": self.landed,
            "crashed": self.crashed,
            "steps": self.steps,
        }

    # -- info: game_engine.py:287-298 --
    def info(self):
        return {
            "steps": self.steps,
            "total_reward": self.total_reward,
            "episode": self.episode,
            "fuel_remaining": self.fuel,
            "distance_to_platform": _dist(self.x, self.y, self.px, self.py),
            "speed": self._speed(),
            "angle": self.angle,
        }

    _get_info = info


OBS_KEYS = (
    "drone_x", "drone_y", "drone_vx", "drone_vy", "drone_angle", "drone_angular_vel",
    "drone_fuel", "platform_x", "platform_y", "distance_to_platform", "dx_to_platform",
    "dy_to_platform", "speed", "landed", "crashed",
)


def obs_vector(state: dict) -> np.ndarray:
    """The 15-vector the policy sees (Actor_Critic_PPO.ipynb c10:L3-19)."""
    return np.array([float(state[k]) for k in OBS_KEYS], dtype=np.float64)


# ---------------------------------------------------------------------------
# CPU-baseline workload (BASELINE.json configs[0]; SURVEY.md 8d "Cfg 1")
# ---------------------------------------------------------------------------
def cpu_rollout_worker(args):
    """One process: ``games`` instances, uniform random actions, reset on done
    or after ``cap`` steps, for ``ticks`` timed ticks after ``warmup`` untimed ones (optional
    fifth element).  Returns (env_steps, seconds)."""
    import time

    seed, ticks, games, cap = args[:4]
    warmup = args[4] if len(args) > 4 else 0
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    envs = [PortDroneGame(None, True, True) for _ in range(games)]
    for e in envs:
        e.reset()
    acts = rng.integers(0, 2, (ticks + warmup, games, 3))
    t0 = time.perf_counter()
    n = 0
    for t in range(ticks + warmup):
        if t == warmup:
            t0 = time.perf_counter()
            n = 0
        row = acts[t]
        for gi, e in enumerate(envs):
            a = row[gi]
            _, _, done, _ = e.step({"main_thrust": a[0], "left_thrust": a[1], "right_thrust": a[2]})
            n += 1
            if done or e.steps >= cap:
                e.reset()
    return n, time.perf_counter() - t0
