/*
 * drone_oracle.c -- batched float64 CPU restatement of the delivery_drone hot path.
 *
 * ORACLE / TEST INFRASTRUCTURE ONLY.  Nothing in the product package links or
 * loads this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg do, and only as the checker.
 *
 * It states, in plain C doubles and in the reference's statement order, what one
 * DroneGame.step()/reset() does, for n independent environments held as
 * structure-of-arrays.  Reference anchors (under /root/reference/delivery_drone/game/):
 *   config.py:4-5,18-20,23-29,32-36,39-40,45,54-58,61-68   constants
 *   physics.py:6-23,26-39,42-44                            rotate_point / normalize_angle / distance
 *   drone.py:44-76,78-103,130-153,221-238                  apply_thrust / update / bottom centre / reset
 *   platform.py:51-74                                      closed bbox test
 *   game_engine.py:59-93,95-138,140-177,179-279            reset / step / get_state / reward + terminals
 *
 * Parity status: PINNED against the live reference -- tests/golden/make_golden.py
 * imports the unmodified reference DroneGame and records KAT1..7 and a randomised
 * corpus; tests/test_oracle_golden.py checks this file against them.  (libm
 * sin/cos may differ from numpy's by <= 1 ulp; everything else is bit-identical.
 * Build with -ffp-contract=off so no FMA contraction changes a rounding.)
 *
 * Things that are NOT in the reference and are specified here instead (the CUDA
 * path must match them bit for bit):
 *   - spawn randomisation draws come from Philox4x32-10 keyed by (seed) with
 *     counter (env_id_lo, env_id_hi, episode_index, 0): numpy's global MT19937
 *     stream cannot be reproduced per-env on a GPU.  Ranges are the reference's
 *     (game_engine.py:66-83): x in [100,700], y in [50,250], px in [100,699],
 *     py in [100,549]; value = lo + mulhi32(r, span).
 *   - same-step auto-reset and max_steps truncation (the reference freezes after
 *     done, game_engine.py:107-111, and leaves time-outs to the notebooks,
 *     Actor_Critic_PPO.ipynb c16:L89-93).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define DD_DONE       0x01
#define DD_LANDED     0x02
#define DD_CRASHED    0x04
#define DD_TRUNCATED  0x08
#define DD_CAUSE_GROUND 0x10
#define DD_CAUSE_FUEL   0x20
#define DD_CAUSE_OOB    0x30

#define ACT_MAIN  0x01
#define ACT_LEFT  0x02
#define ACT_RIGHT 0x04
#define ACT_SKIP  0x80

typedef struct {
    double *x, *y, *vx, *vy, *angle, *angvel, *fuel;  /* drone.py:19-32 */
    double *px, *py;                                   /* platform.py:19-20 */
    double *ep_return;                                 /* game_engine.py:51 total_reward */
    int32_t *steps;                                    /* game_engine.py:50 */
    uint32_t *episode;                                 /* game_engine.py:52 */
    uint8_t *flags;                                    /* done/landed/crashed (+truncated, cause) */
} OracleState;

typedef struct {
    uint64_t episodes, landed, crashed, truncated;
    double sum_return, sum_length;
} OracleStats;

/* ---- Philox4x32-10 (Salmon et al., SC'11; Random123 v1.09 constants) ---- */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    philox4x32_10(ctr, key, out);
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* spawn draw: (x, y, px, py), all integer valued */
void oracle_spawn(uint64_t seed, uint64_t env_id, uint32_t episode_index,
                  int randomize_drone, int randomize_platform, double out[4])
{
    uint32_t ctr[4] = { (uint32_t)env_id, (uint32_t)(env_id >> 32), episode_index, 0u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t r[4];
    philox4x32_10(ctr, key, r);
    out[0] = randomize_drone ? 100.0 + (double)mulhi32(r[0], 601u) : 400.0;   /* game_engine.py:66-71 */
    out[1] = randomize_drone ? 50.0 + (double)mulhi32(r[1], 201u) : 100.0;
    out[2] = randomize_platform ? 100.0 + (double)mulhi32(r[2], 600u) : 400.0; /* game_engine.py:74-85 */
    out[3] = randomize_platform ? 100.0 + (double)mulhi32(r[3], 450u) : 500.0;
}

/* ---- observation: game_engine.py:146-177, first 15 keys in policy order ---- */
static void write_obs(double *o, double x, double y, double vx, double vy, double angle,
                      double angvel, double fuel, double px, double py, int landed, int crashed)
{
    double dx = px - x, dy = py - y;
    double dist = sqrt((px - x) * (px - x) + (py - y) * (py - y));   /* physics.py:44 */
    o[0] = x / 800.0;
    o[1] = y / 600.0;
    o[2] = vx / 10.0;
    o[3] = vy / 10.0;
    o[4] = angle / 180.0;
    o[5] = angvel / 10.0;
    o[6] = fuel / 1000.0;
    o[7] = px / 800.0;
    o[8] = py / 600.0;
    o[9] = dist / 800.0;
    o[10] = dx / 800.0;
    o[11] = dy / 600.0;
    o[12] = sqrt(vx * vx + vy * vy) / 10.0;                          /* drone.py:145 */
    o[13] = landed ? 1.0 : 0.0;
    o[14] = crashed ? 1.0 : 0.0;
}

static int bottom_on_platform(double x, double y, double angle, double px, double py)
{
    /* drone.py:136-137 rotate_point(0, 10, angle); platform.py:57-74 closed bbox */
    double rad = angle * (M_PI / 180.0);       /* np.radians == x * (pi/180) */
    double c = cos(rad), s = sin(rad);
    double bx = x + (0.0 * c - 10.0 * s);
    double by = y + (0.0 * s + 10.0 * c);
    double left = px - 100.0 / 2, right = px + 100.0 / 2;
    double top = py - 20.0 / 2, bottom = py + 20.0 / 2;
    return (left <= bx && bx <= right) && (top <= by && by <= bottom);
}

static void reset_one(OracleState *s, int64_t i, uint64_t seed, uint64_t env_id,
                      int randomize_drone, int randomize_platform)
{
    double sp[4];
    oracle_spawn(seed, env_id, s->episode[i], randomize_drone, randomize_platform, sp);
    s->x[i] = sp[0]; s->y[i] = sp[1]; s->px[i] = sp[2]; s->py[i] = sp[3];
    s->vx[i] = 0.0; s->vy[i] = 0.0; s->angle[i] = 0.0; s->angvel[i] = 0.0;  /* drone.py:229-234 */
    s->fuel[i] = 1000.0;
    s->steps[i] = 0; s->ep_return[i] = 0.0; s->flags[i] = 0;                  /* game_engine.py:88-90 */
    s->episode[i] += 1;                                                       /* game_engine.py:91 */
}

void oracle_reset(OracleState *s, const uint8_t *mask, double *obs, int obs_stride,
                  int randomize_drone, int randomize_platform,
                  uint64_t seed, uint64_t env_id_base, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        if (mask && !mask[i]) continue;
        reset_one(s, i, seed, env_id_base + (uint64_t)i, randomize_drone, randomize_platform);
        if (obs)
            write_obs(obs + (size_t)i * obs_stride, s->x[i], s->y[i], 0, 0, 0, 0, 1000.0, s->px[i], s->py[i], 0, 0);
    }
}

/* One DroneGame.step for env i.  Returns the reward; updates state in place. */
static double step_one(OracleState *s, int64_t i, unsigned act)
{
    double x = s->x[i], y = s->y[i], vx = s->vx[i], vy = s->vy[i];
    double angle = s->angle[i], angvel = s->angvel[i], fuel = s->fuel[i];
    const double px = s->px[i], py = s->py[i];

    /* drone.py:58-76 : sequential fuel gating */
    if ((act & ACT_MAIN) && fuel > 0) {
        double rad = angle * (M_PI / 180.0);
        double c = cos(rad), sn = sin(rad);
        double tx = 0.0 * c - (-0.6) * sn;     /* physics.py:20 */
        double ty = 0.0 * sn + (-0.6) * c;     /* physics.py:21 */
        vx += tx; vy += ty; fuel -= 2.0;
    }
    if ((act & ACT_LEFT) && fuel > 0) { angvel -= 0.3; fuel -= 1.0; }
    if ((act & ACT_RIGHT) && fuel > 0) { angvel += 0.3; fuel -= 1.0; }
    if (fuel < 0) fuel = 0;

    /* drone.py:88-103 */
    vy += 0.3;
    vx *= 0.99; vy *= 0.99;
    x += vx; y += vy;
    angle += angvel;
    angvel *= 0.95;
    while (angle > 180) angle -= 360;
    while (angle < -180) angle += 360;

    /* game_engine.py:185-216 */
    double r = -0.1;
    unsigned f = 0;
    double speed = sqrt(vx * vx + vy * vy);
    int on = bottom_on_platform(x, y, angle, px, py);
    if (on && !(speed > 3.0) && fabs(angle) <= 20.0) {          /* :224-242 */
        f = DD_DONE | DD_LANDED; r += 100.0;
    } else if (y > 600 - 50) {                                   /* :256-265 (== y>550 once landing failed) */
        f = DD_DONE | DD_CRASHED | DD_CAUSE_GROUND; r += -100.0;
    } else if (fuel <= 0) {                                      /* :200-204 */
        f = DD_DONE | DD_CRASHED | DD_CAUSE_FUEL; r += -50.0;
    } else if (x < -50 || x > 800 + 50 || y < -50 || y > 600 + 50) {   /* :275-279 */
        f = DD_DONE | DD_CRASHED | DD_CAUSE_OOB; r += -50.0;
    } else {
        double d = sqrt((px - x) * (px - x) + (py - y) * (py - y));
        r += (500 - d) / 5000;                                   /* :213-214 */
    }
    s->x[i] = x; s->y[i] = y; s->vx[i] = vx; s->vy[i] = vy;
    s->angle[i] = angle; s->angvel[i] = angvel; s->fuel[i] = fuel;
    s->ep_return[i] += r;                                        /* :131 */
    s->steps[i] += 1;                                            /* :132 */
    s->flags[i] = (uint8_t)f;
    return r;
}

/*
 * Batched step.  auto_reset == 0 reproduces the reference's freeze-after-done
 * (reward 0, done stays set, no state change).  auto_reset != 0: on the
 * terminating step, reward/done_flags/final_obs describe the finished episode
 * and obs/state hold the first observation of the next one.
 */
void oracle_step(OracleState *s, const uint8_t *actions,
                 double *obs, int obs_stride, double *reward, uint8_t *done_flags,
                 double *final_obs, OracleStats *stats,
                 int32_t max_steps, int auto_reset, int randomize_drone, int randomize_platform,
                 uint64_t seed, uint64_t env_id_base, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        unsigned act = actions[i];
        double r = 0.0;
        unsigned f = s->flags[i];
        if (!(act & ACT_SKIP) && !(f & DD_DONE)) {
            r = step_one(s, i, act);
            f = s->flags[i];
            if (!(f & DD_DONE) && max_steps > 0 && s->steps[i] >= max_steps) {
                f = DD_DONE | DD_TRUNCATED;
                s->flags[i] = (uint8_t)f;
            }
            if ((f & DD_DONE) && stats) {
                stats->episodes += 1;
                stats->landed += (f & DD_LANDED) ? 1 : 0;
                stats->crashed += (f & DD_CRASHED) ? 1 : 0;
                stats->truncated += (f & DD_TRUNCATED) ? 1 : 0;
                stats->sum_return += s->ep_return[i];
                stats->sum_length += (double)s->steps[i];
            }
        }
        if (reward) reward[i] = r;
        if (done_flags) done_flags[i] = (uint8_t)f;
        int fresh_done = (f & DD_DONE) && !(act & ACT_SKIP);
        if (final_obs && fresh_done)
            write_obs(final_obs + (size_t)i * obs_stride, s->x[i], s->y[i], s->vx[i], s->vy[i], s->angle[i],
                      s->angvel[i], s->fuel[i], s->px[i], s->py[i], (f & DD_LANDED) != 0, (f & DD_CRASHED) != 0);
        if (auto_reset && fresh_done)
            reset_one(s, i, seed, env_id_base + (uint64_t)i, randomize_drone, randomize_platform);
        if (obs) {
            unsigned g = s->flags[i];
            write_obs(obs + (size_t)i * obs_stride, s->x[i], s->y[i], s->vx[i], s->vy[i], s->angle[i],
                      s->angvel[i], s->fuel[i], s->px[i], s->py[i], (g & DD_LANDED) != 0, (g & DD_CRASHED) != 0);
        }
    }
}

/* 3-bit action for (env, t) of the synthetic random policy (p = 0.5 per thruster):
 * one Philox call, counter (env_lo, env_hi, t / 32, 1), serves 32 consecutive steps;
 * action k = t % 32 is bits [3k, 3k+3) of the 96-bit string r0 | r1<<32 | r2<<64. */
unsigned oracle_random_action(uint64_t seed, uint64_t env_id, uint32_t t)
{
    uint32_t ctr[4] = { (uint32_t)env_id, (uint32_t)(env_id >> 32), t >> 5, 1u };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t r[4];
    philox4x32_10(ctr, key, r);
    unsigned k = t & 31u, bit = 3u * k;
    unsigned w = bit >> 5, sh = bit & 31u;
    uint64_t two = (uint64_t)r[w] | ((uint64_t)r[w + 1] << 32);   /* w+1 <= 3 always valid */
    return (unsigned)((two >> sh) & 7u);
}

void oracle_fill_random_actions(uint8_t *actions, uint64_t seed, uint64_t env_id_base,
                                uint32_t t0, int64_t T, int64_t n)
{
    for (int64_t t = 0; t < T; ++t)
        for (int64_t i = 0; i < n; ++i)
            actions[t * n + i] = (uint8_t)oracle_random_action(seed, env_id_base + (uint64_t)i, t0 + (uint32_t)t);
}

/* policy ids for rollouts that generate their own actions */
#define POL_TRACE    0   /* actions[t*n+i] */
#define POL_RANDOM   1   /* oracle_random_action */
#define POL_BANGBANG 2   /* main = vy > 1.5 (SURVEY 8d cfg 3 (ii); KAT6) */

/* T steps of n envs; the CPU baseline ("port" in C) and the checker for the
 * CUDA rollout kernel.  Optional per-step outputs are [T, n] (reward, done). */
void oracle_rollout(OracleState *s, int policy, const uint8_t *actions,
                    double *reward_tn, uint8_t *done_tn, OracleStats *stats,
                    int32_t max_steps, int auto_reset, int randomize_drone, int randomize_platform,
                    uint64_t seed, uint64_t env_id_base, uint32_t t0, int64_t T, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) {
        OracleStats local; memset(&local, 0, sizeof local);
        for (int64_t t = 0; t < T; ++t) {
            uint8_t a;
            if (policy == POL_TRACE) a = actions[t * n + i];
            else if (policy == POL_RANDOM) a = (uint8_t)oracle_random_action(seed, env_id_base + (uint64_t)i, t0 + (uint32_t)t);
            else a = (s->vy[i] > 1.5) ? ACT_MAIN : 0;
            OracleState one = { s->x + i, s->y + i, s->vx + i, s->vy + i, s->angle + i, s->angvel + i,
                                s->fuel + i, s->px + i, s->py + i, s->ep_return + i, s->steps + i,
                                s->episode + i, s->flags + i };
            double r; uint8_t f;
            oracle_step(&one, &a, NULL, 15, &r, &f, NULL, &local, max_steps, auto_reset,
                        randomize_drone, randomize_platform, seed, env_id_base + (uint64_t)i, 1);
            if (reward_tn) reward_tn[t * n + i] = r;
            if (done_tn) done_tn[t * n + i] = f;
        }
        if (stats) {
            {
                stats->episodes += local.episodes; stats->landed += local.landed;
                stats->crashed += local.crashed; stats->truncated += local.truncated;
                stats->sum_return += local.sum_return; stats->sum_length += local.sum_length;
            }
        }
    }
}

/* K4 checker: n, sum, sum of squares (Actor_Critic_PPO.ipynb c21:L105) */
void oracle_moments(const float *x, int64_t n, double out[3])
{
    double s = 0, q = 0;
    for (int64_t i = 0; i < n; ++i) { s += x[i]; q += (double)x[i] * x[i]; }
    out[0] = (double)n; out[1] = s; out[2] = q;
}

/* N3 checker: compute_gae (Actor_Critic_PPO.ipynb c15:L49-53) over [T, n] with
 * bootstrap values[T*n .. (T+1)*n).  fp32 arithmetic like the torch original. */
void oracle_gae(const float *rewards, const float *values, const uint8_t *dones,
                float *adv, double gamma, double lambda, int64_t T, int64_t n)
{
    /* python floats are combined in double first, then meet fp32 tensors */
    const float g = (float)gamma, gl = (float)(gamma * lambda);
    for (int64_t i = 0; i < n; ++i) {
        float gae = 0.0f;
        for (int64_t t = T - 1; t >= 0; --t) {
            float mask = 1.0f - (dones[t * n + i] ? 1.0f : 0.0f);
            float delta = rewards[t * n + i] + (g * values[(t + 1) * n + i]) * mask - values[t * n + i];
            gae = delta + (gl * mask) * gae;
            adv[t * n + i] = gae;
        }
    }
}
