/*
 * vss_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, one field at a time) of the VSS hot path of
 * FelipeMartins96/rsoccer-isaac-cleanrl. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (rsoccer_isaac_cleanrl_b200/) never does.
 *
 * What is pinned and what is not
 *   - obs / rewards / dones (reference envs/vss.py:530-655): pinned against golden
 *     vectors produced by executing the reference's own torch-jit functions
 *     (tests/golden/make_golden.py, run in the build container).
 *   - GAE (ppo_continuous_action_isaacgym.py:282-296): pinned the same way.
 *   - step sequencing (envs/vss.py:180-203 + VecTask.step) and the agent views
 *     (envs/wrappers.py:89-180): restated from source; VecTask is an un-vendored
 *     dependency (IsaacGymEnvs @ dee7c56) -> its part is "parity unpinned".
 *   - physics: PARITY UNPINNED. The reference delegates to closed-source PhysX
 *     (IsaacGym Preview, not in the tree, not installable). orc_physics() is a
 *     double-precision restatement of the NEW 2-D model specified in DESIGN.md §3,
 *     written independently of the CUDA kernel from that text.
 *   - reset RNG: torch's global generator cannot be reproduced by a fused kernel;
 *     both sides use Philox4x32-10 keyed by (seed, global field id, episode), checked
 *     against the Random123 known-answer vectors. Distributional parity with
 *     envs/vss.py:267-333 is tested statistically.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/vss_b200.h"

#define ORC_API __attribute__((visibility("default")))

#define NT 2  /* teams */
#define NR 3  /* robots per team */
#define NB 6  /* robots per field */

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11; Random123). Counter-based RNG.        */
/* ------------------------------------------------------------------------- */
static void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}

/* [0,1) with 24 bits, like torch.rand for float32 */
static float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
/* (0,1] for log() */
static float u01_open(uint32_t x) { return (float)((x >> 8) + 1u) * (1.0f / 16777216.0f); }

/* RNG streams: ctr = {gid_lo, gid_hi, a, (stream << 28) | b} */
#define STREAM_RESET_POS 0u
#define STREAM_RESET_MISC 1u
#define STREAM_OU 2u

static void rng_block(uint64_t seed, uint64_t gid, uint32_t a, uint32_t stream, uint32_t b,
                      uint32_t out[4]) {
  const uint32_t ctr[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), a, (stream << 28) | b};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  orc_philox4x32_10(ctr, key, out);
}

/* ------------------------------------------------------------------------- */
/* State: arrays in the reference's tensor layouts (envs/vss.py:112-141).      */
/* ------------------------------------------------------------------------- */
typedef struct orc_state {
  int64_t n;
  float* ball_pos;  /* (N,2)     self.ball_pos      */
  float* ball_vel;  /* (N,2)     self.ball_vel      */
  float* r_pos;     /* (N,2,3,2) self.robots_pos    */
  float* r_vel;     /* (N,2,3,2) self.robots_vel    */
  float* r_rot;     /* (N,2,3,2) cos/sin of the yaw held by self.robots_quats */
  float* r_w;       /* (N,2,3)   self.robots_ang_vel */
  float* r_act;     /* (N,2,3,2) self.dof_velocity_buf */
  int64_t* progress;/* (N)       self.progress_buf  */
  uint32_t* episode;/* (N)       reset counter (RNG key; not in the reference) */
} orc_state;

/* ------------------------------------------------------------------------- */
/* compute_obs — envs/vss.py:530-575                                           */
/* ------------------------------------------------------------------------- */
static const int PERMS[3][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}}; /* vss.py:173-175 */
static const float MIRROR[9] = {-1.f, -1.f, -1.f, -1.f, -1.f, -1.f, 1.f, 1.f, 1.f}; /* :533 */

static void obs_field(int64_t n, const float* b_pos, const float* b_vel, const float* r_pos,
                      const float* r_vel, const float* r_rot, const float* r_w, const float* r_acts,
                      float* obs) {
  {
    float ball[4] = {b_pos[2 * n], b_pos[2 * n + 1], b_vel[2 * n], b_vel[2 * n + 1]}; /* :539 */
    float robots[NT][NR][9];                                                          /* :541-551 */
    for (int t = 0; t < NT; ++t)
      for (int j = 0; j < NR; ++j) {
        const int64_t r = (n * NT + t) * NR + j;
        float* f = robots[t][j];
        f[0] = r_pos[2 * r]; f[1] = r_pos[2 * r + 1];
        f[2] = r_vel[2 * r]; f[3] = r_vel[2 * r + 1];
        f[4] = r_rot[2 * r]; f[5] = r_rot[2 * r + 1]; /* cos(yaw), sin(yaw) */
        f[6] = r_w[r];
        f[7] = r_acts[2 * r]; f[8] = r_acts[2 * r + 1];
      }
    float* o = obs + n * VSS_OBS_PER_FIELD;
    /* blue rows, :552-559 */
    for (int i = 0; i < NR; ++i) {
      float* row = o + (0 * NR + i) * VSS_NUM_OBS;
      int k = 0;
      for (int c = 0; c < 4; ++c) row[k++] = ball[c];
      for (int m = 0; m < NR; ++m)
        for (int c = 0; c < 9; ++c) row[k++] = robots[0][PERMS[i][m]][c];
      for (int m = 0; m < NR; ++m)
        for (int c = 0; c < 7; ++c) row[k++] = robots[1][m][c];
    }
    /* robots *= mirror_tensor, :560 */
    for (int t = 0; t < NT; ++t)
      for (int j = 0; j < NR; ++j)
        for (int c = 0; c < 9; ++c) robots[t][j][c] *= MIRROR[c];
    /* yellow rows, :561-574 */
    for (int i = 0; i < NR; ++i) {
      float* row = o + (1 * NR + i) * VSS_NUM_OBS;
      int k = 0;
      for (int c = 0; c < 4; ++c) row[k++] = -ball[c];
      for (int m = 0; m < NR; ++m)
        for (int c = 0; c < 9; ++c) row[k++] = robots[1][PERMS[i][m]][c];
      for (int m = 0; m < NR; ++m)
        for (int c = 0; c < 7; ++c) row[k++] = robots[0][m][c];
    }
  }
}

ORC_API void orc_compute_obs(int64_t N, const float* b_pos, const float* b_vel, const float* r_pos,
                             const float* r_vel, const float* r_rot, const float* r_w,
                             const float* r_acts, float* obs /* (N,2,3,52) */) {
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) obs_field(n, b_pos, b_vel, r_pos, r_vel, r_rot, r_w, r_acts, obs);
}

/* ------------------------------------------------------------------------- */
/* rewards and dones — envs/vss.py:578-655                                     */
/* ------------------------------------------------------------------------- */
static float norm2f(float x, float y) { return sqrtf(x * x + y * y); }

/* compute_goal_rew, vss.py:578-594: (N,2,3) int64 */
ORC_API void orc_compute_goal_rew(int64_t N, const float* ball_pos, float field_width,
                                  float goal_height, int64_t* out) {
  for (int64_t n = 0; n < N; ++n) {
    const float bx = ball_pos[2 * n], by = ball_pos[2 * n + 1];
    const int is_goal = (fabsf(bx) > (field_width / 2)) && (fabsf(by) < (goal_height / 2));
    int64_t g = 0;
    if (is_goal && bx > 0) g = 1;
    if (is_goal && bx < 0) g = -1;
    for (int j = 0; j < NR; ++j) {
      out[(n * NT + 0) * NR + j] = g;
      out[(n * NT + 1) * NR + j] = -g;
    }
  }
}

/* compute_grad_rew, vss.py:597-612 */
static float ball_potential(float bx, float by, float gx, float gy) {
  const float d_left = norm2f(bx - (-gx), by - (-gy));
  const float d_right = norm2f(bx - gx, by - gy);
  return d_left - d_right;
}
ORC_API void orc_compute_grad_rew(int64_t N, const float* prev_ball_pos, const float* ball_pos,
                                  const float* yellow_goal, float* out) {
  for (int64_t n = 0; n < N; ++n) {
    const float prev_pot =
        ball_potential(prev_ball_pos[2 * n], prev_ball_pos[2 * n + 1], yellow_goal[0], yellow_goal[1]);
    const float pot = ball_potential(ball_pos[2 * n], ball_pos[2 * n + 1], yellow_goal[0], yellow_goal[1]);
    const float grad = pot - prev_pot;
    for (int j = 0; j < NR; ++j) {
      out[(n * NT + 0) * NR + j] = grad;
      out[(n * NT + 1) * NR + j] = -grad;
    }
  }
}

/* compute_move_rew, vss.py:615-625 */
ORC_API void orc_compute_move_rew(int64_t N, const float* p_robots, const float* robots,
                                  const float* p_ball, const float* ball, float* out) {
  for (int64_t n = 0; n < N; ++n)
    for (int r = 0; r < NB; ++r) {
      const int64_t i = n * NB + r;
      const float p_dist = norm2f(p_robots[2 * i] - p_ball[2 * n], p_robots[2 * i + 1] - p_ball[2 * n + 1]);
      const float dist = norm2f(robots[2 * i] - ball[2 * n], robots[2 * i + 1] - ball[2 * n + 1]);
      out[i] = p_dist - dist;
    }
}

/* compute_energy_rew, vss.py:628-631 */
ORC_API void orc_compute_energy_rew(int64_t N, const float* actions, float* out) {
  for (int64_t i = 0; i < N * NB; ++i)
    out[i] = -((fabsf(actions[2 * i]) + fabsf(actions[2 * i + 1])) / 2.0f);
}

/* compute_vss_dones, vss.py:634-655 */
ORC_API void orc_compute_dones(int64_t N, const float* ball_pos, const int64_t* progress,
                               float max_episode_length, float field_width, float goal_height,
                               int64_t* reset) {
  for (int64_t n = 0; n < N; ++n) {
    const int is_goal = (fabsf(ball_pos[2 * n]) > (field_width / 2)) &&
                        (fabsf(ball_pos[2 * n + 1]) < (goal_height / 2));
    int64_t r = 0;
    if (is_goal) r = 1;
    if ((float)progress[n] >= max_episode_length) r = 1;
    reset[n] = r;
  }
}

/* compute_rewards_and_dones for ONE field, vss.py:218-265 (after the refresh). */
static void rewards_one(const vss_params* p, const float prev_ball[2], const float prev_rpos[NB][2],
                        const float ball[2], const float rpos[NB][2], const float acts[NB][2],
                        float rew[NB][4]) {
  const float field_width = 2.0f * p->field_half_length, goal_height = 2.0f * p->goal_half_width;
  const float yellow_goal[2] = {field_width / 2, 0.0f}; /* vss.py:154-159 */
  for (int r = 0; r < NB; ++r)
    for (int c = 0; c < 4; ++c) rew[r][c] = 0.0f; /* rew_buf *= 0, :223 */
  if (p->w_goal > 0) { /* :225-232 */
    int64_t g[NB];
    orc_compute_goal_rew(1, ball, field_width, goal_height, g);
    for (int r = 0; r < NB; ++r) rew[r][0] = (float)g[r] * p->w_goal;
  }
  if (p->w_grad > 0) { /* :234-239 */
    float g[NB];
    orc_compute_grad_rew(1, prev_ball, ball, yellow_goal, g);
    for (int r = 0; r < NB; ++r) rew[r][1] = g[r] * p->w_grad;
  }
  if (p->w_move > 0) { /* :241-251 */
    float g[NB];
    orc_compute_move_rew(1, &prev_rpos[0][0], &rpos[0][0], prev_ball, ball, g);
    for (int r = 0; r < NB; ++r) rew[r][2] += g[r] * p->w_move;
  }
  if (p->w_energy > 0) { /* :253-255 */
    float g[NB];
    orc_compute_energy_rew(1, &acts[0][0], g);
    for (int r = 0; r < NB; ++r) rew[r][3] += g[r] * p->w_energy;
  }
}

/* ------------------------------------------------------------------------- */
/* reset_dones for ONE field — envs/vss.py:267-333, RNG = Philox streams       */
/* ------------------------------------------------------------------------- */
#define RESET_MAX_ATTEMPTS 64

static void reset_one(const vss_params* p, uint64_t seed, uint64_t gid, orc_state* s, int64_t n) {
  const uint32_t ep = s->episode[n];
  float pos[7][2]; /* entity 0 = ball, 1..6 = robots, vss.py:274-279 */
  const float min_d2 = p->min_placement_dist * p->min_placement_dist;
  for (uint32_t attempt = 0; attempt < RESET_MAX_ATTEMPTS; ++attempt) {
    uint32_t u[16];
    for (uint32_t b = 0; b < 4; ++b) rng_block(seed, gid, ep, STREAM_RESET_POS, (attempt << 2) | b, u + 4 * b);
    for (int e = 0; e < 7; ++e) { /* (rand - 0.5) * field_scale, :283-291 */
      pos[e][0] = (u01(u[2 * e]) - 0.5f) * p->reset_scale_x;
      pos[e][1] = (u01(u[2 * e + 1]) - 0.5f) * p->reset_scale_y;
    }
    int too_close = 0; /* :293-299, 21 pairs */
    for (int a = 0; a < 7; ++a)
      for (int b = a + 1; b < 7; ++b) {
        const float dx = pos[a][0] - pos[b][0], dy = pos[a][1] - pos[b][1];
        const float d2 = fmaf(dy, dy, dx * dx);
        if (d2 < min_d2) too_close = 1;
      }
    if (!too_close) break;
  }
  uint32_t m[8];
  rng_block(seed, gid, ep, STREAM_RESET_MISC, 0, m);
  rng_block(seed, gid, ep, STREAM_RESET_MISC, 1, m + 4);
  s->ball_pos[2 * n] = pos[0][0]; /* :301 */
  s->ball_pos[2 * n + 1] = pos[0][1];
  s->ball_vel[2 * n] = (u01(m[6]) - 0.5f) * p->ball_reset_speed; /* :318-327 */
  s->ball_vel[2 * n + 1] = (u01(m[7]) - 0.5f) * p->ball_reset_speed;
  for (int r = 0; r < NB; ++r) {
    const int64_t i = n * NB + r;
    s->r_pos[2 * i] = pos[1 + r][0]; /* :302-304 */
    s->r_pos[2 * i + 1] = pos[1 + r][1];
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    const float yaw = u01(m[r]) * two_pi - pi; /* torch_rand_float(-pi, pi), :307-312 */
    s->r_rot[2 * i] = cosf(yaw);               /* quat_from_angle_axis about z, :313-315 */
    s->r_rot[2 * i + 1] = sinf(yaw);
    s->r_vel[2 * i] = s->r_vel[2 * i + 1] = 0.0f; /* template root state, :272 */
    s->r_w[i] = 0.0f;
    s->r_act[2 * i] = s->r_act[2 * i + 1] = 0.0f; /* dof_velocity_buf[env_ids] *= 0, :333 */
  }
  s->episode[n] = ep + 1;
}

ORC_API void orc_reset_dones(const vss_params* p, uint64_t seed, int64_t global_offset, orc_state* s,
                             const int64_t* reset_buf) {
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < s->n; ++n)
    if (reset_buf[n] != 0) reset_one(p, seed, (uint64_t)(global_offset + n), s, n);
}

/* ------------------------------------------------------------------------- */
/* Physics: double-precision restatement of the NEW 2-D model, DESIGN.md §3.   */
/* Replaces gym.simulate (PhysX). PARITY UNPINNED vs PhysX.                    */
/* ------------------------------------------------------------------------- */
typedef struct { double x, y, vx, vy, c, s, w, invm, invi; } body_t;
typedef struct {
  double h, H, rb, b, rw, rwc;
  double m_eff, i_eff, kd_imp, tmax, wmax, dv_max, du_max, ball_decay;
  double inv_mr, inv_ir, inv_mb, e;
  double mu_br, mu_bw, mu_rw;
  double HL, HW, GH, GD;
} phys_t;

static void phys_derive(const vss_params* p, phys_t* q) {
  q->h = (double)p->dt / (double)p->substeps;
  q->H = p->robot_half_size; q->rb = p->ball_radius; q->b = p->wheel_half_track;
  q->rw = p->wheel_radius; q->rwc = p->wheel_coll_radius;
  const double jw_lin = (double)p->wheel_inertia / (q->rw * q->rw); /* wheel inertia as linear mass */
  q->m_eff = p->robot_mass + 2.0 * jw_lin;
  q->i_eff = p->robot_inertia + 2.0 * jw_lin * q->b * q->b;
  const double j_wheel_eq = 0.5 * q->m_eff * q->rw * q->rw;
  q->kd_imp = p->drive_damping / (1.0 + q->h * p->drive_damping / j_wheel_eq);
  q->tmax = p->drive_max_torque; q->wmax = p->max_wheel_rad_s;
  q->dv_max = (double)p->mu_traction * p->gravity * q->h;
  q->du_max = (double)p->mu_lateral * p->gravity * q->h;
  q->ball_decay = exp(-(double)p->ball_drag * q->h);
  q->inv_mr = 1.0 / p->robot_mass; q->inv_ir = 1.0 / p->robot_inertia; q->inv_mb = 1.0 / p->ball_mass;
  q->e = p->restitution;
  q->mu_br = p->mu_ball_robot; q->mu_bw = p->mu_ball_wall; q->mu_rw = p->mu_robot_wall;
  q->HL = p->field_half_length; q->HW = p->field_half_width; q->GH = p->goal_half_width; q->GD = p->goal_depth;
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
static double sgn(double v) { return v < 0.0 ? -1.0 : 1.0; }

/* Inelastic contact between P and Q; unit normal n points from P to Q; cp = contact
 * point (world). kt_extra = extra tangential compliance of a rolling ball. */
static void resolve(body_t* P, body_t* Q, double nx, double ny, double depth, double cpx, double cpy,
                    double e, double mu, double kt_extra) {
  const double rpx = cpx - P->x, rpy = cpy - P->y, rqx = cpx - Q->x, rqy = cpy - Q->y;
  const double wsum = P->invm + Q->invm;
  if (wsum <= 0.0) return;
  const double wp = P->invm / wsum, wq = Q->invm / wsum;
  P->x -= nx * depth * wp; P->y -= ny * depth * wp;
  Q->x += nx * depth * wq; Q->y += ny * depth * wq;
  double vrx = (Q->vx - Q->w * rqy) - (P->vx - P->w * rpy);
  double vry = (Q->vy + Q->w * rqx) - (P->vy + P->w * rpx);
  const double vn = vrx * nx + vry * ny;
  if (vn >= 0.0) return;
  const double rnp = rpx * ny - rpy * nx, rnq = rqx * ny - rqy * nx;
  const double kn = wsum + rnp * rnp * P->invi + rnq * rnq * Q->invi;
  const double jn = -(1.0 + e) * vn / kn;
  P->vx -= jn * nx * P->invm; P->vy -= jn * ny * P->invm; P->w -= jn * rnp * P->invi;
  Q->vx += jn * nx * Q->invm; Q->vy += jn * ny * Q->invm; Q->w += jn * rnq * Q->invi;
  if (mu > 0.0) {
    const double tx = -ny, ty = nx;
    vrx = (Q->vx - Q->w * rqy) - (P->vx - P->w * rpy);
    vry = (Q->vy + Q->w * rqx) - (P->vy + P->w * rpx);
    const double vt = vrx * tx + vry * ty;
    const double rtp = rpx * ty - rpy * tx, rtq = rqx * ty - rqy * tx;
    const double kt = wsum + rtp * rtp * P->invi + rtq * rtq * Q->invi + kt_extra;
    const double jt = clampd(-vt / kt, -mu * jn, mu * jn);
    P->vx -= jt * tx * P->invm; P->vy -= jt * ty * P->invm; P->w -= jt * rtp * P->invi;
    Q->vx += jt * tx * Q->invm; Q->vy += jt * ty * Q->invm; Q->w += jt * rtq * Q->invi;
  }
}

/* Circle (centre px,py radius rho; rho = 0 -> point) against the oriented box of body B.
 * On hit returns 1 with unit normal pointing OUT of B towards the circle. */
static int circle_vs_box(const body_t* B, double H, double px, double py, double rho, double* nx,
                         double* ny, double* depth, double* cpx, double* cpy) {
  const double dx = px - B->x, dy = py - B->y;
  const double lx = dx * B->c + dy * B->s, ly = -dx * B->s + dy * B->c;
  double nlx, nly, clx, cly;
  int face = fabs(lx) <= H && fabs(ly) <= H; /* centre inside (or exactly on) the box */
  if (!face) {
    const double qx = clampd(lx, -H, H), qy = clampd(ly, -H, H);
    const double ex = lx - qx, ey = ly - qy;
    const double d2 = ex * ex + ey * ey;
    if (d2 >= rho * rho) return 0;
    if (d2 > 1e-20) {
      const double d = sqrt(d2);
      nlx = ex / d; nly = ey / d; *depth = rho - d; clx = qx; cly = qy;
    } else {
      face = 1; /* within rounding of the surface: use the face rule */
    }
  }
  if (face) { /* face of least penetration */
    const double pxd = H - fabs(lx), pyd = H - fabs(ly);
    if (pxd < pyd) { nlx = sgn(lx); nly = 0.0; *depth = pxd + rho; clx = sgn(lx) * H; cly = ly; }
    else { nlx = 0.0; nly = sgn(ly); *depth = pyd + rho; clx = lx; cly = sgn(ly) * H; }
  }
  *nx = nlx * B->c - nly * B->s; *ny = nlx * B->s + nly * B->c;
  *cpx = B->x + clx * B->c - cly * B->s; *cpy = B->y + clx * B->s + cly * B->c;
  return 1;
}

static const double CORNER[4][2] = {{1, 1}, {-1, 1}, {-1, -1}, {1, -1}};

/* Robot-robot contact, DESIGN.md §3 C: all 12 features are tested against the poses at entry
 * (snapshot A0, B0); hits are resolved in feature order with the penetration reduced by the
 * separation already gained along that normal. */
static void robot_robot(const phys_t* q, body_t* A, body_t* B) {
  const body_t A0 = *A, B0 = *B;
  for (int k = 0; k < 12; ++k) {
    /* order: corners of A in B (0-3), corners of B in A (4-7), wheels of A (8,9), wheels of B (10,11) */
    const int a_owns = (k < 8) ? (k < 4) : (k < 10);
    double lx, ly, rho;
    if (k < 8) { lx = CORNER[k & 3][0] * q->H; ly = CORNER[k & 3][1] * q->H; rho = 0.0; }
    else { lx = 0.0; ly = (k & 1) ? -q->b : q->b; rho = q->rwc; }
    const body_t* F0 = a_owns ? &A0 : &B0;
    const body_t* G0 = a_owns ? &B0 : &A0;
    body_t* F = a_owns ? A : B;
    body_t* G = a_owns ? B : A;
    const double px = F0->x + lx * F0->c - ly * F0->s, py = F0->y + lx * F0->s + ly * F0->c;
    double nx, ny, depth, cpx, cpy;
    if (!circle_vs_box(G0, q->H, px, py, rho, &nx, &ny, &depth, &cpx, &cpy)) continue;
    const double gained = ((F->x - F0->x) - (G->x - G0->x)) * nx + ((F->y - F0->y) - (G->y - G0->y)) * ny;
    const double d = depth - gained;
    resolve(G, F, nx, ny, d > 0.0 ? d : 0.0, cpx, cpy, q->e, 0.0, 0.0);
  }
}

static void ball_robot(const phys_t* q, body_t* ball, body_t* R) {
  double nx, ny, depth, cpx, cpy;
  if (circle_vs_box(R, q->H, ball->x, ball->y, q->rb, &nx, &ny, &depth, &cpx, &cpy))
    resolve(R, ball, nx, ny, depth, cpx, cpy, q->e, q->mu_br, 2.5 * ball->invm);
}

/* point/circle against the static field walls; each hit is resolved at once. */
static void body_vs_walls(const phys_t* q, body_t* Q, double lx, double ly, double rho, double mu,
                          double kt_extra) {
  body_t wall = {0};
  double nx, ny, depth;
  for (int pass = 0; pass < 3; ++pass) {
    const double px = Q->x + lx * Q->c - ly * Q->s, py = Q->y + lx * Q->s + ly * Q->c;
    const double ax = fabs(px), ay = fabs(py), sx = sgn(px), sy = sgn(py);
    int hit = 0;
    if (pass == 0) { /* side walls y = +-HW */
      if (ay > q->HW - rho) { nx = 0.0; ny = -sy; depth = ay - (q->HW - rho); hit = 1; }
    } else if (pass == 1) { /* end-wall blocks [HL,inf) x [GH,inf) per quadrant */
      if (ax >= q->HL && ay >= q->GH) {
        const double dx = ax - q->HL, dy = ay - q->GH;
        if (dx < dy) { nx = -sx; ny = 0.0; depth = dx + rho; }
        else { nx = 0.0; ny = -sy; depth = dy + rho; }
        hit = 1;
      } else {
        const double qx = ax > q->HL ? ax : q->HL, qy = ay > q->GH ? ay : q->GH;
        const double ex = ax - qx, ey = ay - qy;
        const double d2 = ex * ex + ey * ey;
        if (d2 < rho * rho && d2 > 1e-20) {
          const double d = sqrt(d2);
          nx = sx * ex / d; ny = sy * ey / d; depth = rho - d; hit = 1;
        } else if (d2 < rho * rho) { /* on the block's surface within rounding: push out along x */
          nx = -sx; ny = 0.0; depth = rho; hit = 1;
        }
      }
    } else { /* goal back wall x = +-(HL+GD) */
      if (ax > q->HL + q->GD - rho) { nx = -sx; ny = 0.0; depth = ax - (q->HL + q->GD - rho); hit = 1; }
    }
    if (hit) resolve(&wall, Q, nx, ny, depth, px - nx * rho, py - ny * rho, q->e, mu, kt_extra);
  }
}

static void robot_walls(const phys_t* q, body_t* R) {
  /* reject: the box's axis-aligned extent H(|c|+|s|) cannot reach any wall or goal post */
  const double ext = q->H * (fabs(R->c) + fabs(R->s));
  if (fabs(R->x) + ext < q->HL && fabs(R->y) + ext < q->HW) return;
  for (int k = 0; k < 4; ++k) body_vs_walls(q, R, CORNER[k][0] * q->H, CORNER[k][1] * q->H, 0.0, q->mu_rw, 0.0);
  /* goal-post corners (+-HL, +-GH) poking into a box face */
  body_t wall = {0};
  for (int k = 0; k < 4; ++k) {
    const double px = CORNER[k][0] * q->HL, py = CORNER[k][1] * q->GH;
    double nx, ny, depth, cpx, cpy;
    if (circle_vs_box(R, q->H, px, py, 0.0, &nx, &ny, &depth, &cpx, &cpy))
      resolve(R, &wall, nx, ny, depth, cpx, cpy, q->e, q->mu_rw, 0.0);
  }
}

static void substep(const phys_t* q, body_t* ball, body_t R[NB], const double act[NB][2]) {
  /* A. wheel drive + integration */
  for (int k = 0; k < NB; ++k) {
    body_t* r = &R[k];
    double v = r->vx * r->c + r->vy * r->s, u = -r->vx * r->s + r->vy * r->c;
    const double wl = (v - r->w * q->b) / q->rw, wr = (v + r->w * q->b) / q->rw;
    const double tl = clampd(q->kd_imp * (q->wmax * act[k][0] - wl), -q->tmax, q->tmax);
    const double tr = clampd(q->kd_imp * (q->wmax * act[k][1] - wr), -q->tmax, q->tmax);
    const double fl = tl / q->rw, fr = tr / q->rw;
    v += clampd((fl + fr) / q->m_eff * q->h, -q->dv_max, q->dv_max);
    r->w += q->b * (fr - fl) / q->i_eff * q->h;
    u -= clampd(u, -q->du_max, q->du_max);
    r->vx = v * r->c - u * r->s; r->vy = v * r->s + u * r->c;
    r->x += r->vx * q->h; r->y += r->vy * q->h;
    const double a = r->w * q->h, ca = cos(a), sa = sin(a);
    const double c2 = r->c * ca - r->s * sa, s2 = r->s * ca + r->c * sa;
    const double inv = 1.0 / sqrt(c2 * c2 + s2 * s2);
    r->c = c2 * inv; r->s = s2 * inv;
  }
  /* B. ball */
  ball->vx *= q->ball_decay; ball->vy *= q->ball_decay;
  ball->x += ball->vx * q->h; ball->y += ball->vy * q->h;
  /* C. pair contacts. Broadphase flags are taken ONCE, from the positions right after
   * integration; flagged pairs are then resolved in fixed order: (ball, r0..r5), then the
   * robot pairs i<j in lexicographic order. */
  const double br_reach = q->rb + q->H * 1.4142135623730951 + 0.005;
  const double rr = q->H * 1.4142135623730951, wr = q->b + q->rwc;
  const double rr_reach = 2.0 * (rr > wr ? rr : wr) + 0.01;
  int near_ball[NB], near_pair[NB][NB];
  for (int k = 0; k < NB; ++k) {
    const double dx = ball->x - R[k].x, dy = ball->y - R[k].y;
    near_ball[k] = dx * dx + dy * dy < br_reach * br_reach;
  }
  for (int i = 0; i < NB; ++i)
    for (int j = i + 1; j < NB; ++j) {
      const double dx = R[i].x - R[j].x, dy = R[i].y - R[j].y;
      near_pair[i][j] = dx * dx + dy * dy < rr_reach * rr_reach;
    }
  for (int k = 0; k < NB; ++k)
    if (near_ball[k]) ball_robot(q, ball, &R[k]);
  for (int i = 0; i < NB; ++i)
    for (int j = i + 1; j < NB; ++j)
      if (near_pair[i][j]) robot_robot(q, &R[i], &R[j]);
  /* D. robots vs walls, E. ball vs walls */
  for (int k = 0; k < NB; ++k) robot_walls(q, &R[k]);
  body_vs_walls(q, ball, 0.0, 0.0, q->rb, q->mu_bw, 2.5 * ball->invm);
}

/* one control step (dt) of ONE field; state in/out as float arrays */
static void physics_one(const vss_params* p, const phys_t* q, orc_state* s, int64_t n) {
  body_t ball = {s->ball_pos[2 * n], s->ball_pos[2 * n + 1], s->ball_vel[2 * n], s->ball_vel[2 * n + 1],
                 1.0, 0.0, 0.0, q->inv_mb, 0.0};
  body_t R[NB];
  double act[NB][2];
  for (int k = 0; k < NB; ++k) {
    const int64_t i = n * NB + k;
    R[k] = (body_t){s->r_pos[2 * i], s->r_pos[2 * i + 1], s->r_vel[2 * i], s->r_vel[2 * i + 1],
                    s->r_rot[2 * i], s->r_rot[2 * i + 1], s->r_w[i], q->inv_mr, q->inv_ir};
    act[k][0] = s->r_act[2 * i]; act[k][1] = s->r_act[2 * i + 1];
  }
  for (int it = 0; it < p->substeps; ++it) substep(q, &ball, R, act);
  s->ball_pos[2 * n] = (float)ball.x; s->ball_pos[2 * n + 1] = (float)ball.y;
  s->ball_vel[2 * n] = (float)ball.vx; s->ball_vel[2 * n + 1] = (float)ball.vy;
  for (int k = 0; k < NB; ++k) {
    const int64_t i = n * NB + k;
    s->r_pos[2 * i] = (float)R[k].x; s->r_pos[2 * i + 1] = (float)R[k].y;
    s->r_vel[2 * i] = (float)R[k].vx; s->r_vel[2 * i + 1] = (float)R[k].vy;
    s->r_rot[2 * i] = (float)R[k].c; s->r_rot[2 * i + 1] = (float)R[k].s;
    s->r_w[i] = (float)R[k].w;
  }
}

ORC_API void orc_physics(const vss_params* p, orc_state* s) {
  phys_t q;
  phys_derive(p, &q);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < s->n; ++n) physics_one(p, &q, s, n);
}

/* ------------------------------------------------------------------------- */
/* VSS.step = VecTask.step -> pre / simulate / post (envs/vss.py:180-203)      */
/*   post_state != NULL: skip physics, take the post-physics state from it     */
/*   (58 x ld SoA as in include/vss_b200.h) — mirrors vss_step_injected.       */
/* ------------------------------------------------------------------------- */
static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void obs_one(const orc_state* s, int64_t n, float* obs_row) {
  obs_field(0, s->ball_pos + 2 * n, s->ball_vel + 2 * n, s->r_pos + 2 * NB * n, s->r_vel + 2 * NB * n,
            s->r_rot + 2 * NB * n, s->r_w + NB * n, s->r_act + 2 * NB * n, obs_row);
}

static void load_injected(orc_state* s, int64_t n, const float* post, int64_t ld) {
  s->ball_pos[2 * n] = post[0 * ld + n]; s->ball_pos[2 * n + 1] = post[1 * ld + n];
  s->ball_vel[2 * n] = post[2 * ld + n]; s->ball_vel[2 * n + 1] = post[3 * ld + n];
  for (int k = 0; k < NB; ++k) {
    const int64_t i = n * NB + k;
    const float* w = post + (4 + 9 * k) * ld + n;
    s->r_pos[2 * i] = w[0 * ld]; s->r_pos[2 * i + 1] = w[1 * ld];
    s->r_vel[2 * i] = w[2 * ld]; s->r_vel[2 * i + 1] = w[3 * ld];
    s->r_rot[2 * i] = w[4 * ld]; s->r_rot[2 * i + 1] = w[5 * ld];
    s->r_w[i] = w[6 * ld];
  }
}

ORC_API void orc_step(const vss_params* p, uint64_t seed, int64_t global_offset, orc_state* s,
                      const float* actions /* (N,2,3,2) */, const float* post_state, int64_t post_ld,
                      int64_t* reset_buf, float* obs, float* term_obs, float* rew, uint8_t* timeout,
                      float* progress_f) {
  phys_t q;
  phys_derive(p, &q);
  const float field_width = 2.0f * p->field_half_length, goal_height = 2.0f * p->goal_half_width;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < s->n; ++n) {
    /* VecTask.step: clamp actions to +-clipActions (vss.yaml:7) */
    float act[NB][2];
    for (int k = 0; k < NB; ++k)
      for (int c = 0; c < 2; ++c) act[k][c] = clampf(actions[(n * NB + k) * 2 + c], -1.0f, 1.0f);
    /* pre_physics_step, vss.py:180-187 */
    if (reset_buf[n] != 0) s->progress[n] = 0;
    for (int k = 0; k < NB; ++k)
      for (int c = 0; c < 2; ++c) s->r_act[(n * NB + k) * 2 + c] = act[k][c];
    /* prev_* clones taken before the refresh, vss.py:219-220 */
    float prev_ball[2] = {s->ball_pos[2 * n], s->ball_pos[2 * n + 1]};
    float prev_rpos[NB][2];
    memcpy(prev_rpos, s->r_pos + 2 * NB * n, sizeof(prev_rpos));
    /* gym.simulate */
    if (post_state) load_injected(s, n, post_state, post_ld);
    else physics_one(p, &q, s, n);
    /* post_physics_step, vss.py:189-203 */
    s->progress[n] += 1;
    float rw[NB][4];
    /* safety net (engine-specific, not in the reference): a field whose state is not finite is
     * re-randomised on the spot and reported as done with zero reward and no timeout */
    int finite = isfinite(s->ball_pos[2 * n]) && isfinite(s->ball_pos[2 * n + 1]) &&
                 isfinite(s->ball_vel[2 * n]) && isfinite(s->ball_vel[2 * n + 1]);
    for (int k = 0; k < NB; ++k) {
      const int64_t i = n * NB + k;
      finite = finite && isfinite(s->r_pos[2 * i]) && isfinite(s->r_pos[2 * i + 1]) && isfinite(s->r_vel[2 * i]) &&
               isfinite(s->r_vel[2 * i + 1]) && isfinite(s->r_rot[2 * i]) && isfinite(s->r_rot[2 * i + 1]) &&
               isfinite(s->r_w[i]);
    }
    if (finite) {
      rewards_one(p, prev_ball, prev_rpos, s->ball_pos + 2 * n, (const float(*)[2])(s->r_pos + 2 * NB * n),
                  (const float(*)[2])act, rw);
    } else {
      reset_one(p, seed, (uint64_t)(global_offset + n), s, n);
      memset(rw, 0, sizeof(rw));
    }
    memcpy(rew + n * VSS_REW_PER_FIELD, rw, sizeof(rw));
    orc_compute_dones(1, s->ball_pos + 2 * n, s->progress + n, (float)p->max_episode_length, field_width,
                      goal_height, reset_buf + n);
    if (!finite) reset_buf[n] = 1;
    if (term_obs) obs_one(s, n, term_obs + n * VSS_OBS_PER_FIELD);            /* :195-196 */
    if (progress_f) progress_f[n] = (float)s->progress[n];                    /* :198-200 */
    if (reset_buf[n] != 0 && finite) reset_one(p, seed, (uint64_t)(global_offset + n), s, n); /* :202 */
    obs_one(s, n, obs + n * VSS_OBS_PER_FIELD);                               /* :203 */
    /* VecTask.step: timeout_buf = (progress_buf >= max_len - 1) & (reset_buf != 0) */
    timeout[n] = (uint8_t)(finite && (s->progress[n] >= p->max_episode_length - 1) && (reset_buf[n] != 0));
  }
}

/* ------------------------------------------------------------------------- */
/* random_ou — envs/wrappers.py:5-19, RNG = Philox stream 2 keyed by step      */
/* ------------------------------------------------------------------------- */
static void ou_one(const vss_params* p, uint64_t seed, uint64_t gid, uint32_t step, float* abuf /* 12 */) {
  const float two_pi = 6.283185307179586f;
  for (uint32_t b = 0; b < 3; ++b) {
    uint32_t u[4];
    rng_block(seed, gid, step, STREAM_OU, b, u);
    for (int h = 0; h < 2; ++h) { /* Box-Muller: two normals per uniform pair */
      const float rad = sqrtf(-2.0f * logf(u01_open(u[2 * h])));
      const float ang = two_pi * u01(u[2 * h + 1]);
      const float z[2] = {rad * cosf(ang), rad * sinf(ang)};
      for (int c = 0; c < 2; ++c) {
        float* a = abuf + 4 * b + 2 * h + c;
        const float v = *a - p->ou_theta * *a + p->ou_sigma * z[c];
        *a = clampf(v, -1.0f, 1.0f);
      }
    }
  }
}

/* View outputs of SingleAgent / CMA / DMA .step + RecordEpisodeStatisticsTorch.step given the
 * raw VSS.step outputs — envs/wrappers.py:104-115, 136-148, 166-180, 68-81 */
ORC_API void orc_view_outputs(int64_t N, int view, const float* obs, const float* tobs, const float* rew,
                              const int64_t* reset_buf, const uint8_t* tout, const float* prog,
                              float* action_buf, float* obs_v, float* term_obs_v, float* rews_v,
                              float* reward_v, int64_t* done_v, uint8_t* timeout_v, float* progress_v,
                              float* ep_ret, int32_t* ep_len, float* ret_ret, int32_t* ret_len) {
  const int per = (view == VSS_VIEW_DMA) ? 3 : 1;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    /* action_buf[env_ids] *= 0 for done envs, wrappers.py:105-107 */
    if (reset_buf[n] != 0)
      for (int c = 0; c < VSS_ACT_PER_FIELD; ++c) action_buf[n * VSS_ACT_PER_FIELD + c] *= 0.0f;
    for (int j = 0; j < per; ++j) {
      const int64_t v = n * per + j;
      memcpy(obs_v + v * VSS_NUM_OBS, obs + n * VSS_OBS_PER_FIELD + j * VSS_NUM_OBS, sizeof(float) * VSS_NUM_OBS);
      memcpy(term_obs_v + v * VSS_NUM_OBS, tobs + n * VSS_OBS_PER_FIELD + j * VSS_NUM_OBS, sizeof(float) * VSS_NUM_OBS);
      float r4[4];
      if (view == VSS_VIEW_CMA) { /* rewards[:, 0, :].mean(1), wrappers.py:140 */
        for (int c = 0; c < 4; ++c) {
          const float* r = rew + n * VSS_REW_PER_FIELD;
          r4[c] = ((r[c] + r[4 + c]) + r[8 + c]) / 3.0f;
        }
      } else { /* rewards[:, 0, 0] (sa) / rewards[:, 0, :].reshape(-1, 4) (dma) */
        memcpy(r4, rew + n * VSS_REW_PER_FIELD + j * 4, sizeof(r4));
      }
      memcpy(rews_v + 4 * v, r4, sizeof(r4));
      reward_v[v] = ((r4[0] + r4[1]) + r4[2]) + r4[3]; /* infos['rews'].sum(-1) */
      done_v[v] = reset_buf[n];
      timeout_v[v] = tout[n];
      progress_v[v] = prog[n];
      if (ep_ret) { /* RecordEpisodeStatisticsTorch.step, wrappers.py:68-75 */
        for (int c = 0; c < 4; ++c) {
          ep_ret[4 * v + c] += r4[c];
          ret_ret[4 * v + c] = ep_ret[4 * v + c];
          ep_ret[4 * v + c] *= (float)(1 - done_v[v]);
        }
        ep_len[v] += 1;
        ret_len[v] = ep_len[v];
        ep_len[v] *= (int32_t)(1 - done_v[v]);
      }
    }
  }
}

/* SingleAgent / CMA / DMA .step — envs/wrappers.py:101-115, 133-148, 163-180 */
ORC_API void orc_step_view(const vss_params* p, uint64_t seed, int64_t global_offset, uint32_t step_index,
                           orc_state* s, int view, const float* policy_action, float* action_buf,
                           int64_t* reset_buf, float* obs_v, float* term_obs_v, float* rews_v,
                           float* reward_v, int64_t* done_v, uint8_t* timeout_v, float* progress_v,
                           float* ep_ret, int32_t* ep_len, float* ret_ret, int32_t* ret_len) {
  const int64_t N = s->n;
  /* the raw task's persistent buffers (obs_buf, rew_buf, ... of envs/vss.py:75-97): kept between calls, as
   * the reference keeps them, so that a timed run does not pay 2.6 KB per field of page faults per step */
  static float *obs = NULL, *tobs = NULL, *rew = NULL, *prog = NULL;
  static uint8_t* tout = NULL;
  static int64_t cap = 0;
  if (N > cap) {
    free(obs); free(tobs); free(rew); free(tout); free(prog);
    obs = (float*)malloc(sizeof(float) * N * VSS_OBS_PER_FIELD);
    tobs = (float*)malloc(sizeof(float) * N * VSS_OBS_PER_FIELD);
    rew = (float*)malloc(sizeof(float) * N * VSS_REW_PER_FIELD);
    tout = (uint8_t*)malloc(N);
    prog = (float*)malloc(sizeof(float) * N);
    cap = N;
  }
  /* action_buf = random_ou(action_buf); act_view[:] = action  (wrappers.py:102-103) */
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    float* a = action_buf + n * VSS_ACT_PER_FIELD;
    ou_one(p, seed, (uint64_t)(global_offset + n), step_index, a);
    if (view == VSS_VIEW_SA) { a[0] = policy_action[2 * n]; a[1] = policy_action[2 * n + 1]; }
    else for (int c = 0; c < 6; ++c) a[c] = policy_action[6 * n + c]; /* cma (N,6) == dma (3N,2) */
  }
  orc_step(p, seed, global_offset, s, action_buf, NULL, 0, reset_buf, obs, tobs, rew, tout, prog);
  orc_view_outputs(N, view, obs, tobs, rew, reset_buf, tout, prog, action_buf, obs_v, term_obs_v, rews_v,
                   reward_v, done_v, timeout_v, progress_v, ep_ret, ep_len, ret_ret, ret_len);
}

/* ------------------------------------------------------------------------- */
/* GAE — ppo_continuous_action_isaacgym.py:282-296                             */
/* ------------------------------------------------------------------------- */
ORC_API void orc_gae(const float* rewards, const float* values, const float* next_values,
                     const float* next_dones, const float* next_timeouts, float* advantages,
                     float* returns, int32_t T, int64_t N, double gamma_d, double gae_lambda_d) {
  /* python scalars are doubles; torch rounds each to float32 when it meets a tensor:
   * `args.gamma * next_values[t]` and `args.gamma * args.gae_lambda * (1.0 - ...)` */
  const float gamma = (float)gamma_d, gamma_lambda = (float)(gamma_d * gae_lambda_d);
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    float lastgaelam = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
      const int64_t i = (int64_t)t * N + n;
      /* next_non_terminal = 1 - (done & !timeout), :286 */
      const float nnt = 1.0f - (float)((next_dones[i] != 0.0f) && !(next_timeouts[i] != 0.0f));
      const float delta = rewards[i] + gamma * next_values[i] * nnt - values[i];
      lastgaelam = delta + gamma_lambda * (1.0f - next_dones[i]) * lastgaelam;
      advantages[i] = lastgaelam;
      returns[i] = lastgaelam + values[i];
    }
  }
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
