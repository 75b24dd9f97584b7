// vss_emu.cpp — TEST HARNESS: runs the product's per-lane step code (csrc/vss_lane.cuh,
// the exact source the CUDA kernel compiles) on the host, one 32-field tile at a time, so the
// step logic can be checked against the oracle in the GPU-less build container. The
// orchestration below mirrors k_step / k_reset_dones in csrc/vss_step.cu (phases 1-5).
// Not part of the product: libvss_b200.so has no host execution path.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../rsoccer_isaac_cleanrl_b200/csrc/vss_lane.cuh"

using namespace vss;

static const ObsTable g_tab = make_obs_table();

template <int VIEW, bool INJECT>
static void emu_step_t(const StepArgs& a, const DevParams& P) {
  constexpr int PER_FIELD = ViewShape<VIEW>::F4_PER;
  const long long tiles = (a.n + 31) / 32;
#pragma omp parallel for schedule(static)
  for (long long tile = 0; tile < tiles; ++tile) {
    std::vector<float> Tbuf(TILE_WORDS, 0.0f);
    float* T = Tbuf.data();
    const long long env0 = tile * 32;
    const int valid = (int)std::min(32LL, a.n - env0);
    bool done[32] = {false}, ended[32] = {false};
    uint32_t done_mask = 0;
    for (int lane = 0; lane < valid; ++lane) lane_phase1a<VIEW>(T + lane, env0 + lane, a, P, make_key(a, env0 + lane));
    if (INJECT) {
      for (int lane = 0; lane < valid; ++lane) lane_inject(T + lane, env0 + lane, a);
    } else {  // physics_tile of vss_step.cu: wall tasks compacted over the warp
      for (int it = 0; it < P.substeps; ++it) {
        std::vector<int> queue;
        for (int lane = 0; lane < valid; ++lane) {
          uint32_t m = substep_pre_lane(T + lane, P);
          for (int r = 0; r < 6; ++r) if (m & (1u << r)) queue.push_back((lane << 3) | r);
        }
        for (int q : queue) robot_walls_task(T + (q >> 3), q & 7, P);
        for (int lane = 0; lane < valid; ++lane) substep_ball_walls_lane(T + lane, P);
      }
    }
    for (int lane = 0; lane < valid; ++lane) {
      const int code = lane_phase1d<VIEW>(T + lane, env0 + lane, a, P, make_key(a, env0 + lane));
      done[lane] = code == LANE_DONE;
      ended[lane] = code != LANE_RUNNING;
      if (done[lane]) done_mask |= 1u << lane;
    }
    float* ob = a.obs + env0 * (PER_FIELD * 4);
    float* tob = a.term_obs ? a.term_obs + env0 * (PER_FIELD * 4) : nullptr;
    void* pk = (VIEW != VIEW_FULL && a.packed)
                   ? static_cast<void*>(static_cast<char*>(a.packed) + env0 * (ViewShape<VIEW>::AGENTS * VSS_PACKED_ROW_BYTES))
                   : nullptr;
    for (int lane = 0; lane < 32; ++lane) {
      if (PER_FIELD >= 32) write_obs_tile_rows<PER_FIELD>(T, g_tab.v, lane, valid, tob, ob, done_mask, nullptr, pk);
      else write_obs_tile(T, g_tab.v, lane, valid, PER_FIELD, tob, ob, done_mask, nullptr, pk);
    }
    for (int lane = 0; lane < valid; ++lane)
      if (done[lane]) reset_lane(T + lane, P, make_key(a, env0 + lane));
    for (int lane = 0; lane < 32; ++lane) write_obs_fields(T, g_tab.v, lane, PER_FIELD, ob, done_mask, nullptr, pk);
    for (int lane = 0; lane < valid; ++lane) lane_phase5<VIEW>(T + lane, env0 + lane, a, ended[lane]);
  }
}

// Mirror of k_step_cta (one tile shared by NW warps: W = NW - 1 body warps with the bodies dealt to them,
// and a helper warp that computes the resets speculatively into shadow columns): the phases below are the
// kernel's, each CTA barrier becomes the end of a loop over (warp, lane).
template <int VIEW, bool INJECT>
static void emu_step_cta_t(const StepArgs& a, const DevParams& P, int NW, int F) {
  constexpr int PER_FIELD = ViewShape<VIEW>::F4_PER;
  const int W = NW - 1;
  const long long tiles = (a.n + F - 1) / F;  // F = 32, 16 or 8 fields per tile (vss_set_step_fields_per_tile)
#pragma omp parallel for schedule(static)
  for (long long tile = 0; tile < tiles; ++tile) {
    std::vector<float> Tbuf(TILE_CTA_WORDS, 0.0f);
    float* T = Tbuf.data();
    const long long env0 = tile * F;
    const int valid = (int)std::min((long long)F, a.n - env0);
#define EACH_THREAD for (int warp = 0; warp < NW; ++warp) for (int lane = 0; lane < valid; ++lane)
#define EACH_BODY_THREAD for (int warp = 0; warp < W; ++warp) for (int lane = 0; lane < valid; ++lane)
    EACH_THREAD load_state_words(T + lane, a.state, a.ld, env0 + lane, warp, NW);
    EACH_BODY_THREAD {
      float* S = T + lane;
      const long long env = env0 + lane;
      const RngKey key = make_key(a, env);
      for (int b = warp; b < 3; b += W) actions_block<VIEW>(S, env, a, P, key, b);
      if (warp == 3 % W && a.reset_buf[env] != 0) S[VSS_W_PROGRESS * LDS] = bitsf(0u);
      for (int b = warp; b < 7; b += W) prev_term_body(S, b, P);
    }
    for (int lane = 0; lane < valid; ++lane) {  // the helper warp
      float* S = T + lane;
      float* Sh = S + W_SHADOW * LDS;
      Sh[VSS_W_EPISODE * LDS] = S[VSS_W_EPISODE * LDS];
      reset_lane(Sh, P, make_key(a, env0 + lane));
    }
    if (INJECT) {
      for (int lane = 0; lane < valid; ++lane) lane_inject(T + lane, env0 + lane, a);
    } else {
      for (int it = 0; it < P.substeps; ++it) {
        EACH_BODY_THREAD for (int b = warp; b < 7; b += W) { if (b < 6) integrate_robot(T + lane, b, P); else integrate_ball(T + lane, P); }
        EACH_BODY_THREAD {
          uint32_t m = 0u;
          for (int q = warp; q < 21; q += W) m |= broadphase_pair(T + lane, q, P);
          T[(W_SCR + warp) * LDS + lane] = bitsf(m);
        }
        EACH_BODY_THREAD if (lane % W == warp) {
          uint32_t m = 0u;
          for (int j = 0; j < W; ++j) m |= fbits(T[(W_SCR + j) * LDS + lane]);
          if (m) contacts_task(T + lane, m, P);
        }
        EACH_BODY_THREAD for (int b = warp; b < 7; b += W) walls_body(T + lane, b, P);
      }
    }
    bool finite[32];
    for (int lane = 0; lane < valid; ++lane) {
      finite[lane] = state_finite(T + lane);
      if (!finite[lane]) lane_sanitise(T + lane, a, P, make_key(a, env0 + lane));
    }
    float* ob = a.obs + env0 * (PER_FIELD * 4);
    float* tob = a.term_obs ? a.term_obs + env0 * (PER_FIELD * 4) : nullptr;
    void* pk = (VIEW != VIEW_FULL && a.packed)
                   ? static_cast<void*>(static_cast<char*>(a.packed) + env0 * (ViewShape<VIEW>::AGENTS * VSS_PACKED_ROW_BYTES))
                   : nullptr;
    // warps 1 .. NW-1 write the rows of ALL fields (before warp 0's outputs are known) ...
    for (int warp = 1; warp < NW; ++warp)
      for (int lane = 0; lane < 32; ++lane) {
        if (PER_FIELD >= 32) write_obs_tile_rows<PER_FIELD, true>(T, g_tab.v, lane, valid, tob, ob, 0u, nullptr, pk, warp - 1, NW - 1);
        else write_obs_tile(T, g_tab.v, lane, valid, PER_FIELD, tob, ob, 0u, nullptr, pk, warp - 1, NW - 1);
      }
    // ... while warp 0 computes them
    uint32_t done_mask = 0, ended_mask = 0;
    for (int lane = 0; lane < valid; ++lane) {
      const int code = lane_outputs<VIEW>(T + lane, env0 + lane, a, P, finite[lane]);
      if (code == LANE_DONE) done_mask |= 1u << lane;
      if (code != LANE_RUNNING) ended_mask |= 1u << lane;
    }
    EACH_THREAD if ((done_mask >> lane) & 1u)
      for (int w = warp; w < VSS_STATE_WORDS; w += NW)
        if (w != VSS_W_PROGRESS) T[w * LDS + lane] = T[(W_SHADOW + w) * LDS + lane];
    for (int warp = 0; warp < NW; ++warp)
      for (int lane = 0; lane < 32; ++lane) write_obs_fields(T, g_tab.v, lane, PER_FIELD, ob, done_mask, nullptr, pk, warp, NW);
    EACH_THREAD store_state_words(T + lane, a.state, a.ld, env0 + lane, warp, NW);
    if (VIEW != VIEW_FULL)
      for (int lane = 0; lane < valid; ++lane)
        if ((ended_mask >> lane) & 1u) zero_action_row(env0 + lane, a);
#undef EACH_THREAD
#undef EACH_BODY_THREAD
  }
}

extern "C" {

__attribute__((visibility("default"))) int emu_step(
    const vss_params* p, int view, float* state, long long n, long long ld, unsigned long long goff,
    unsigned long long seed, unsigned int step, const float* actions, const float* inject, long long* reset_buf,
    float* obs, float* term_obs, float* rew, uint8_t* timeout, float* progress_f, const float* policy_action,
    float* action_buf, float* reward_v, long long* done_v, float* ep_ret, int* ep_len, float* ret_ret,
    int* ret_len, void* packed, int wpt, int fpt) {
  const DevParams P = derive_params(*p);
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.state = state; a.n = n; a.ld = ld; a.goff = goff;
  a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); unsigned long long step_ctr[3] = {step, 0, 0};  // index, CTA count, sanitised fields
  a.step_ctr = step_ctr; a.grid = 1;
  a.actions = actions; a.inject = inject; a.reset_buf = reset_buf; a.obs = obs; a.term_obs = term_obs;
  a.rew = rew; a.timeout = timeout; a.progress_f = progress_f; a.policy_action = policy_action;
  a.action_buf = action_buf; a.reward_v = reward_v; a.done_v = done_v; a.ep_ret = ep_ret; a.ep_len = ep_len;
  a.ret_ret = ret_ret; a.ret_len = ret_len; a.packed = packed;
  if (wpt > 1) {  // the k_step_cta structure
    if (wpt > MAX_WPT || (fpt != 8 && fpt != 16 && fpt != 32)) return -1;
    if (view == VIEW_FULL) { if (inject) emu_step_cta_t<VIEW_FULL, true>(a, P, wpt, fpt); else emu_step_cta_t<VIEW_FULL, false>(a, P, wpt, fpt); }
    else if (view == VSS_VIEW_SA) emu_step_cta_t<VSS_VIEW_SA, false>(a, P, wpt, fpt);
    else if (view == VSS_VIEW_CMA) emu_step_cta_t<VSS_VIEW_CMA, false>(a, P, wpt, fpt);
    else if (view == VSS_VIEW_DMA) emu_step_cta_t<VSS_VIEW_DMA, false>(a, P, wpt, fpt);
    else return -1;
    return 0;
  }
  if (view == VIEW_FULL) { if (inject) emu_step_t<VIEW_FULL, true>(a, P); else emu_step_t<VIEW_FULL, false>(a, P); }
  else if (view == VSS_VIEW_SA) emu_step_t<VSS_VIEW_SA, false>(a, P);
  else if (view == VSS_VIEW_CMA) emu_step_t<VSS_VIEW_CMA, false>(a, P);
  else if (view == VSS_VIEW_DMA) emu_step_t<VSS_VIEW_DMA, false>(a, P);
  else return -1;
  return 0;
}

__attribute__((visibility("default"))) int emu_reset_dones(const vss_params* p, float* state, long long n,
                                                           long long ld, unsigned long long goff,
                                                           unsigned long long seed, const long long* reset_buf,
                                                           float* obs) {
  const DevParams P = derive_params(*p);
  StepArgs a;
  memset(&a, 0, sizeof(a));
  a.goff = goff; a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32);
  const long long tiles = (n + 31) / 32;
  for (long long tile = 0; tile < tiles; ++tile) {
    std::vector<float> Tbuf(TILE_WORDS, 0.0f);
    float* T = Tbuf.data();
    const long long env0 = tile * 32;
    const int valid = (int)std::min(32LL, n - env0);
    for (int lane = 0; lane < valid; ++lane) {
      load_state(T + lane, state, ld, env0 + lane);
      if (reset_buf[env0 + lane] != 0) {
        reset_lane(T + lane, P, make_key(a, env0 + lane));
        store_state(T + lane, state, ld, env0 + lane);
      }
    }
    if (obs)
      for (int lane = 0; lane < 32; ++lane)
        write_obs_tile(T, g_tab.v, lane, valid, F4_PER_FIELD, nullptr, obs + env0 * VSS_OBS_PER_FIELD, 0u);
  }
  return 0;
}

}  // extern "C"
