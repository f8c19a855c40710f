// Fused one-step model rollout on a CTA PAIR (cluster 2x1x1, tcgen05 cta_group::2), two row tiles in flight.
//
// Same math and the same per-layer schedule as step_tc.cu (reference
// algo/dynamics/mobody_dynamics.py:193-265, algo/dynamics/mobody_module.py:217-330,
// algo/offline_offline/mobody.py:60-72).  What changes is the shape of the machine it runs on:
//
//  * step_tc.cu walks ONE 128-row tile per SM through a strictly serial chain: a 256-wide epilogue cannot start
//    before its MMA finished, and the next wide MMA cannot overwrite the single 128 KB (bf16 hi+lo) A operand any
//    earlier, so the tensor pipe and the SFU/ALU pipes mostly take turns (ncu: 35 % / 38 % busy).
//  * here a CTA pair owns TWO 128-row tiles (X and Y).  Every MMA is M = 128 across the pair (cta_group::2): each
//    CTA holds 64 rows of a tile's A operand (64 KB with both planes — two tiles fit), stages HALF of each weight
//    K step (N/2 rows of B), and its TMEM holds 64 rows x N as 128 lanes x N/2 columns ("2x2" layout: lanes 0-63
//    = columns [0, N/2), lanes 64-127 = columns [N/2, N)), so four 128-column accumulators (2 tiles x ping-pong)
//    fill the 512 columns.  The two tiles are staggered by one layer: while the epilogue warps work on tile X's
//    layer l, the tensor pipe runs tile Y's layer l, then X's layer l+1 trailing X's epilogue chunk by chunk.
//
// Roles per CTA (18 warps): 16 epilogue warps (warp & 3 = TMEM lane quadrant, warp >> 2 = 8-column group),
// one producer warp (weights: its half of every K step, L2 -> SMEM ring with 1-D TMA bulk copies; biases), one MMA
// warp (leader CTA: issues tcgen05.mma.cta_group::2 for the pair; peer CTA: relays "my half landed" to the leader).
// Barriers: every barrier an epilogue warp touches is LOCAL to its CTA (a cluster-scope release per warp and chunk
// costs more than the chunk's math); the leader's commits are multicast to both CTAs (w_empty, d_full); a relay warp
// in the peer CTA forwards the completion of the peer's local a_ready / d_empty barriers to the leader's a_peer /
// d_peer barriers with ONE cluster-scope arrive each (and the peer's MMA warp forwards w_full -> w_peer).
#include "common.cuh"
#include "philox.cuh"
#include "term.cuh"
#include "tc_prims.cuh"
#include "tc_layout.h"
#include "tc_epi.cuh"
#include <stdlib.h>

namespace tcp {
using namespace tce;

constexpr int RH = 64;                     // rows of a tile held by one CTA
constexpr uint32_t MAIN_PLANE = 32768;     // 64 rows x 256 k x bf16
constexpr uint32_t KG_BYTES = RH * 16;     // one k-group (8 k) of 64 rows
constexpr int MAX_NST = 12;
constexpr int NB = 4;                      // bias ring slots (one layer's bias each, shared by both tiles)
constexpr int BSLOT = 528;                 // 256 bias + reward_model3 vector (272)
constexpr int EPI_WARPS = 16, PROD_WARP = 16, MMA_WARP = 17, RELAY_WARP = 18, NTHREADS = 32 * 19;

struct Cfg {
  int nst;
  uint32_t stage_bytes, small_plane, sa_off, obs_kp, sas_kp;
  uint32_t dyn_bias_base, pol_bias_base;
  int has_policy, first_dyn;
  long long* trace;     // debug: clock64 stamps of pair 0's leader CTA, [(layer * 2 + tile)][8]; nullptr in production
};

struct Bars {
  uint64_t w_full[MAX_NST], w_empty[MAX_NST], w_peer[MAX_NST];
  uint64_t a_ready[2][4], d_full[2][2], d_empty[2][2], b_full[NB], b_empty[NB];   // local: this CTA's 16 epilogue warps
  uint64_t a_peer[2][4], d_peer[2][2];                                            // leader only: relayed from the peer CTA
  uint32_t tmem_slot, pad;
};

// K order of a 256-deep layer: epilogue chunk c (32 TMEM columns) produces k-groups {4c..4c+3} (lanes 0-63) and
// {16+4c..16+4c+3} (lanes 64-127), i.e. K steps {2c, 2c+1, 8+2c, 9+2c}; MMAs and weight stages follow that order.
__device__ __forceinline__ int kperm(int i) { return (i & 1) + 2 * (i >> 2) + 8 * ((i >> 1) & 1); }

template <int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
step_pair_kernel(const StepArgs a, const unsigned char* __restrict__ dynb, const unsigned char* __restrict__ polb,
                 const __grid_constant__ TcSched sched, const Cfg cfg) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int S = a.S, A = a.A, B = a.B;
  const int live = a.n_rows_dev ? min(*a.n_rows_dev, B) : B;
  const int pair_row0 = (int)(blockIdx.x >> 1) * 256;
  if (pair_row0 >= live) return;                              // the whole pair is past the live rows
  const uint32_t rank = tc::cluster_ctarank();                // == blockIdx.x & 1

  unsigned char* A_main = smem;                                        // [tile][plane][kgroup 0..31][64 rows][8 k]
  unsigned char* A_small = A_main + 2 * NS * MAIN_PLANE;               // [tile][plane][small_plane]
  unsigned char* wst = A_small + 2 * NS * cfg.small_plane;
  float* red = reinterpret_cast<float*>(wst + (size_t)cfg.nst * cfg.stage_bytes);      // [2][8][64]
  float* bias_s = red + 2 * 8 * RH;                                                    // [NB][BSLOT]
  Bars* bars = reinterpret_cast<Bars*>(bias_s + NB * BSLOT);
  auto main_of = [&](int u) { return A_main + (size_t)u * NS * MAIN_PLANE; };
  auto small_of = [&](int u) { return A_small + (size_t)u * NS * cfg.small_plane; };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < cfg.nst; ++i) { tc::mbar_init(&bars->w_full[i], 1); tc::mbar_init(&bars->w_empty[i], 1); tc::mbar_init(&bars->w_peer[i], 1); }
    for (int u = 0; u < 2; ++u) {
      for (int i = 0; i < 4; ++i) { tc::mbar_init(&bars->a_ready[u][i], EPI_WARPS); tc::mbar_init(&bars->a_peer[u][i], 1); }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&bars->d_full[u][i], 1); tc::mbar_init(&bars->d_empty[u][i], EPI_WARPS); tc::mbar_init(&bars->d_peer[u][i], 1);
      }
    }
    for (int i = 0; i < NB; ++i) { tc::mbar_init(&bars->b_full[i], 1); tc::mbar_init(&bars->b_empty[i], 2 * EPI_WARPS); }   // 2 tiles x 16 warps
    tc::mbar_fence_init();
  }
  if (warp == PROD_WARP) tc::tmem_alloc2(&bars->tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == PROD_WARP) {
    // ================= producer: this CTA's half of every weight K step, once per (layer, tile) unit =================
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int li = 0; li < sched.n_layers; ++li) {
        const TcLayer L = sched.L[li];
        const unsigned char* src = (L.blob ? polb : dynb) + L.w_off;
        const uint32_t nh16 = (uint32_t)(L.n >> 1) * 16u;                 // bytes of one (plane, k-group) piece of my half
        const uint32_t step_bytes = (uint32_t)NS * L.n * 32u;             // one K step of all planes in the packed image
        {   // this layer's bias (both tiles use it) -> bias ring slot li % NB
          const int slot = li & (NB - 1);
          const float* bsrc = reinterpret_cast<const float*>(L.blob ? polb + cfg.pol_bias_base : dynb + cfg.dyn_bias_base) + L.b_off;
          const uint32_t bb = (L.kind == EPI_REWARD ? (uint32_t)BSLOT : (uint32_t)L.n) * 4u;
          tc::mbar_wait(&bars->b_empty[slot], (uint32_t)(((li / NB) & 1) ^ 1));
          tc::mbar_arrive_expect_tx(&bars->b_full[slot], bb);
          tc::bulk_g2s(bias_s + slot * BSLOT, bsrc, bb, &bars->b_full[slot]);
        }
        const bool perm = (L.ksteps == 16 && L.a_region == REG_MAIN);
        for (int u = 0; u < 2; ++u)
          for (int i = 0; i < L.ksteps; ++i) {
            const int s = perm ? kperm(i) : i;
            tc::mbar_wait(&bars->w_empty[stage], phase ^ 1u);
            const uint32_t half_bytes = (uint32_t)NS * 2u * nh16;             // packed image: [kstep][half][plane][kgroup][n_local][8 k]
            tc::mbar_arrive_expect_tx(&bars->w_full[stage], half_bytes);
            tc::bulk_g2s(wst + (size_t)stage * cfg.stage_bytes, src + (size_t)s * step_bytes + (size_t)rank * half_bytes, half_bytes,
                         &bars->w_full[stage]);
            if (++stage == cfg.nst) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == MMA_WARP) {
    if (rank != 0) {
      // ================= peer CTA: relay "my half of this stage landed" to the leader =================
      if (tc::elect_one()) {
        int stage = 0; uint32_t wphase = 0;
        for (int li = 0; li < sched.n_layers; ++li)
          for (int i = 0; i < 2 * sched.L[li].ksteps; ++i) {
            tc::mbar_wait(&bars->w_full[stage], wphase);
            tc::mbar_arrive_remote(&bars->w_peer[stage], 0);
            if (++stage == cfg.nst) { stage = 0; wphase ^= 1u; }
          }
      }
    } else if (tc::elect_one()) {
      // ================= leader CTA: MMA issuer for the pair, units in (layer, tile) order =================
      int stage = 0; uint32_t wphase = 0, aph[2] = {0u, 0u};
      const uint32_t wst0 = tc::smem_u32(wst);
      for (int li = 0; li < sched.n_layers; ++li) {
        const TcLayer L = sched.L[li];
        const int buf = li & 1;
        const uint32_t idesc = tc::make_idesc_bf16(128, L.n);                  // M = 128 over the CTA pair
        const uint32_t nh16 = (uint32_t)(L.n >> 1) * 16u;
        const uint32_t bplane = 2u * nh16, blbo = nh16;                        // my half of B: [plane][kgroup][n_local][8 k]
        const bool perm = (L.ksteps == 16 && L.a_region == REG_MAIN);
        for (int u = 0; u < 2; ++u) {
          const uint32_t dcol = tmem + (uint32_t)(u * 2 + buf) * 128u;
          if (li >= 2) {
            tc::mbar_wait(&bars->d_empty[u][buf], (uint32_t)(((li >> 1) - 1) & 1));
            tc::mbar_wait_cluster(&bars->d_peer[u][buf], (uint32_t)(((li >> 1) - 1) & 1));
          }
          tc::tc_fence_after();
          if (cfg.trace && blockIdx.x == 0) cfg.trace[(li * 2 + u) * 8 + 0] = clock64();
          uint32_t abase, aplane;
          if (L.a_region == REG_MAIN) { abase = tc::smem_u32(main_of(u)); aplane = MAIN_PLANE; }
          else { aplane = cfg.small_plane; abase = tc::smem_u32(small_of(u)) + (L.a_region == REG_SA ? cfg.sa_off : 0u); }
          for (int i = 0; i < L.ksteps; ++i) {
            const int s = perm ? kperm(i) : i;
            if (L.a_wait && (perm ? (i & 3) == 0 : i == 0)) {
              const int c = perm ? (i >> 2) : 0;
              tc::mbar_wait(&bars->a_ready[u][c], (aph[u] >> c) & 1u);
              tc::mbar_wait_cluster(&bars->a_peer[u][c], (aph[u] >> c) & 1u); aph[u] ^= (1u << c);
              tc::tc_fence_after();
            }
            tc::mbar_wait(&bars->w_full[stage], wphase);
            tc::mbar_wait_cluster(&bars->w_peer[stage], wphase);
            tc::tc_fence_after();
            if (cfg.trace && blockIdx.x == 0 && i == 0) cfg.trace[(li * 2 + u) * 8 + 1] = clock64();
            const uint32_t ao = abase + (uint32_t)s * (2u * KG_BYTES), bo = wst0 + (uint32_t)stage * cfg.stage_bytes;
            const uint64_t ah = tc::make_smem_desc(ao, KG_BYTES, 128), bh = tc::make_smem_desc(bo, blbo, 128);
            tc::umma2_bf16(dcol, ah, bh, idesc, i > 0 ? 1u : 0u);
            if (NS == 2) {
              const uint64_t al = tc::make_smem_desc(ao + aplane, KG_BYTES, 128), bl = tc::make_smem_desc(bo + bplane, blbo, 128);
              tc::umma2_bf16(dcol, al, bh, idesc, 1u);
              tc::umma2_bf16(dcol, ah, bl, idesc, 1u);
            }
            tc::umma_commit2(&bars->w_empty[stage], 3);                        // frees this stage in both CTAs
            if (++stage == cfg.nst) { stage = 0; wphase ^= 1u; }
          }
          tc::umma_commit2(&bars->d_full[u][buf], 3);                          // accumulator ready in both CTAs' TMEM
          if (cfg.trace && blockIdx.x == 0) cfg.trace[(li * 2 + u) * 8 + 2] = clock64();
        }
      }
    }
  } else if (warp == RELAY_WARP) {
    // ================= peer CTA: forward the completion of my local a_ready / d_empty barriers to the leader =================
    // Walks the epilogue's event order: prologue (A0 per tile), then per (layer, tile) unit the A chunks it announces,
    // its accumulator release, and after the last member's transition3 the statistics phase's A0.
    if (rank != 0 && tc::elect_one()) {
      uint32_t aph[2] = {0u, 0u}, dph[2] = {0u, 0u};
      auto fwd_a = [&](int u, int c) {
        tc::mbar_wait(&bars->a_ready[u][c], (aph[u] >> c) & 1u); aph[u] ^= (1u << c);
        tc::mbar_arrive_remote(&bars->a_peer[u][c], 0);
      };
      auto fwd_d = [&](int u, int b) {
        tc::mbar_wait(&bars->d_empty[u][b], (dph[u] >> b) & 1u); dph[u] ^= (1u << b);
        tc::mbar_arrive_remote(&bars->d_peer[u][b], 0);
      };
      fwd_a(0, 0); fwd_a(1, 0);
      for (int l = 0; l < sched.n_layers; ++l) {
        const int kind = sched.L[l].kind;
        const bool last_mean = kind == EPI_MEAN && ((l - cfg.first_dyn) >> 3) == MB_E - 1;
        for (int u = 0; u < 2; ++u) {
          if (kind == EPI_SWISH256 || kind == EPI_RELU256) { fwd_a(u, 0); fwd_a(u, 1); fwd_a(u, 2); fwd_a(u, 3); }
          else if (kind == EPI_ZS || kind == EPI_G || kind == EPI_Z) fwd_a(u, 0);
          fwd_d(u, l & 1);
          if (last_mean) fwd_a(u, 0);
        }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3, group = warp >> 2;
    const int hi = q >> 1;                                      // column half held by this warp's TMEM lanes
    const int rl = (q & 1) * 32 + lane;                         // CTA-local row 0..63
    const int sub = hi * 4 + group;                             // 0..7: the 8 warps that share a row set
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const uint32_t sp = cfg.small_plane;
    float zs[2][8];
    float racc[2] = {0.f, 0.f}, pen[2] = {0.f, 0.f};

    auto grow0 = [&](int u) { return (size_t)pair_row0 + (size_t)u * 128 + (size_t)rank * RH; };
    auto dbase = [&](int u, int l) { return lane_base + (uint32_t)(u * 2 + (l & 1)) * 128u; };
    const bool tracer = cfg.trace && blockIdx.x < 2 && warp == 0 && lane == 0;     // warp 0 of the leader and of the peer
    auto wait_d = [&](int u, int l) {
      if (tracer) cfg.trace[(l * 2 + u) * 8 + 3 + 3 * (int)rank] = clock64();
      tc::mbar_wait(&bars->d_full[u][l & 1], (uint32_t)((l >> 1) & 1)); tc::tc_fence_after();
      if (tracer) cfg.trace[(l * 2 + u) * 8 + 4 + 3 * (int)rank] = clock64();
    };
    auto release_d = [&](int u, int l) {
      if (tracer && rank == 0) cfg.trace[(l * 2 + u) * 8 + 5] = clock64();
      tc::tc_fence_before(); __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->d_empty[u][l & 1]);
    };
    auto signal_a = [&](int u, int c) {
      tc::fence_proxy_async_smem(); __syncwarp();               // my smem writes -> visible to this CTA's tensor core
      if (lane == 0) tc::mbar_arrive(&bars->a_ready[u][c]);
    };
    auto bias_of = [&](int l) -> const float* {
      tc::mbar_wait(&bars->b_full[l & (NB - 1)], (uint32_t)((l / NB) & 1));
      return bias_s + (l & (NB - 1)) * BSLOT;
    };
    auto bias_done = [&](int l) { __syncwarp(); if (lane == 0) tc::mbar_arrive(&bars->b_empty[l & (NB - 1)]); };
    auto ld8s = [&](const float* b, float (&bv)[8]) {
      const float4 t0 = *reinterpret_cast<const float4*>(b), t1 = *(reinterpret_cast<const float4*>(b) + 1);
      bv[0] = t0.x; bv[1] = t0.y; bv[2] = t0.z; bv[3] = t0.w; bv[4] = t1.x; bv[5] = t1.y; bv[6] = t1.z; bv[7] = t1.w;
    };

    // 256-wide hidden layer of tile u: act(x + b) -> A_main[u]; 4 chunks of 32 TMEM columns, 8 columns per warp
    auto epi_act256 = [&](int l, int u, bool relu) {
      const float* bias = bias_of(l) + hi * 128 + group * 8;
      const uint32_t t0 = dbase(u, l) + (uint32_t)(group * 8);
      unsigned char* Am = main_of(u);
      wait_d(u, l);
      uint32_t xa[8], xb[8];
      auto chunk = [&](int c, const uint32_t (&x)[8]) {
        float bv[8], v[8];
        ld8s(bias + c * 32, bv);
        act8<NS>(x, bv, v, relu);
        store8<NS>(Am, MAIN_PLANE, (uint32_t)(hi * 16 + c * 4 + group) * KG_BYTES + (uint32_t)rl * 16u, v);
        signal_a(u, c);
      };
      tc::tmem_ld8(t0, xa);
#pragma unroll 1
      for (int c = 0; c < 4; c += 2) {
        tc::tmem_ld_wait();
        tc::tmem_ld8(t0 + (uint32_t)(c + 1) * 32u, xb);
        chunk(c, xa);
        tc::tmem_ld_wait();
        if (c + 2 < 4) tc::tmem_ld8(t0 + (uint32_t)(c + 2) * 32u, xa);
        chunk(c + 1, xb);
      }
      release_d(u, l);
      bias_done(l);
    };

    // narrow layer (N = 16 or 32 over the pair): this warp's columns are hi * N/2 + group * 8 + [0, 8) if they exist
    auto narrow_ld = [&](int l, int u, int n, uint32_t (&x)[8]) -> bool {
      const bool mine = group * 8 < (n >> 1);
      if (mine) { tc::tmem_ld8(dbase(u, l) + (uint32_t)(group * 8), x); tc::tmem_ld_wait(); }
      return mine;
    };

    // ---------------- prologue: obs (and given actions) of both tiles -> bf16 operand planes ----------------
    for (int u = 0; u < 2; ++u) {
      const size_t grow = grow0(u) + rl;
      const bool valid = grow < (size_t)live;
      const float* orow = a.obs + grow * S;
      unsigned char* As = small_of(u);
      for (int kg = sub; kg < (int)cfg.obs_kp / 8; kg += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { int j = kg * 8 + i; v[i] = (valid && j < S) ? __ldg(orow + j) : 0.f; }
        store8<NS>(As, sp, (uint32_t)kg * KG_BYTES + (uint32_t)rl * 16u, v);
      }
      if (!cfg.has_policy && sub < 2) {
        const float* arow = a.act + grow * A;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { int j = sub * 8 + i; v[i] = (valid && j < A) ? __ldg(arow + j) : 0.f; }
        store8<NS>(As, sp, cfg.sa_off + (uint32_t)(2 + sub) * KG_BYTES + (uint32_t)rl * 16u, v);
      }
      signal_a(u, 0);
    }

    // ---------------- ensemble statistics, noise, pick, penalty, termination, reward-head operand (tile u) ----------------
    auto stats = [&](int u) {
      constexpr int NEPI = 32 * EPI_WARPS;
      const size_t g0 = grow0(u);
      const int n_el = RH * S;
      float* nobs_s = reinterpret_cast<float*>(main_of(u));              // [64*S]  next_obs of this CTA's rows
      int* member_s = reinterpret_cast<int*>(nobs_s + n_el);             // [64]    picked member per row
      float* dsq_s = reinterpret_cast<float*>(member_s + RH) + RH;       // [batch][64*S] squared deviations
      const int scratch_floats = (int)(NS * MAIN_PLANE / 4) - n_el - 2 * RH;
      const int mbatch = min(MB_E, scratch_floats / n_el);
      float* redu = red + u * 8 * RH;
      const size_t grow = g0 + rl;
      const bool valid = grow < (size_t)live;
      if (tid < RH) {
        int mem = 0;
        if (g0 + tid < (size_t)live) {
          const size_t gr = g0 + tid;
          const unsigned long long gid = a.row_ids ? (unsigned long long)a.row_ids[gr] : a.row0 + gr;
          mem = a.idx ? (int)a.idx[gr] : (int)a.elites[philox_elite_slot(a.seed, a.step, gid, a.n_elites)];
        }
        member_s[tid] = mem;
      }
      epi_bar<EPI_WARPS>();   // every mean column of this tile half has been written; member_s is visible
      float row_pmax = 0.f;
      for (int e0 = 0; e0 < MB_E; e0 += mbatch) {
#pragma unroll 2
        for (int i = tid; i < n_el; i += NEPI) {
          const int rr = i / S, j = i - rr * S;
          const size_t gr = g0 + rr;
          if (gr < (size_t)live) {
            float mv[MB_E], sum = 0.f;
#pragma unroll
            for (int e = 0; e < MB_E; ++e) { mv[e] = a.mean[((size_t)e * B + g0) * S + i]; sum += mv[e]; }
            const float mbar = sum / (float)MB_E;
            const int member = member_s[rr];
            float ss = 0.f, mk = 0.f;
#pragma unroll
            for (int e = 0; e < MB_E; ++e) {
              const float d = mv[e] - mbar, dd = d * d;
              ss += dd;
              if (e >= e0 && e < e0 + mbatch) dsq_s[(e - e0) * n_el + i] = (j < S - 1) ? dd : 0.f;   // quirk: last dim excluded (:246)
              if (e == member) mk = mv[e];
            }
            if (e0 == 0) {
              const float sd = sqrtf(ss / (float)(MB_E - 1));
              float ep;
              if (a.eps) ep = a.eps[((size_t)member * B + gr) * S + j];
              else {
                const unsigned long long gid = a.row_ids ? (unsigned long long)a.row_ids[gr] : a.row0 + gr;
                ep = philox_normal1(philox_noise_block(a.seed, a.step, gid, (unsigned)(j >> 2)), j & 3);
              }
              const float nv = mk + ep * sd;
              a.next_obs[g0 * S + i] = nv;
              nobs_s[i] = nv;
            }
          }
        }
        epi_bar<EPI_WARPS>();
        // (row, member) sums in a fixed order; the 8 warps of a row set take members sub, sub + 8, ...
        if (valid) {
          for (int e = sub; e < mbatch && e0 + e < MB_E; e += 8) {
            const float* p = dsq_s + e * n_el + rl * S;
            float v0 = 0.f, v1 = 0.f;
            int j = 0;
            for (; j + 2 <= S - 1; j += 2) { v0 += p[j]; v1 += p[j + 1]; }
            if (j < S - 1) v0 += p[j];
            row_pmax = fmaxf(row_pmax, sqrtf(v0 + v1));
          }
        }
        if (e0 + mbatch < MB_E) epi_bar<EPI_WARPS>();   // dsq_s is rewritten by the next batch
      }
      redu[sub * RH + rl] = row_pmax;
      if (sub == 0 && valid) a.terminal[grow] = (unsigned char)mb_terminal(a.term_kind, nobs_s + rl * S, S);
      epi_bar<EPI_WARPS>();
      if (sub == 0 && valid) {
        float pm = 0.f;
#pragma unroll
        for (int g2 = 0; g2 < 8; ++g2) pm = fmaxf(pm, redu[g2 * RH + rl]);
        pen[u] = pm;
      }
      {   // sas = [obs, act, next_obs, 0-pad] operand of the reward head (mobody_module.py:296); aliases obs/sa planes
        const float* actp = cfg.has_policy ? a.act_out : a.act;
        unsigned char* As = small_of(u);
        for (int kg = sub; kg < (int)cfg.sas_kp / 8; kg += 8) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = kg * 8 + i;
            float t = 0.f;
            if (valid) {
              if (k < S) t = __ldg(a.obs + grow * S + k);
              else if (k < S + A) t = actp[grow * A + (k - S)];
              else if (k < 2 * S + A) t = nobs_s[rl * S + (k - S - A)];
            }
            v[i] = t;
          }
          store8<NS>(As, sp, (uint32_t)kg * KG_BYTES + (uint32_t)rl * 16u, v);
        }
        tc::fence_proxy_async_smem();
      }
      epi_bar<EPI_WARPS>();   // every sas plane is written (and nobs_s is no longer needed) before the operand is announced
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->a_ready[u][0]);
    };

    // ---------------- units in (layer, tile) order: the same order the MMA issuer follows ----------------
#pragma unroll 1
    for (int l = 0; l < sched.n_layers; ++l) {
      const int kind = sched.L[l].kind, n = sched.L[l].n;
      const int e = (l >= cfg.first_dyn && kind != EPI_REWARD) ? (l - cfg.first_dyn) >> 3 : 0;
#pragma unroll 1
      for (int u = 0; u < 2; ++u) {
        const size_t grow = grow0(u) + rl;
        const bool valid = grow < (size_t)live;
        switch (kind) {
          case EPI_SWISH256: epi_act256(l, u, false); break;
          case EPI_RELU256: epi_act256(l, u, true); break;
          case EPI_ACTION: {                                    // policy head: tanh * max_action -> [zs | act] operand, act_out
            const float* bias = bias_of(l);
            wait_d(u, l);
            uint32_t x[8];
            if (narrow_ld(l, u, n, x)) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int j = hi * (n >> 1) + group * 8 + i;
                v[i] = (j < A) ? tanhf(__uint_as_float(x[i]) + bias[j]) * a.max_action : 0.f;
                if (valid && j < A && a.act_out) a.act_out[grow * A + j] = v[i];
              }
              store8<NS>(small_of(u), sp, cfg.sa_off + (uint32_t)(2 + hi * (n >> 4) + group) * KG_BYTES + (uint32_t)rl * 16u, v);
            }
            release_d(u, l); bias_done(l);
          } break;
          case EPI_ZS: {                                        // zs3 mu half: keep zs, write the zs part of [zs | act]
            const float* bias = bias_of(l);
            wait_d(u, l);
            uint32_t x[8];
            if (narrow_ld(l, u, n, x)) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) { zs[u][i] = __uint_as_float(x[i]) + bias[hi * 8 + i]; v[i] = zs[u][i]; }
              store8<NS>(small_of(u), sp, cfg.sa_off + (uint32_t)hi * KG_BYTES + (uint32_t)rl * 16u, v);
            }
            signal_a(u, 0);
            release_d(u, l); bias_done(l);
          } break;
          case EPI_G: {                                         // za1: swish -> 32-wide operand (aliases A_main[u])
            const float* bias = bias_of(l);
            wait_d(u, l);
            uint32_t x[8];
            if (narrow_ld(l, u, n, x)) {
              float v[8];
              act8<NS>(x, bias + hi * 16 + group * 8, v, false);
              store8<NS>(main_of(u), MAIN_PLANE, (uint32_t)(hi * 2 + group) * KG_BYTES + (uint32_t)rl * 16u, v);
            }
            signal_a(u, 0);
            release_d(u, l); bias_done(l);
          } break;
          case EPI_Z: {                                         // za2 mu half: z = zs + za -> 16-wide operand
            const float* bias = bias_of(l);
            wait_d(u, l);
            uint32_t x[8];
            if (narrow_ld(l, u, n, x)) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = zs[u][i] + (__uint_as_float(x[i]) + bias[hi * 8 + i]);
              store8<NS>(main_of(u), MAIN_PLANE, (uint32_t)hi * KG_BYTES + (uint32_t)rl * 16u, v);
            }
            signal_a(u, 0);
            release_d(u, l); bias_done(l);
          } break;
          case EPI_MEAN: {                                      // transition3 -> mean[e] (info['samples'])
            const float* bias = bias_of(l);
            const int nh = n >> 1;
            wait_d(u, l);
            for (int c0 = group * 8; c0 < nh; c0 += 32) {
              uint32_t x[8];
              tc::tmem_ld8(dbase(u, l) + (uint32_t)c0, x);
              tc::tmem_ld_wait();
              if (valid) {
                float* mrow = a.mean + ((size_t)e * B + grow) * S;
#pragma unroll
                for (int i = 0; i < 8; ++i) { const int col = hi * nh + c0 + i; if (col < S) mrow[col] = __uint_as_float(x[i]) + bias[col]; }
              }
            }
            release_d(u, l); bias_done(l);
            if (e == MB_E - 1) stats(u);
          } break;
          case EPI_REWARD: {                                    // reward_model2 -> swish -> dot reward_model3[:,0]
            const float* bslot = bias_of(l);
            const float* bias = bslot + hi * 128 + group * 8;
            const float* w3 = bslot + 256 + hi * 128 + group * 8;
            const uint32_t t0 = dbase(u, l) + (uint32_t)(group * 8);
            float part = 0.f;
            wait_d(u, l);
            uint32_t xa[8], xb[8];
            auto chunk = [&](int c, const uint32_t (&x)[8]) {
              float bv[8], wv[8], v[8];
              ld8s(bias + c * 32, bv); ld8s(w3 + c * 32, wv);
              act8<NS>(x, bv, v, false);
#pragma unroll
              for (int i = 0; i < 8; ++i) part = fmaf(v[i], wv[i], part);
            };
            tc::tmem_ld8(t0, xa);
#pragma unroll 1
            for (int c = 0; c < 4; c += 2) {
              tc::tmem_ld_wait();
              tc::tmem_ld8(t0 + (uint32_t)(c + 1) * 32u, xb);
              chunk(c, xa);
              tc::tmem_ld_wait();
              if (c + 2 < 4) tc::tmem_ld8(t0 + (uint32_t)(c + 2) * 32u, xa);
              chunk(c + 1, xb);
            }
            const float b3 = bslot[512];
            release_d(u, l); bias_done(l);
            float* redu = red + u * 8 * RH;
            redu[sub * RH + rl] = part;
            epi_bar<EPI_WARPS>();
            if (sub == 0) {
              float sum = 0.f;
#pragma unroll
              for (int g2 = 0; g2 < 8; ++g2) sum += redu[g2 * RH + rl];
              racc[u] += sum + b3;
            }
          } break;
          default: break;
        }
      }
    }
    if (sub == 0) {
      for (int u = 0; u < 2; ++u) {
        const size_t grow = grow0(u) + rl;
        if (grow < (size_t)live) {
          const float raw = racc[u] / (float)MB_E;
          if (a.raw_reward) a.raw_reward[grow] = raw;
          a.penalty[grow] = pen[u];
          a.reward[grow] = (a.coef != 0.f && a.use_penalty) ? raw - a.coef * pen[u] : raw;     // :261-263
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();                      // the peer's smem / barriers stay alive until both CTAs are done
  if (warp == PROD_WARP) tc::tmem_dealloc2(tmem, 512);
}

}  // namespace tcp

long long* mb_tc_get_trace();

const char* mb_tc_pair_step_launch(const StepArgs& a, const unsigned char* dynb, const unsigned char* polb, int ns, cudaStream_t st) {
  if (a.B <= 0) return nullptr;
  const int S = a.S, A = a.A;
  if (S < 2 || S > 64 || A < 1 || A > 16) return "tensor-core step kernel supports 2 <= S <= 64, 1 <= A <= 16";
  if (!dynb) return "tensor-core step needs dyn_pack (mobody_dyn_pack)";
  const bool has_policy = polb != nullptr;
  if (has_policy && !a.act_out) return "tensor-core step with a fused policy needs act_out";
  const TcDynLayout DL = tc_dyn_layout(S, A, ns);
  TcSched sc{}; int n = 0;
  auto add = [&](size_t w_off, uint32_t b_off, const TcGeom& g, int region, int a_wait, int kind, int blob) {
    TcLayer& L = sc.L[n++];
    L.w_off = (uint32_t)w_off; L.b_off = b_off; L.ksteps = (uint16_t)(g.Kp / 16); L.n = (uint16_t)g.Np;
    L.a_region = (uint8_t)region; L.a_wait = (uint8_t)a_wait; L.kind = (uint8_t)kind; L.blob = (uint8_t)blob;
  };
  tcp::Cfg cfg{};
  // a_wait: 1 = the layer's A operand is announced by the preceding epilogue / prologue / statistics phase of the
  // same tile (once, or chunk by chunk for a 256-deep operand); 0 = static operand that is already in place
  if (has_policy) {
    const TcMlpLayout PL = tc_mlp_layout(S, A, ns);
    add(PL.w_off[0], PL.b_off[0], PL.g[0], REG_OBS, 1, EPI_RELU256, 1);
    add(PL.w_off[1], PL.b_off[1], PL.g[1], REG_MAIN, 1, EPI_RELU256, 1);
    add(PL.w_off[2], PL.b_off[2], PL.g[2], REG_MAIN, 1, EPI_ACTION, 1);
    cfg.pol_bias_base = (uint32_t)PL.bias_base;
  }
  sc.first_dyn = n; cfg.first_dyn = n;
  const int za1 = a.use_trg ? PK_ZATRG1 : PK_ZASRC1, za2 = a.use_trg ? PK_ZATRG2 : PK_ZASRC2;
  const int sas_kp = tc_rup16(2 * S + A);
  for (int e = 0; e < MB_E; ++e) {
    const size_t wb = (size_t)e * DL.member_w_bytes; const uint32_t bb = (uint32_t)e * DL.member_b_floats;
    auto lay = [&](int pk, int region, int a_wait, int kind) { add(wb + DL.w_off[pk], bb + DL.b_off[pk], tc_dyn_geom(pk, S, A), region, a_wait, kind, 0); };
    lay(PK_ZS1, REG_OBS, (e == 0 && !has_policy) ? 1 : 0, EPI_SWISH256);
    lay(PK_ZS2, REG_MAIN, 1, EPI_SWISH256);
    lay(PK_ZS3, REG_MAIN, 1, EPI_ZS);
    lay(za1, REG_SA, 1, EPI_G);
    lay(za2, REG_MAIN, 1, EPI_Z);
    lay(PK_T1, REG_MAIN, 1, EPI_SWISH256);
    lay(PK_T2, REG_MAIN, 1, EPI_SWISH256);
    lay(PK_T3, REG_MAIN, 1, EPI_MEAN);
  }
  for (int e = 0; e < MB_E; ++e) {
    const size_t wb = (size_t)e * DL.member_w_bytes; const uint32_t bb = (uint32_t)e * DL.member_b_floats;
    add(wb + DL.w_off[PK_R1], bb + DL.b_off[PK_R1], tc_dyn_geom(PK_R1, S, A), REG_SAS, e == 0 ? 1 : 0, EPI_SWISH256, 0);
    add(wb + DL.w_off[PK_R2], bb + DL.b_off[PK_R2], tc_dyn_geom(PK_R2, S, A), REG_MAIN, 1, EPI_REWARD, 0);
  }
  sc.n_layers = n;
  cfg.obs_kp = (uint32_t)tc_rup16(S); cfg.sas_kp = (uint32_t)sas_kp;
  cfg.sa_off = (cfg.obs_kp / 8) * tcp::KG_BYTES;
  uint32_t small = cfg.sa_off + 4 * tcp::KG_BYTES, sasb = (cfg.sas_kp / 8) * tcp::KG_BYTES;
  cfg.small_plane = small > sasb ? small : sasb;
  cfg.stage_bytes = (uint32_t)ns * 128u * 32u;                 // one K step of this CTA's half of B, all planes
  cfg.dyn_bias_base = (uint32_t)DL.bias_base;
  cfg.has_policy = has_policy ? 1 : 0;
  cfg.trace = mb_tc_get_trace();
  const size_t fixed = 2 * ((size_t)ns * tcp::MAIN_PLANE + (size_t)ns * cfg.small_plane) +
                       (2 * 8 * tcp::RH + tcp::NB * tcp::BSLOT) * sizeof(float) + sizeof(tcp::Bars) + 128;
  const size_t budget = 227 * 1024;
  if (fixed + 2 * cfg.stage_bytes > budget) return "tensor-core pair step kernel: shared memory budget exceeded for this (S, A)";
  int nst = (int)((budget - fixed) / cfg.stage_bytes);
  if (nst > tcp::MAX_NST) nst = tcp::MAX_NST;
  cfg.nst = nst;
  const size_t bytes = fixed + (size_t)nst * cfg.stage_bytes;
  auto kern = ns == 2 ? tcp::step_pair_kernel<2> : tcp::step_pair_kernel<1>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return "cudaFuncSetAttribute(step_pair_kernel) failed";
  const int grid = 2 * ((a.B + 255) / 256);                    // CTA pairs (cluster dims 2x1x1), 256 rows per pair
  kern<<<grid, tcp::NTHREADS, bytes, st>>>(a, dynb, polb ? polb : dynb, sc, cfg);
  return nullptr;
}
