// Fused one-step model rollout on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same math as step_simt.cu / SURVEY.md Appendix A.1 (reference
// algo/dynamics/mobody_dynamics.py:193-265, algo/dynamics/mobody_module.py:217-330,
// algo/offline_offline/mobody.py:60-72), restructured for Blackwell:
//
//  * one CTA owns 128 start states (UMMA M = 128) and walks the whole per-row chain
//      [policy 3 GEMMs] + 7 x 8 dynamics GEMMs + ensemble statistics + 7 x 2 reward GEMMs
//    without the activations ever leaving the SM;
//  * layer weights are pre-packed (mobody_dyn_pack) as bf16 planes in the exact UMMA K-major
//    shared-memory image and streamed L2 -> SMEM by a producer warp with 1-D TMA bulk copies
//    (cp.async.bulk + mbarrier complete_tx) through an NST-deep ring;
//  * one elected thread issues tcgen05.mma (128 x N x 16 per instruction), accumulators live in
//    TMEM (two 256-column buffers, ping-pong across layers);
//  * 8 epilogue warps read TMEM (tcgen05.ld), apply bias + swish/relu, split to bf16 planes and write
//    the next layer's A operand in place in SMEM, 32 columns at a time; the next layer's MMAs start
//    as soon as the first 32-column chunk is announced (mbarrier per chunk), so tensor pipe and
//    epilogue overlap inside a single row tile.
//
// Precision: NS = 1 -> single bf16 pass (stated bound 5e-3); NS = 2 -> bf16 hi+lo split of both
// operands, 3 MMAs per K step (hi*hi + lo*hi + hi*lo), ~2^-16 per product: inside the 1e-4 bound.
#include "common.cuh"
#include "philox.cuh"
#include "term.cuh"
#include "tc_prims.cuh"
#include "tc_layout.h"

namespace tcs {

constexpr int EPI_WARPS = 16;      // warps 0..15: epilogue, 4 groups x 4 TMEM lane quadrants
constexpr int NGROUPS = EPI_WARPS / 4;
constexpr int PROD_WARP = 16;      // weight producer (+ TMEM alloc/dealloc)
constexpr int MMA_WARP = 17;       // MMA issuer
constexpr int NTHREADS = 32 * (EPI_WARPS + 2);
constexpr int TM = 128;
constexpr uint32_t MAIN_PLANE = 65536;   // 128 rows x 256 k x bf16
constexpr int MAX_NST = 12;

struct Cfg {
  int nst;
  uint32_t stage_bytes, small_plane, sa_off, obs_kp, sas_kp;
  uint32_t dyn_bias_base, pol_bias_base, member_b_floats, r3_b_off;
  int has_policy;
};

struct Bars {
  uint64_t w_full[MAX_NST], w_empty[MAX_NST], a_ready[8], d_full[2], d_empty[2];
  uint32_t tmem_slot, pad;
};

template <int NS>
__device__ __forceinline__ void store8(unsigned char* base, uint32_t plane_stride, uint32_t off, const float (&v)[8]) {
  if (NS == 2) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) tc::split2(v[2 * i], v[2 * i + 1], h[i], l[i]);
    *reinterpret_cast<uint4*>(base + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(base + plane_stride + off) = make_uint4(l[0], l[1], l[2], l[3]);
  } else {
    *reinterpret_cast<uint4*>(base + off) = make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]),
                                                       tc::pack_bf16(v[4], v[5]), tc::pack_bf16(v[6], v[7]));
  }
}

template <int NS> __device__ __forceinline__ float swish_ns(float x) { return NS == 1 ? tc::swish_tanh(x) : tc::swish_ex2_rcp(x); }

template <int NS>
__device__ __forceinline__ void store1(unsigned char* base, uint32_t plane_stride, uint32_t off, float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(base + off) = h;
  if (NS == 2) *reinterpret_cast<__nv_bfloat16*>(base + plane_stride + off) = __float2bfloat16_rn(v - __bfloat162float(h));
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory"); }   // epilogue warps only

template <int NS>
__global__ void __launch_bounds__(NTHREADS, 1)
step_tc_kernel(const StepArgs a, const unsigned char* __restrict__ dynb, const unsigned char* __restrict__ polb,
               const __grid_constant__ TcSched sched, const Cfg cfg) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int S = a.S, A = a.A, B = a.B;
  const int live = a.n_rows_dev ? min(*a.n_rows_dev, B) : B;
  const int row0 = blockIdx.x * TM;
  if (row0 >= live) return;

  unsigned char* A_main = smem;
  unsigned char* A_small = A_main + NS * MAIN_PLANE;
  unsigned char* wst = A_small + NS * cfg.small_plane;
  float* red = reinterpret_cast<float*>(wst + (size_t)cfg.nst * cfg.stage_bytes);      // [2][NGROUPS][128]
  Bars* bars = reinterpret_cast<Bars*>(red + 2 * NGROUPS * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < cfg.nst; ++i) { tc::mbar_init(&bars->w_full[i], 1); tc::mbar_init(&bars->w_empty[i], 1); }
    for (int i = 0; i < 8; ++i) tc::mbar_init(&bars->a_ready[i], 4);       // 4 warps own each 32-column chunk
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&bars->d_full[i], 1); tc::mbar_init(&bars->d_empty[i], EPI_WARPS); }
    tc::mbar_fence_init();
  }
  if (warp == PROD_WARP) tc::tmem_alloc(&bars->tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = bars->tmem_slot;

  if (warp == PROD_WARP) {
    // ================= weight producer: L2 -> SMEM ring, one K step (all planes) per stage =================
    if (tc::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int li = 0; li < sched.n_layers; ++li) {
        const TcLayer L = sched.L[li];
        const unsigned char* src = (L.blob ? polb : dynb) + L.w_off;
        const uint32_t bytes = (uint32_t)NS * L.n * 32u;
        for (int s = 0; s < L.ksteps; ++s) {
          tc::mbar_wait(&bars->w_empty[stage], phase ^ 1u);
          tc::mbar_arrive_expect_tx(&bars->w_full[stage], bytes);
          tc::bulk_g2s(wst + (size_t)stage * cfg.stage_bytes, src + (size_t)s * bytes, bytes, &bars->w_full[stage]);
          if (++stage == cfg.nst) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================= MMA issuer =================
    if (tc::elect_one()) {
      int stage = 0; uint32_t wphase = 0, aph = 0;
      const uint32_t small0 = tc::smem_u32(A_small), main0 = tc::smem_u32(A_main), wst0 = tc::smem_u32(wst);
      for (int li = 0; li < sched.n_layers; ++li) {
        const TcLayer L = sched.L[li];
        const int buf = li & 1;
        const uint32_t dcol = tmem + (uint32_t)buf * 256u;
        if (li >= 2) tc::mbar_wait(&bars->d_empty[buf], (uint32_t)(((li >> 1) - 1) & 1));
        tc::tc_fence_after();
        uint32_t abase, aplane;
        if (L.a_region == REG_MAIN) { abase = main0; aplane = MAIN_PLANE; }
        else { aplane = cfg.small_plane; abase = small0 + (L.a_region == REG_SA ? cfg.sa_off : 0u); }
        const uint32_t idesc = tc::make_idesc_bf16(128, L.n);
        const uint32_t bplane = (uint32_t)L.n * 32u, blbo = (uint32_t)L.n * 16u;
        for (int s = 0; s < L.ksteps; ++s) {
          if ((s & 1) == 0 && (s >> 1) < L.a_wait) {
            const int c = s >> 1;
            tc::mbar_wait(&bars->a_ready[c], (aph >> c) & 1u); aph ^= (1u << c);
            tc::tc_fence_after();
          }
          tc::mbar_wait(&bars->w_full[stage], wphase);
          tc::tc_fence_after();
          const uint32_t ao = abase + (uint32_t)s * 4096u, bo = wst0 + (uint32_t)stage * cfg.stage_bytes;
          const uint64_t ah = tc::make_smem_desc(ao, 2048, 128), bh = tc::make_smem_desc(bo, blbo, 128);
          tc::umma_bf16(dcol, ah, bh, idesc, s > 0 ? 1u : 0u);
          if (NS == 2) {
            const uint64_t al = tc::make_smem_desc(ao + aplane, 2048, 128), bl = tc::make_smem_desc(bo + bplane, blbo, 128);
            tc::umma_bf16(dcol, al, bh, idesc, 1u);
            tc::umma_bf16(dcol, ah, bl, idesc, 1u);
          }
          tc::umma_commit(&bars->w_empty[stage]);
          if (++stage == cfg.nst) { stage = 0; wphase ^= 1u; }
        }
        tc::umma_commit(&bars->d_full[buf]);
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3, group = warp >> 2;
    const int r = q * 32 + lane;
    const bool valid = (row0 + r) < live;
    const size_t grow = (size_t)row0 + r;                     // row in this launch's arrays
    const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
    const float* dyn_bias = reinterpret_cast<const float*>(dynb + cfg.dyn_bias_base);
    const float* pol_bias = reinterpret_cast<const float*>(polb + cfg.pol_bias_base);
    const uint32_t sp = cfg.small_plane;
    int li = 0;
    float zs[16];
    float racc = 0.f, pen = 0.f;

    auto wait_d = [&](int l) { tc::mbar_wait(&bars->d_full[l & 1], (uint32_t)((l >> 1) & 1)); tc::tc_fence_after(); };
    auto release_d = [&](int l) {
      tc::tc_fence_before(); __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->d_empty[l & 1]);
    };
    auto signal_a = [&](int c) {
      tc::fence_proxy_async_smem(); __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars->a_ready[c]);
    };
    auto bias_of = [&](int l) { const TcLayer L = sched.L[l]; return (L.blob ? pol_bias : dyn_bias) + L.b_off; };

    // 256-wide hidden layer: act(x + b) -> A_main planes, chunk by chunk (group g owns chunks g, g+4)
    auto epi_act256 = [&](int l, bool relu) {
      const float* bias = bias_of(l);
      const uint32_t t0 = lane_addr + (uint32_t)(l & 1) * 256u;
      float4 bv[8];                                            // bias of the first chunk: fetched before the wait
#pragma unroll
      for (int i = 0; i < 8; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(bias + group * 32) + i);
      wait_d(l);
#pragma unroll 1
      for (int cc = 0; cc < 8 / NGROUPS; ++cc) {
        const int c = cc * NGROUPS + group;
        uint32_t x[32];
        tc::tmem_ld32(t0 + (uint32_t)c * 32u, x);
        tc::tmem_ld_wait();
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = bv[kg * 2 + (i >> 2)];
            const float bb = (i & 3) == 0 ? b4.x : (i & 3) == 1 ? b4.y : (i & 3) == 2 ? b4.z : b4.w;
            const float t = __uint_as_float(x[kg * 8 + i]) + bb;
            v[i] = relu ? fmaxf(t, 0.f) : swish_ns<NS>(t);
          }
          store8<NS>(A_main, MAIN_PLANE, (uint32_t)(c * 4 + kg) * 2048u + (uint32_t)r * 16u, v);
        }
        signal_a(c);
        if (cc + 1 < 8 / NGROUPS) {
#pragma unroll
          for (int i = 0; i < 8; ++i) bv[i] = __ldg(reinterpret_cast<const float4*>(bias + (c + NGROUPS) * 32) + i);
        }
      }
      release_d(l);
    };

    // ---------------- prologue: obs (and given actions) -> bf16 operand planes ----------------
    if (group == 0) {
      const float* orow = a.obs + grow * S;
      for (int kg = 0; kg < (int)cfg.obs_kp / 8; ++kg) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { int j = kg * 8 + i; v[i] = (valid && j < S) ? __ldg(orow + j) : 0.f; }
        store8<NS>(A_small, sp, (uint32_t)kg * 2048u + (uint32_t)r * 16u, v);
      }
      if (!cfg.has_policy) {
        const float* arow = a.act + grow * A;
#pragma unroll
        for (int kg = 0; kg < 2; ++kg) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) { int j = kg * 8 + i; v[i] = (valid && j < A) ? __ldg(arow + j) : 0.f; }
          store8<NS>(A_small, sp, cfg.sa_off + (uint32_t)(2 + kg) * 2048u + (uint32_t)r * 16u, v);
        }
      }
      signal_a(0);
    }

    // ---------------- policy: relu MLP, tanh * max_action (mobody.py:35-72) ----------------
    if (cfg.has_policy) {
      epi_act256(li++, true);
      epi_act256(li++, true);
      {
        const int l = li++;
        const float* bias = bias_of(l);
        wait_d(l);
        if (group == 0) {
          uint32_t x[16];
          tc::tmem_ld16(lane_addr + (uint32_t)(l & 1) * 256u, x);
          tc::tmem_ld_wait();
          float act[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) act[j] = (j < A) ? tanhf(__uint_as_float(x[j]) + __ldg(bias + j)) * a.max_action : 0.f;
          if (valid && a.act_out)
            for (int j = 0; j < A; ++j) a.act_out[grow * A + j] = act[j];
#pragma unroll
          for (int kg = 0; kg < 2; ++kg) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = act[kg * 8 + i];
            store8<NS>(A_small, sp, cfg.sa_off + (uint32_t)(2 + kg) * 2048u + (uint32_t)r * 16u, v);
          }
        }
        release_d(l);
      }
    }

    // ---------------- 7 members: forward_trg / forward_src (mobody_module.py:315-330) ----------------
#pragma unroll 1
    for (int e = 0; e < MB_E; ++e) {
      epi_act256(li++, false);                                  // zs1
      epi_act256(li++, false);                                  // zs2
      {                                                         // zs3 mu half -> [zs | act] operand
        const int l = li++;
        const float* bias = bias_of(l);
        wait_d(l);
        if (group == 0) {
          uint32_t x[16];
          tc::tmem_ld16(lane_addr + (uint32_t)(l & 1) * 256u, x);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) zs[j] = __uint_as_float(x[j]) + __ldg(bias + j);
#pragma unroll
          for (int kg = 0; kg < 2; ++kg) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = zs[kg * 8 + i];
            store8<NS>(A_small, sp, cfg.sa_off + (uint32_t)kg * 2048u + (uint32_t)r * 16u, v);
          }
          signal_a(0);
        }
        release_d(l);
      }
      {                                                         // za1: swish -> 32-wide operand (aliases A_main)
        const int l = li++;
        const float* bias = bias_of(l);
        wait_d(l);
        if (group == 0) {
          uint32_t x[32];
          tc::tmem_ld32(lane_addr + (uint32_t)(l & 1) * 256u, x);
          tc::tmem_ld_wait();
#pragma unroll
          for (int kg = 0; kg < 4; ++kg) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = swish_ns<NS>(__uint_as_float(x[kg * 8 + i]) + __ldg(bias + kg * 8 + i));
            store8<NS>(A_main, MAIN_PLANE, (uint32_t)kg * 2048u + (uint32_t)r * 16u, v);
          }
          signal_a(0);
        }
        release_d(l);
      }
      {                                                         // za2 mu half: z = zs + za -> 16-wide operand
        const int l = li++;
        const float* bias = bias_of(l);
        wait_d(l);
        if (group == 0) {
          uint32_t x[16];
          tc::tmem_ld16(lane_addr + (uint32_t)(l & 1) * 256u, x);
          tc::tmem_ld_wait();
#pragma unroll
          for (int kg = 0; kg < 2; ++kg) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = zs[kg * 8 + i] + (__uint_as_float(x[kg * 8 + i]) + __ldg(bias + kg * 8 + i));
            store8<NS>(A_main, MAIN_PLANE, (uint32_t)kg * 2048u + (uint32_t)r * 16u, v);
          }
          signal_a(0);
        }
        release_d(l);
      }
      epi_act256(li++, false);                                  // transition1
      epi_act256(li++, false);                                  // transition2
      {                                                         // transition3 -> mean[e] (info['samples'])
        const int l = li++;
        const float* bias = bias_of(l);
        const int np = sched.L[l].n;
        wait_d(l);
        for (int c = group; c * 32 < np; c += NGROUPS) {
          uint32_t x[32];
          if (np - c * 32 >= 32) tc::tmem_ld32(lane_addr + (uint32_t)(l & 1) * 256u + (uint32_t)c * 32u, x);
          else {
            uint32_t y[16];
            tc::tmem_ld16(lane_addr + (uint32_t)(l & 1) * 256u + (uint32_t)c * 32u, y);
#pragma unroll
            for (int j = 0; j < 16; ++j) { x[j] = y[j]; x[16 + j] = 0u; }
          }
          tc::tmem_ld_wait();
          if (valid) {
            float* mrow = a.mean + ((size_t)e * B + grow) * S;
#pragma unroll
            for (int j = 0; j < 32; ++j) { int col = c * 32 + j; if (col < S) mrow[col] = __uint_as_float(x[j]) + __ldg(bias + col); }
          }
        }
        release_d(l);
      }
    }

    // ---------------- ensemble statistics, noise, pick, penalty, termination ----------------
    epi_bar();   // mean columns >= 32 were written by the other group's thread of this row
    if (group == 0) {
      int member = 0;
      float d2[MB_E];
#pragma unroll
      for (int e = 0; e < MB_E; ++e) d2[e] = 0.f;
      const unsigned long long gid = a.row_ids ? (unsigned long long)a.row_ids[valid ? grow : 0] : a.row0 + grow;
      if (valid) {
        member = a.idx ? (int)a.idx[grow] : (int)a.elites[philox_elite_slot(a.seed, a.step, gid, a.n_elites)];
        float nrm[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < S; ++j) {
          if (!a.eps && (j & 3) == 0) philox_normal4(philox_noise_block(a.seed, a.step, gid, (unsigned)(j >> 2)), nrm);
          float mv[MB_E], sum = 0.f;
#pragma unroll
          for (int e = 0; e < MB_E; ++e) { mv[e] = a.mean[((size_t)e * B + grow) * S + j]; sum += mv[e]; }
          const float mbar = sum / (float)MB_E;
          float ss = 0.f, mk = 0.f;
#pragma unroll
          for (int e = 0; e < MB_E; ++e) {
            const float d = mv[e] - mbar; ss = fmaf(d, d, ss);
            if (j < S - 1) d2[e] = fmaf(d, d, d2[e]);            // quirk: last state dim excluded (:246)
            if (e == member) mk = mv[e];
          }
          const float sd = sqrtf(ss / (float)(MB_E - 1));
          const float ep = a.eps ? a.eps[((size_t)member * B + grow) * S + j] : nrm[j & 3];
          a.next_obs[grow * S + j] = mk + ep * sd;
        }
        float pmax = 0.f;
#pragma unroll
        for (int e = 0; e < MB_E; ++e) pmax = fmaxf(pmax, sqrtf(d2[e]));
        pen = pmax;
        a.terminal[grow] = (unsigned char)mb_terminal(a.term_kind, a.next_obs + grow * S, S);
      }
      // sas = [obs, act, next_obs, 0-pad] operand of the reward head (mobody_module.py:296)
      const float* actp = cfg.has_policy ? a.act_out : a.act;
      for (int k = 0; k < (int)cfg.sas_kp; ++k) {
        float v = 0.f;
        if (valid) {
          if (k < S) v = __ldg(a.obs + grow * S + k);
          else if (k < S + A) v = actp[grow * A + (k - S)];
          else if (k < 2 * S + A) v = a.next_obs[grow * S + (k - S - A)];
        }
        store1<NS>(A_small, sp, (uint32_t)(k >> 3) * 2048u + (uint32_t)r * 16u + (uint32_t)(k & 7) * 2u, v);
      }
      tc::fence_proxy_async_smem();
    }
    epi_bar();
    for (int c = group; c * 32 < (int)cfg.sas_kp; c += NGROUPS) { __syncwarp(); if (lane == 0) tc::mbar_arrive(&bars->a_ready[c]); }

    // ---------------- reward head, all 7 members (mobody_module.py:295-302; mean over members :236) ----------------
#pragma unroll 1
    for (int e = 0; e < MB_E; ++e) {
      epi_act256(li++, false);                                  // reward_model1
      {                                                         // reward_model2 -> swish -> dot reward_model3[:,0]
        const int l = li++;
        const float* bias = bias_of(l);
        const float* w3 = dyn_bias + (size_t)e * cfg.member_b_floats + cfg.r3_b_off;
        wait_d(l);
        const uint32_t t0 = lane_addr + (uint32_t)(l & 1) * 256u;
        float part = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 8 / NGROUPS; ++cc) {
          const int c = cc * NGROUPS + group;
          uint32_t x[32];
          tc::tmem_ld32(t0 + (uint32_t)c * 32u, x);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c * 32) + i);
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(w3 + c * 32) + i);
            part = fmaf(swish_ns<NS>(__uint_as_float(x[4 * i + 0]) + b4.x), w4.x, part);
            part = fmaf(swish_ns<NS>(__uint_as_float(x[4 * i + 1]) + b4.y), w4.y, part);
            part = fmaf(swish_ns<NS>(__uint_as_float(x[4 * i + 2]) + b4.z), w4.z, part);
            part = fmaf(swish_ns<NS>(__uint_as_float(x[4 * i + 3]) + b4.w), w4.w, part);
          }
        }
        release_d(l);
        red[((e & 1) * NGROUPS + group) * 128 + r] = part;
        epi_bar();
        if (group == 0) {
          float sum = 0.f;
#pragma unroll
          for (int g2 = 0; g2 < NGROUPS; ++g2) sum += red[((e & 1) * NGROUPS + g2) * 128 + r];
          racc += sum + __ldg(w3 + 256);
        }
      }
    }
    if (group == 0 && valid) {
      const float raw = racc / (float)MB_E;
      if (a.raw_reward) a.raw_reward[grow] = raw;
      a.penalty[grow] = pen;
      a.reward[grow] = (a.coef != 0.f && a.use_penalty) ? raw - a.coef * pen : raw;     // :261-263
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == PROD_WARP) tc::tmem_dealloc(tmem, 512);
}

// ---------------- weight packing: fp32 [K][N] (any strides) -> bf16 planes in UMMA B layout ----------------
__global__ void pack_weight_kernel(const float* __restrict__ W, long long stride_k, long long stride_n, int K, int N, int Kp,
                                   int Np, int ns, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)(Kp / 16) * ns * 2 * Np * 8;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int j = (int)(t & 7); long long u = t >> 3;
    int n = (int)(u % Np); u /= Np;
    int g = (int)(u & 1); u >>= 1;
    int p = (int)(u % ns); int s = (int)(u / ns);
    int k = s * 16 + g * 8 + j;
    float v = (k < K && n < N) ? W[k * stride_k + n * stride_n] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[t] = (p == 0) ? h : __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
__global__ void pack_bias_kernel(const float* __restrict__ b, long long stride, int N, int Np, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Np) out[i] = (i < N) ? b[i * stride] : 0.f;
}

}  // namespace tcs

static inline void launch_pack_w(const float* W, long long sk, long long sn, const TcGeom& g, int ns, unsigned char* out, cudaStream_t st) {
  long long total = (long long)(g.Kp / 16) * ns * 2 * g.Np * 8;
  int grid = (int)((total + 255) / 256); if (grid > 1184) grid = 1184;
  tcs::pack_weight_kernel<<<grid, 256, 0, st>>>(W, sk, sn, g.K, g.N, g.Kp, g.Np, ns, reinterpret_cast<__nv_bfloat16*>(out));
}

// mobody_dyn_pack: all 7 members x 12 MMA layers + biases + reward_model3 vector
const char* mb_tc_dyn_pack(const DynPtrs& dp, int S, int A, int ns, unsigned char* blob, cudaStream_t st) {
  if (ns != 1 && ns != 2) return "dyn_pack: nsplit must be 1 or 2";
  const TcDynLayout L = tc_dyn_layout(S, A, ns);
  static const int src[PK_COUNT] = {L_ZS1, L_ZS2, L_ZS3, L_ZASRC1, L_ZASRC2, L_ZATRG1, L_ZATRG2, L_T1, L_T2, L_T3, L_R1, L_R2};
  float* bias = reinterpret_cast<float*>(blob + L.bias_base);
  for (int e = 0; e < MB_E; ++e)
    for (int i = 0; i < PK_COUNT; ++i) {
      const TcGeom g = tc_dyn_geom(i, S, A);
      const int nfull = (i == PK_ZS3 || i == PK_ZASRC2 || i == PK_ZATRG2) ? 32 : g.N;     // mu half of 32 columns
      const float* W = dp.w[src[i]] + (size_t)e * g.K * nfull;
      launch_pack_w(W, nfull, 1, g, ns, blob + (size_t)e * L.member_w_bytes + L.w_off[i], st);
      tcs::pack_bias_kernel<<<(g.Np + 127) / 128, 128, 0, st>>>(dp.b[src[i]] + (size_t)e * nfull, 1, g.N, g.Np,
                                                                bias + (size_t)e * L.member_b_floats + L.b_off[i]);
    }
  for (int e = 0; e < MB_E; ++e) {   // reward_model3: column 0 of [256][2] and bias[0]
    float* o = bias + (size_t)e * L.member_b_floats + L.b_off[PK_COUNT];
    tcs::pack_bias_kernel<<<2, 128, 0, st>>>(dp.w[L_R3] + (size_t)e * MB_H * 2, 2, MB_H, MB_H, o);
    tcs::pack_bias_kernel<<<1, 32, 0, st>>>(dp.b[L_R3] + (size_t)e * 2, 1, 1, 16, o + MB_H);
  }
  return nullptr;
}

const char* mb_tc_mlp_pack(const MlpPtrs& mp, int din, int dout, int ns, unsigned char* blob, cudaStream_t st) {
  if (ns != 1 && ns != 2) return "mlp_pack: nsplit must be 1 or 2";
  const TcMlpLayout L = tc_mlp_layout(din, dout, ns);
  float* bias = reinterpret_cast<float*>(blob + L.bias_base);
  for (int i = 0; i < 3; ++i) {   // nn.Linear weight is [out][in]: stride_k = 1, stride_n = K
    launch_pack_w(mp.w[i], 1, L.g[i].K, L.g[i], ns, blob + L.w_off[i], st);
    tcs::pack_bias_kernel<<<(L.g[i].Np + 127) / 128, 128, 0, st>>>(mp.b[i], 1, L.g[i].N, L.g[i].Np, bias + L.b_off[i]);
  }
  return nullptr;
}

const char* mb_tc_step_launch(const StepArgs& a, const unsigned char* dynb, const unsigned char* polb, int ns, cudaStream_t st) {
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
  tcs::Cfg cfg{};
  if (has_policy) {
    const TcMlpLayout PL = tc_mlp_layout(S, A, ns);
    add(PL.w_off[0], PL.b_off[0], PL.g[0], REG_OBS, 1, EPI_RELU256, 1);
    add(PL.w_off[1], PL.b_off[1], PL.g[1], REG_MAIN, 8, EPI_RELU256, 1);
    add(PL.w_off[2], PL.b_off[2], PL.g[2], REG_MAIN, 8, EPI_ACTION, 1);
    cfg.pol_bias_base = (uint32_t)PL.bias_base;
  }
  sc.first_dyn = n;
  const int za1 = a.use_trg ? PK_ZATRG1 : PK_ZASRC1, za2 = a.use_trg ? PK_ZATRG2 : PK_ZASRC2;
  const int sas_kp = tc_rup16(2 * S + A);
  for (int e = 0; e < MB_E; ++e) {
    const size_t wb = (size_t)e * DL.member_w_bytes; const uint32_t bb = (uint32_t)e * DL.member_b_floats;
    auto lay = [&](int pk, int region, int a_wait, int kind) { add(wb + DL.w_off[pk], bb + DL.b_off[pk], tc_dyn_geom(pk, S, A), region, a_wait, kind, 0); };
    lay(PK_ZS1, REG_OBS, (e == 0 && !has_policy) ? 1 : 0, EPI_SWISH256);
    lay(PK_ZS2, REG_MAIN, 8, EPI_SWISH256);
    lay(PK_ZS3, REG_MAIN, 8, EPI_ZS);
    lay(za1, REG_SA, 1, EPI_G);
    lay(za2, REG_MAIN, 1, EPI_Z);
    lay(PK_T1, REG_MAIN, 1, EPI_SWISH256);
    lay(PK_T2, REG_MAIN, 8, EPI_SWISH256);
    lay(PK_T3, REG_MAIN, 8, EPI_MEAN);
  }
  for (int e = 0; e < MB_E; ++e) {
    const size_t wb = (size_t)e * DL.member_w_bytes; const uint32_t bb = (uint32_t)e * DL.member_b_floats;
    add(wb + DL.w_off[PK_R1], bb + DL.b_off[PK_R1], tc_dyn_geom(PK_R1, S, A), REG_SAS, e == 0 ? (sas_kp + 31) / 32 : 0, EPI_SWISH256, 0);
    add(wb + DL.w_off[PK_R2], bb + DL.b_off[PK_R2], tc_dyn_geom(PK_R2, S, A), REG_MAIN, 8, EPI_REWARD, 0);
  }
  sc.n_layers = n;
  cfg.obs_kp = (uint32_t)tc_rup16(S); cfg.sas_kp = (uint32_t)sas_kp;
  cfg.sa_off = (cfg.obs_kp / 8) * 2048u;
  uint32_t small = cfg.sa_off + 4 * 2048u, sasb = (cfg.sas_kp / 8) * 2048u;
  cfg.small_plane = small > sasb ? small : sasb;
  cfg.stage_bytes = (uint32_t)ns * 256u * 32u;
  cfg.dyn_bias_base = (uint32_t)DL.bias_base; cfg.member_b_floats = DL.member_b_floats; cfg.r3_b_off = DL.b_off[PK_COUNT];
  cfg.has_policy = has_policy ? 1 : 0;
  const size_t fixed = (size_t)ns * tcs::MAIN_PLANE + (size_t)ns * cfg.small_plane + 2 * tcs::NGROUPS * 128 * sizeof(float) + sizeof(tcs::Bars) + 128;
  const size_t budget = 227 * 1024;
  if (fixed + 2 * cfg.stage_bytes > budget) return "tensor-core step kernel: shared memory budget exceeded for this (S, A)";
  int nst = (int)((budget - fixed) / cfg.stage_bytes);
  if (nst > tcs::MAX_NST) nst = tcs::MAX_NST;
  cfg.nst = nst;
  const size_t bytes = fixed + (size_t)nst * cfg.stage_bytes;
  auto kern = ns == 2 ? tcs::step_tc_kernel<2> : tcs::step_tc_kernel<1>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return "cudaFuncSetAttribute(step_tc_kernel) failed";
  const int grid = (a.B + tcs::TM - 1) / tcs::TM;
  kern<<<grid, tcs::NTHREADS, bytes, st>>>(a, dynb, polb ? polb : dynb, sc, cfg);
  return nullptr;
}
