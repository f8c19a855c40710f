// Shared constants and helpers for the MOBODY B200 kernels.
// Layer order mirrors MOBODYModule.__init__ (reference algo/dynamics/mobody_module.py:97-184).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MB_E 7          // ensemble members (train_mobody.py:795; literal 7 in mobody_dynamics.py:218)
#define MB_H 256        // hidden width (train_mobody.py:794)
#define MB_LATENT 16    // latent_dim (mobody_module.py:95)
#define MB_STATS_BLOCKS 148   // cap on the partial-sum blocks of rollout_stats_kernel (sizes mobody_rollout_desc.stats)
#define MB_ZAH 32       // action-encoder hidden (mobody_module.py:110)

enum MbLayer {
  L_ZS1 = 0, L_ZS2, L_ZS3, L_ZASRC1, L_ZASRC2, L_ZATRG1, L_ZATRG2,
  L_T1, L_T2, L_T3, L_R1, L_R2, L_R3, L_COUNT
};

// Raw fp32 parameter pointers of the live nn.Parameters: weight [E,in,out], bias [E,1,out].
struct DynPtrs {
  const float* w[L_COUNT];
  const float* b[L_COUNT];
};
// nn.Linear layout: weight [out,in], bias [out]; three layers (mobody.py:35-48).
struct MlpPtrs {
  const float* w[3];
  const float* b[3];
};

struct StepArgs {
  const float* obs;        // [B,S]
  const float* act;        // [B,A] or nullptr -> computed by the policy MLP
  const float* eps;        // [E,B,S] injected N(0,1) (parity mode) or nullptr -> Philox
  const int64_t* idx;      // [B] injected member index per row or nullptr -> Philox pick among elites
  const int64_t* elites;   // [n_elites] device
  int n_elites;
  int B, S, A;             // B = row capacity (stride of eps/mean); live rows = *n_rows_dev if given
  int obs_ld, act_ld;      // row strides of obs / act in floats (>= S / A; packed buffer rows are read in place)
  const int* n_rows_dev;   // nullable: live row count on device (rollout steps after compaction)
  const long long* row_ids;// nullable: global row id per row (Philox counter); default row0 + r
  int use_trg, use_penalty, term_kind;
  float coef, max_action;
  unsigned long long seed, row0;   // Philox key / global row id of obs[0]
  unsigned int step;
  float* act_out;          // [B,A] or nullptr
  float* next_obs;         // [B,S]
  float* reward;           // [B]
  float* raw_reward;       // [B] or nullptr
  float* penalty;          // [B]
  unsigned char* terminal; // [B]
  float* mean;             // [E,B,S] (info['samples']); always written
};

// ---- programmatic dependent launch (PDL): the launch-bound kernel chains (train step, small rollouts) let the next
// kernel's CTAs become resident while this one drains.  Every kernel of a chain calls mb_pdl_begin() first:
// launch_dependents = "the next grid may start launching", wait = "block until every grid this one depends on has
// completed and its writes are visible" -- so no kernel touches global memory before its predecessors are done.
__device__ __forceinline__ void mb_pdl_begin() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#ifdef __CUDACC__
#include <cstdlib>
#include <utility>
inline bool mb_pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MOBODY_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t mb_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = mb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
#endif

__device__ __forceinline__ float mb_swish(float x) {
  // x * sigmoid(x); ex2.approx + rcp.approx are ~1-2 ulp, far inside the 1e-4 parity bound
  return __fdividef(x, 1.0f + __expf(-x));
}
