// Large-batch MOBODY.train step on tcgen05 (3xTF32 GEMM tiles, umma_gemm.cuh).  See the orchestration below.
#include "umma_gemm.cuh"
#include "../../include/mobody_b200.h"

template <int NP, int AS, int BS, int EPI> static const char* launch_gemm(const ug::Args& a, int max_m, cudaStream_t st) {
  const size_t bytes = 2 * ug::stage_bytes(NP);
  static bool attr_set = false;            // per instantiation
  if (!attr_set) {
    if (cudaFuncSetAttribute(ug::gemm_kernel<NP, AS, BS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
      return "cudaFuncSetAttribute(gemm_kernel) failed";
    attr_set = true;
  }
  const dim3 grid((max_m + ug::BM - 1) / ug::BM, a.njobs, a.nsplit);
  if (mb_launch(ug::gemm_kernel<NP, AS, BS, EPI>, grid, dim3(ug::NT), bytes, st, a) != cudaSuccess) return "gemm_kernel launch failed";
  return nullptr;
}
template <int AS, int BS, int EPI> static const char* launch_np(const ug::Args& a, int max_m, bool narrow, cudaStream_t st) {
  return narrow ? launch_gemm<64, AS, BS, EPI>(a, max_m, st) : launch_gemm<256, AS, BS, EPI>(a, max_m, st);
}
// one launch: every job of `a` must use the same operand sources and epilogue class, and fit the chosen tile width
const char* mb_gemm_launch(const ug::Args& a, cudaStream_t st) {
  int max_m = 0, max_n = 0;
  for (int j = 0; j < a.njobs; ++j) {
    const ug::Job& jb = a.job[j];
    if (jb.M < 1 || jb.N < 1 || jb.N > 256 || jb.K < 0) return "gemm: bad M / N / K";
    if (jb.a_src != a.job[0].a_src || jb.b_src != a.job[0].b_src || ug::epi_class(jb.epi) != ug::epi_class(a.job[0].epi))
      return "gemm: the jobs of one launch must share their operand sources and epilogue class";
    if (jb.epi == ug::EPI_HEAD && !jb.w3) return "gemm: head job without w3";
    if (jb.M > max_m) max_m = jb.M;
    if (jb.N > max_n) max_n = jb.N;
  }
  const int as = a.job[0].a_src, bs = a.job[0].b_src, ec = ug::epi_class(a.job[0].epi);
  const bool narrow = max_n <= 64;
  using namespace ug;
  if (as == SRC_KCONTIG && bs == SRC_KCONTIG) {          // forward layers
    if (ec == EPI_STORE) return launch_np<0, 0, EPI_STORE>(a, max_m, narrow, st);
    if (ec == EPI_TANH) return launch_np<0, 0, EPI_TANH>(a, max_m, narrow, st);
  } else if (as == SRC_KCONTIG && bs == SRC_RCONTIG) {   // backward-data
    if (ec == EPI_MASK) return launch_np<0, 1, EPI_MASK>(a, max_m, narrow, st);
    if (ec == EPI_STORE) return launch_np<0, 1, EPI_STORE>(a, max_m, narrow, st);
  } else if (as == SRC_RCONTIG && bs == SRC_RCONTIG) {   // weight gradients (and the test hook)
    if (ec == EPI_PART) return launch_np<1, 1, EPI_PART>(a, max_m, narrow, st);
    if (ec == EPI_STORE) return launch_np<1, 1, EPI_STORE>(a, max_m, narrow, st);
  } else if (ec == EPI_STORE) return launch_np<1, 0, EPI_STORE>(a, max_m, narrow, st);
  return "gemm: this operand-source / epilogue combination is not instantiated";
}

// Test hook (C ABI: mobody_selftest_gemm): C[M][N] = A * B through the tile kernel, any operand source combination.
const char* mb_gemm_selftest_launch(const float* A, const float* B, int M, int N, int K, int a_src, int b_src, int lda, int ldb,
                                    float* C, cudaStream_t st) {
  ug::Args a{}; a.njobs = 1; a.nsplit = 1;
  ug::Job& j = a.job[0];
  j.A = A; j.lda = lda; j.B = B; j.ldb = ldb; j.M = M; j.N = N; j.K = K; j.a_src = a_src; j.b_src = b_src;
  j.epi = ug::EPI_STORE; j.C = C; j.ldc = N;
  return mb_gemm_launch(a, st);
}
