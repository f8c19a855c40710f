// Large-batch MOBODY.train step on tcgen05 (3xTF32 GEMM tiles, umma_gemm.cuh).  See the orchestration below.
#include "umma_gemm.cuh"
#include "../../include/mobody_b200.h"

template <int NP> static const char* launch_gemm(const ug::Args& a, int max_m, cudaStream_t st) {
  const size_t bytes = 2 * ug::stage_bytes(NP);
  if (cudaFuncSetAttribute(ug::gemm_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return "cudaFuncSetAttribute(gemm_kernel) failed";
  const dim3 grid((max_m + ug::BM - 1) / ug::BM, a.njobs, a.nsplit);
  if (mb_launch(ug::gemm_kernel<NP>, grid, dim3(ug::NT), bytes, st, a) != cudaSuccess) return "gemm_kernel launch failed";
  return nullptr;
}
// one launch: every job of `a` must fit the chosen tile width
const char* mb_gemm_launch(const ug::Args& a, cudaStream_t st) {
  int max_m = 0, max_n = 0;
  for (int j = 0; j < a.njobs; ++j) {
    const ug::Job& jb = a.job[j];
    if (jb.M < 1 || jb.N < 1 || jb.N > 256 || jb.K < 0) return "gemm: bad M / N / K";
    if (jb.M > max_m) max_m = jb.M;
    if (jb.N > max_n) max_n = jb.N;
  }
  if (max_n <= 64) return launch_gemm<64>(a, max_m, st);
  return launch_gemm<256>(a, max_m, st);
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
