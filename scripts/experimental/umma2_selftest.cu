// Experimental (not part of the product library): CTA-pair UMMA self test that accompanied step_pair.cu.
#include "../../mobody-model-based-off-dynamics-offline-reinforcement-learning_b200/csrc/tc_prims.cuh"
namespace tcst {
// CTA-pair variant: D[256,N] = A[256,K] * B[N,K]^T with tcgen05.mma.cta_group::2.
// CTA r stages A rows [128r, 128r+128) and B rows [r N/2, (r+1) N/2) in its own shared memory.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
umma2_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int K, int N, int nsplit, float* __restrict__ D) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int nh = N / 2;
  const uint32_t a_plane = (uint32_t)(K / 8) * 2048u, b_plane = (uint32_t)(K / 8) * (uint32_t)nh * 16u;
  unsigned char* sA = smem;
  unsigned char* sB = smem + 2 * a_plane;
  for (int k = 0; k < K; ++k) {
    float v = A[(size_t)(rank * 128 + tid) * K + k];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    *reinterpret_cast<__nv_bfloat16*>(sA + op_off(tid, k, 128)) = h;
    *reinterpret_cast<__nv_bfloat16*>(sA + a_plane + op_off(tid, k, 128)) = l;
  }
  for (int n = tid; n < nh; n += 128)
    for (int k = 0; k < K; ++k) {
      float v = B[(size_t)(rank * nh + n) * K + k];
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      *reinterpret_cast<__nv_bfloat16*>(sB + op_off(n, k, nh)) = h;
      *reinterpret_cast<__nv_bfloat16*>(sB + b_plane + op_off(n, k, nh)) = l;
    }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc2(&tmem_slot, 256);
  tc::fence_proxy_async_all();
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (rank == 0 && warp == 1 && tc::elect_one()) {
    const uint32_t idesc = tc::make_idesc_bf16(256, N);
    const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
    uint32_t acc = 0;
    for (int s = 0; s < K / 16; ++s) {
      const uint32_t ao = a0 + s * 4096u, bo = b0 + s * (uint32_t)nh * 32u;
      uint64_t ah = tc::make_smem_desc(ao, 2048, 128), al = tc::make_smem_desc(ao + a_plane, 2048, 128);
      uint64_t bh = tc::make_smem_desc(bo, nh * 16, 128), bl = tc::make_smem_desc(bo + b_plane, nh * 16, 128);
      tc::umma2_bf16(tmem, ah, bh, idesc, acc); acc = 1;
      if (nsplit == 2) { tc::umma2_bf16(tmem, al, bh, idesc, 1); tc::umma2_bf16(tmem, ah, bl, idesc, 1); }
    }
    tc::umma_commit2(&bar, 3);
  }
  __syncwarp();
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    uint32_t r[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[(size_t)(rank * 128 + warp * 32 + lane) * N + c + j] = __uint_as_float(r[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync();
  if (warp == 0) tc::tmem_dealloc2(tmem, 256);
}

}  // namespace tcst

const char* mb_umma2_selftest_launch(const float* A, const float* B, int K, int N, int nsplit, float* D, cudaStream_t st) {
  if (K < 16 || K % 16 || K > 256 || N < 16 || N % 16 || N > 256 || (nsplit != 1 && nsplit != 2)) return "selftest2: bad K/N/nsplit";
  size_t bytes = 2 * (size_t)(K / 8) * 2048 + 2 * (size_t)(K / 8) * (N / 2) * 16;
  if (bytes > 200 * 1024) return "selftest2: operands exceed shared memory";
  if (cudaFuncSetAttribute(tcst::umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
    return "selftest2: cudaFuncSetAttribute failed";
  tcst::umma2_selftest_kernel<<<2, 128, bytes, st>>>(A, B, K, N, nsplit, D);
  return nullptr;
}

