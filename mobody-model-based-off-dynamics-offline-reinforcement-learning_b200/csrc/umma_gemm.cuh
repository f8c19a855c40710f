// fp32-class GEMM tiles on the 5th-gen tensor cores: C[128 x N] (+)= A[128 x K] * B[K x N] with tcgen05.mma.kind::tf32,
// both operands split into TF32 hi + lo planes and three MMAs per K step (lo*hi + hi*lo + hi*hi: "3xTF32", ~2^-21 per
// product -- the precision the Q-weighted BC update needs), fp32 accumulation in TMEM.
//
// The train step at large batch is a handful of such GEMMs per layer (forward, backward-data, weight gradient) whose
// operands are fp32 tensors PyTorch owns: activations [rows][256], nn.Linear weights [out][in], batch rows.  They cannot be
// fed to the tensor core as they are (fp32 -> TF32 hi/lo needs ALU work), so a CTA's 8 warps stream K chunks of 32
// global -> registers (next chunk in flight) -> split -> shared memory in the UMMA K-major no-swizzle layout, and one thread
// issues the chunk's 12 MMAs asynchronously (tcgen05.commit -> mbarrier frees the stage): the tensor pipe works on chunk c
// while the CUDA cores split chunk c + 1.  Epilogues read the accumulator with tcgen05.ld and fuse bias / ReLU / ReLU mask /
// the one-output Q head / tanh.
//
// Operand sources (element (row, k) of a K-major operand plane; row = m for A, n for B):
//   SRC_KCONTIG : src[row * ld + k]      (activations, nn.Linear weight used "forward")
//   SRC_RCONTIG : src[k * ld + row]      (weight used "backward", and both operands of a weight gradient)
#pragma once
#include "common.cuh"
#include "tc_prims.cuh"

namespace ug {

constexpr int BM = 128;          // rows of C per CTA (UMMA M)
constexpr int KC = 32;           // K per pipeline stage = 4 UMMA K steps of 8
constexpr int NT = 256;          // threads per CTA

enum { SRC_KCONTIG = 0, SRC_RCONTIG = 1 };
enum { EPI_STORE = 0,   // C = act(acc + bias)                       -> C[m][n]
       EPI_MASK = 1,    // C = mask[m][n] > 0 ? acc : 0              -> C[m][n]          (ReLU backward)
       EPI_HEAD = 2,    // h = relu(acc + bias); out1[m] = h . w3 + b3; optional C = h     (Q network tail)
       EPI_TANH = 3,    // C = tanh(acc + bias) * scale              -> C[m][n]          (policy tail)
       EPI_PART = 4 };  // C[split][m][n] = acc, db[split][m] = row sums of the A operand (weight gradient partials)

struct Job {
  const float* A; int lda; const float* A2; int lda2; int ksplit;     // A2: columns k >= ksplit of a KCONTIG A come from A2[m][k - ksplit]
  const float* B; int ldb;
  int M, N, K;                                                          // N <= 256; operands are zero-padded to Np = roundup16(N), KC
  int a_src, b_src, epi, relu;
  const float* bias; const float* mask; int ldmask;
  const float* w3; const float* b3;
  float* C; int ldc; float* out1; float* db; float scale;
};
struct Args { Job job[8]; int njobs, nsplit; };

__host__ __device__ inline int rup16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline size_t stage_bytes(int np) { return (size_t)2 * (BM + np) * (KC / 4) * 16; }   // hi + lo planes of A and B

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xFFFFE000u;                               // exactly a TF32 value
  lo = __float_as_uint(v - __uint_as_float(hi)) + 0x1000u;             // the MMA truncates: pre-round the residual
}
__device__ __forceinline__ void put_unit(unsigned char* hi_plane, uint32_t plane_bytes, uint32_t off, const float (&v)[4]) {
  uint4 h, l;
  split_tf32(v[0], h.x, l.x); split_tf32(v[1], h.y, l.y); split_tf32(v[2], h.z, l.z); split_tf32(v[3], h.w, l.w);
  *reinterpret_cast<uint4*>(hi_plane + off) = h;
  *reinterpret_cast<uint4*>(hi_plane + plane_bytes + off) = l;
}
// kind::tf32 instruction descriptor: fp32 accumulate, TF32 A/B (format 2), both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// One operand (R rows x KC k) of one chunk: U = R / 32 units of (row, kgroup) per thread.  Mapping: a warp-iteration covers
// 8 rows x 4 kgroups (KCONTIG: each quarter-warp writes 128 contiguous bytes of shared memory, each row reads 64 contiguous
// bytes) or 32 rows x 1 kgroup (RCONTIG: 128-byte coalesced reads per k, 128-byte contiguous writes per quarter-warp).
template <int R>
struct Operand {
  static constexpr int U = R / 32;
  float v[U][4];

  __device__ __forceinline__ void load(const float* __restrict__ src, int ld, int mode, int row0, int rows, int k0, int kend,
                                       const float* __restrict__ src2, int ld2, int ksplit, float* rowsum) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (mode == SRC_KCONTIG) {
      const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && src2 == nullptr;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int b = warp * U + u;                                     // 8-row x 4-kgroup block
        const int row = (b >> 1) * 8 + (lane & 7), k = k0 + ((b & 1) * 4 + (lane >> 3)) * 4;
        const bool rok = row < rows;
        const float* p = src + (size_t)(row0 + row) * ld + k;
        if (vec && rok && k + 4 <= kend) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(p));
          v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x = 0.f;
            if (rok && k + i < kend) x = (src2 && k + i >= ksplit) ? __ldg(src2 + (size_t)(row0 + row) * ld2 + (k + i - ksplit)) : __ldg(p + i);
            v[u][i] = x;
          }
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int row = u * 32 + lane, k = k0 + warp * 4;               // warp = kgroup
        const bool rok = row < rows;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = (rok && k + i < kend) ? __ldg(src + (size_t)(k + i) * ld + row0 + row) : 0.f;
          v[u][i] = x; s += x;
        }
        if (rowsum) rowsum[u] += s;
      }
    }
  }
  __device__ __forceinline__ void store(unsigned char* hi_plane, int mode) const {
    constexpr uint32_t plane = (uint32_t)R * (KC / 4) * 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      int row, kg;
      if (mode == SRC_KCONTIG) { const int b = warp * U + u; row = (b >> 1) * 8 + (lane & 7); kg = (b & 1) * 4 + (lane >> 3); }
      else { row = u * 32 + lane; kg = warp; }
      put_unit(hi_plane, plane, (uint32_t)kg * (R * 16) + (uint32_t)row * 16, v[u]);
    }
  }
};

// NP = padded N of the tile (B rows staged): 256 for the hidden layers, 64 for the narrow ones.
template <int NP>
__global__ void __launch_bounds__(NT, 1) gemm_kernel(const __grid_constant__ Args args) {
  mb_pdl_begin();
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar_free[2], bar_done;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[8 * BM];
  const Job& jb = args.job[blockIdx.y];
  const int m0 = blockIdx.x * BM;
  if (m0 >= jb.M) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rows = min(BM, jb.M - m0);
  constexpr uint32_t A_PLANE = BM * (KC / 4) * 16, B_PLANE = NP * (KC / 4) * 16, STAGE = 2 * (A_PLANE + B_PLANE);
  constexpr uint32_t TCOLS = NP <= 32 ? 32 : NP <= 64 ? 64 : NP <= 128 ? 128 : 256;
  // K range of this CTA: whole K, or one of nsplit ranges of whole chunks (weight gradients)
  int kbeg = 0, kend = jb.K;
  if (args.nsplit > 1) {
    const long long chunks = (jb.K + KC - 1) / KC;
    kbeg = (int)(chunks * blockIdx.z / args.nsplit) * KC;
    kend = min(jb.K, (int)(chunks * (blockIdx.z + 1) / args.nsplit) * KC);
  }
  const int nchunks = kend > kbeg ? (kend - kbeg + KC - 1) / KC : 0;

  if (tid == 0) { tc::mbar_init(&bar_free[0], 1); tc::mbar_init(&bar_free[1], 1); tc::mbar_init(&bar_done, 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TCOLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  Operand<BM> oa; Operand<NP> ob;
  float rowsum[Operand<BM>::U];
#pragma unroll
  for (int u = 0; u < Operand<BM>::U; ++u) rowsum[u] = 0.f;
  float* rs = (jb.epi == EPI_PART && jb.db) ? rowsum : nullptr;
  const int nrows_b = jb.N;                                             // valid rows of the B operand (n < N)
  if (nchunks > 0) {
    oa.load(jb.A, jb.lda, jb.a_src, m0, rows, kbeg, kend, jb.A2, jb.lda2, jb.ksplit, rs);
    ob.load(jb.B, jb.ldb, jb.b_src, 0, nrows_b, kbeg, kend, nullptr, 0, 0, nullptr);
  }
  const uint32_t idesc = make_idesc_tf32(BM, NP);
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    unsigned char* st = smem + (size_t)buf * STAGE;
    if (c >= 2) tc::mbar_wait(&bar_free[buf], (uint32_t)(((c >> 1) - 1) & 1));      // the MMAs that read this stage are done
    oa.store(st, jb.a_src);
    ob.store(st + 2 * A_PLANE, jb.b_src);
    if (c + 1 < nchunks) {                                              // next chunk's global loads fly during this chunk's MMAs
      const int k0 = kbeg + (c + 1) * KC;
      oa.load(jb.A, jb.lda, jb.a_src, m0, rows, k0, kend, jb.A2, jb.lda2, jb.ksplit, rs);
      ob.load(jb.B, jb.ldb, jb.b_src, 0, nrows_b, k0, kend, nullptr, 0, 0, nullptr);
    }
    tc::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t a0 = tc::smem_u32(st), b0 = a0 + 2 * A_PLANE;
#pragma unroll
      for (int s = 0; s < KC / 8; ++s) {
        const uint32_t ao = a0 + (uint32_t)s * 2u * (BM * 16), bo = b0 + (uint32_t)s * 2u * (NP * 16);
        const uint64_t ah = tc::make_smem_desc(ao, BM * 16, 128), al = tc::make_smem_desc(ao + A_PLANE, BM * 16, 128);
        const uint64_t bh = tc::make_smem_desc(bo, NP * 16, 128), bl = tc::make_smem_desc(bo + B_PLANE, NP * 16, 128);
        umma_tf32(tmem, al, bh, idesc, (c > 0 || s > 0) ? 1u : 0u);
        umma_tf32(tmem, ah, bl, idesc, 1u);
        umma_tf32(tmem, ah, bh, idesc, 1u);
      }
      tc::umma_commit(&bar_free[buf]);
      if (c + 1 == nchunks) tc::umma_commit(&bar_done);
    }
  }
  if (nchunks > 0) tc::mbar_wait(&bar_done, 0);
  tc::tc_fence_after();

  // ---------------- epilogue: warp w reads TMEM lane quadrant w & 3, column half w >> 2 ----------------
  const int q = warp & 3, half = warp >> 2;
  const int r = q * 32 + lane, m = m0 + r;
  const bool valid = r < rows;
  constexpr int CH = 16;                                                // columns per tcgen05.ld
  constexpr int NCH = NP / CH;
  const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
  float head = 0.f;
  for (int ch = half; ch < NCH; ch += 2) {
    uint32_t x[CH];
    if (nchunks > 0) { tc::tmem_ld16(taddr + (uint32_t)(ch * CH), x); tc::tmem_ld_wait(); }
    else {
#pragma unroll
      for (int j = 0; j < CH; ++j) x[j] = 0u;
    }
    const int n0 = ch * CH;
    if (jb.epi == EPI_PART) {
      if (valid) {
        float* dst = jb.C + ((size_t)blockIdx.z * jb.M + m) * jb.N + n0;
#pragma unroll
        for (int j = 0; j < CH; ++j) if (n0 + j < jb.N) dst[j] = __uint_as_float(x[j]);
      }
      continue;
    }
    float v[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int n = n0 + j;
      float t = __uint_as_float(x[j]);
      if (jb.bias && n < jb.N) t += __ldg(jb.bias + n);
      if (jb.epi == EPI_MASK) t = (valid && n < jb.N && __ldg(jb.mask + (size_t)m * jb.ldmask + n) > 0.f) ? t : 0.f;
      else if (jb.epi == EPI_TANH) t = tanhf(t) * jb.scale;
      else if (jb.relu) t = fmaxf(t, 0.f);
      v[j] = t;
    }
    if (jb.epi == EPI_HEAD) {
#pragma unroll
      for (int j = 0; j < CH; ++j) if (n0 + j < jb.N) head = fmaf(v[j], __ldg(jb.w3 + n0 + j), head);
    }
    if (jb.C && valid) {
      float* dst = jb.C + (size_t)m * jb.ldc + n0;
      if (n0 + CH <= jb.N && (jb.ldc & 3) == 0) {
#pragma unroll
        for (int j = 0; j < CH; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < CH; ++j) if (n0 + j < jb.N) dst[j] = v[j];
      }
    }
  }
  if (jb.epi == EPI_HEAD) {                                             // the two column halves of a row meet in shared memory
    red[half * BM + r] = head;
    __syncthreads();
    if (half == 0 && valid) jb.out1[m] = red[r] + red[BM + r] + __ldg(jb.b3);
  }
  if (jb.epi == EPI_PART && jb.db && jb.a_src == SRC_RCONTIG) {         // bias gradient: row sums of A (= column sums of D), fixed order
#pragma unroll
    for (int u = 0; u < Operand<BM>::U; ++u) red[warp * BM + u * 32 + lane] = rowsum[u];
    __syncthreads();
    if (tid < rows) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w * BM + tid];
      jb.db[(size_t)blockIdx.z * jb.M + m0 + tid] = s;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

}  // namespace ug
