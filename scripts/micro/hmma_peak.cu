// Microbenchmark: issue rate of the legacy warp-level tensor-core path on sm_100a (mma.sync TF32 m16n8k8 and BF16
// m16n8k16), independent accumulators, W warps per SM sub-partition.  Prints dense TFLOP/s per instruction shape.
// Evidence for DESIGN.md section 4.5 (ceiling of the 3xTF32 train-step kernels).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND, int NACC>
__global__ void k(float* out, int iters) {
  float c[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  uint32_t a[4] = {threadIdx.x, 2u, 3u, 4u}, b[2] = {5u, 6u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456f) out[0] = s;
}
template <int KIND> void run(const char* name, double flop_per_inst) {
  float* out; cudaMalloc(&out, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int warps : {4, 8, 16, 32}) {
    const int iters = 20000; constexpr int NACC = 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, NACC><<<sms, warps * 32>>>(out, 100);
    cudaEventRecord(e0);
    k<KIND, NACC><<<sms, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double insts = (double)sms * warps * iters * NACC;
    printf("%s warps/SM=%2d: %.1f TFLOP/s dense, %.2f cycles per instruction per SM sub-partition at 1.9 GHz\n", name, warps,
           insts * flop_per_inst / (ms * 1e-3) / 1e12, (ms * 1e-3) * 1.9e9 / (insts / sms / 4));
  }
}
int main() { run<0>("tf32 m16n8k8 ", 2048.0); run<1>("bf16 m16n8k16", 4096.0); return 0; }
