// Multi-GPU assembly of the synthetic transitions over peer memory.
//
// The reference is single-GPU; when rollout start states are sharded over the GPUs of a box, every rank needs all
// transitions in its own (device-resident) fake buffer.  Instead of an NCCL all-gather of padded slabs, a rank's rollout
// packs its kept transitions straight into slot `rank` of its OWN receive buffer, and a narrow streaming kernel on a
// high-priority side stream copies exactly those rows into the same slot of every peer's buffer through peer-mapped
// pointers (NVLink stores; one NVSwitch-replicated multicast store when the fabric offers it), then raises a flag there.
// Only kept rows travel, the copy overlaps the next rollout, and the consumer waits on the device.
//
// Receive buffer of one rank (mapped by all):  [half 0 | half 1 | flag block]
//   half h    = world slots; slot r = [cap_rows][W] packed transitions of rank r + one header row
//               header (int32 words): [0] kept rows, [1..2] produced transitions (int64), [3..4] reward sum (double)
//   flag block = unsigned arrive[8] (arrive[r] = last epoch rank r delivered here), unsigned ack[8] (ack[r] = last epoch
//               rank r has finished consuming in ITS buffer), unsigned ticket
#include "common.cuh"
#include <stdlib.h>
#include "../../include/mobody_b200.h"

namespace peer {

struct Layout { long long slot_floats, half_floats, flags_off; };
__host__ __device__ inline Layout layout(int world, long long cap_rows, int W) {
  Layout L;
  L.slot_floats = ((cap_rows + 1) * W + 3) & ~3LL;
  L.half_floats = (long long)world * L.slot_floats;
  L.flags_off = 2 * L.half_floats;
  return L;
}
constexpr int FLAG_WORDS = 32;   // arrive[8] ack[8] ticket pad
constexpr int ACK = 8, TICKET = 16;

struct PushArgs {
  const float* src;                   // local slot `rank` of half (epoch & 1): rows written by the rollout's pack kernel
  const int* m_dev; long long m_cap; const double* stats;
  float* dst[MOBODY_MAX_PEERS];       // the same slot at every rank (dst[rank] == src)
  float* mc_dst;                      // the same slot through the multicast mapping (NVLS: one store reaches every rank), or nullptr
  unsigned* flags[MOBODY_MAX_PEERS];  // flag block of every rank
  int world, rank; unsigned epoch; long long cap_rows; int W;
};

__device__ __forceinline__ void st_multimem(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Streams this rank's packed transitions (contiguous, 128-bit) from its own slot to the same slot of every peer over
// NVLink -- one unicast store per peer, or ONE multicast store replicated by the NVSwitch -- then publishes header + flag.
// Deliberately narrow (a few CTAs on a high-priority stream): it is NVLink bound and runs beside the next rollout's step kernel.
__global__ void __launch_bounds__(512) peer_push_kernel(const PushArgs a) {
  __shared__ bool last;
  // a peer's half (epoch & 1) may be overwritten once that peer has consumed epoch - 2 (its ack lands in OUR flag block)
  if ((int)threadIdx.x < a.world) {
    const volatile unsigned* ack = a.flags[a.rank] + ACK;
    const unsigned need = a.epoch >= 2u ? a.epoch - 2u : 0u;
    while (ack[threadIdx.x] < need) __nanosleep(64);
  }
  __syncthreads();
  const long long m = min((long long)*a.m_dev, a.m_cap);
  const long long n4 = (m * a.W + 3) >> 2;                        // the slot is padded to 16 bytes: the tail quad is in bounds
  const float4* src = reinterpret_cast<const float4*>(a.src);
  constexpr int U = 4;                                            // independent 128-bit loads in flight per thread
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < n4; t0 += U * stride) {
    float4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const long long t = t0 + u * stride; if (t < n4) q[u] = __ldcs(src + t); }
    if (a.mc_dst) {
#pragma unroll
      for (int u = 0; u < U; ++u) { const long long t = t0 + u * stride; if (t < n4) st_multimem(a.mc_dst + 4 * t, q[u]); }
    } else {
      for (int r = 0; r < a.world; ++r) {
        if (r == a.rank) continue;
        float4* d = reinterpret_cast<float4*>(a.dst[r]);
#pragma unroll
        for (int u = 0; u < U; ++u) { const long long t = t0 + u * stride; if (t < n4) d[t] = q[u]; }
      }
    }
  }
  // publish: every CTA's stores are released at system scope before it takes a ticket; the last CTA writes header + flag
  __threadfence_system();
  __syncthreads();
  unsigned* ticket = a.flags[a.rank] + TICKET;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if ((int)threadIdx.x < a.world) {
    const int r = threadIdx.x;
    int* h = reinterpret_cast<int*>(a.dst[r] + a.cap_rows * a.W);
    const long long produced = (long long)a.stats[1];
    h[0] = (int)m;
    h[1] = (int)(produced & 0xffffffffLL); h[2] = (int)(produced >> 32);
    h[3] = __double2loint(a.stats[0]); h[4] = __double2hiint(a.stats[0]);
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(a.flags[r] + a.rank) = a.epoch;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

// Header of this rank's slot, written LOCALLY (the copy-engine push ships the whole slot, header row included)
__global__ void peer_header_kernel(int* __restrict__ h, const int* __restrict__ m_dev, long long m_cap, const double* __restrict__ stats) {
  if (threadIdx.x == 0) {
    const long long m = min((long long)*m_dev, m_cap), produced = (long long)stats[1];
    h[0] = (int)m; h[1] = (int)(produced & 0xffffffffLL); h[2] = (int)(produced >> 32);
    h[3] = __double2loint(stats[0]); h[4] = __double2hiint(stats[0]);
  }
}

__global__ void peer_wait_kernel(const unsigned* __restrict__ flags_local, int world, unsigned epoch) {
  if ((int)threadIdx.x < world) {
    const volatile unsigned* f = flags_local;
    while (f[threadIdx.x] < epoch) __nanosleep(100);
  }
  __threadfence_system();
}

struct AckArgs { unsigned* flags[MOBODY_MAX_PEERS]; int world, rank; unsigned value; };
__global__ void peer_ack_kernel(const AckArgs a) {
  __threadfence_system();
  if ((int)threadIdx.x < a.world) *reinterpret_cast<volatile unsigned*>(a.flags[threadIdx.x] + ACK + a.rank) = a.value;
}

}  // namespace peer

static const char* check_peer(const mobody_peer_desc* p) {
  if (!p) return "null peer descriptor";
  if (p->world < 1 || p->world > MOBODY_MAX_PEERS || p->rank < 0 || p->rank >= p->world) return "bad world / rank";
  if (p->cap_rows < 1 || p->W < 5) return "bad cap_rows / W";
  for (int r = 0; r < p->world; ++r) if (!p->base[r]) return "null peer buffer";
  return nullptr;
}
static unsigned* flags_of(const mobody_peer_desc* p, int r) {
  return reinterpret_cast<unsigned*>(reinterpret_cast<float*>(p->base[r]) + peer::layout(p->world, p->cap_rows, p->W).flags_off);
}

long long mb_peer_slot_floats(long long cap_rows, int W) { return peer::layout(1, cap_rows, W).slot_floats; }
long long mb_peer_buffer_bytes(int world, long long cap_rows, int W) {
  return (peer::layout(world, cap_rows, W).flags_off + peer::FLAG_WORDS) * 4;
}
const char* mb_peer_slot(const mobody_peer_desc* p, int r, float** rows, int** header) {
  if (const char* e = check_peer(p)) return e;
  if (r < 0 || r >= p->world) return "bad slot index";
  const peer::Layout L = peer::layout(p->world, p->cap_rows, p->W);
  float* slot = reinterpret_cast<float*>(p->base[p->rank]) + (long long)(p->epoch & 1u) * L.half_floats + (long long)r * L.slot_floats;
  if (rows) *rows = slot;
  if (header) *header = reinterpret_cast<int*>(slot + p->cap_rows * p->W);
  return nullptr;
}
const char* mb_peer_header_launch(const mobody_peer_desc* p, const int* kept_dev, const double* stats_dev, cudaStream_t st) {
  if (const char* e = check_peer(p)) return e;
  if (!kept_dev || !stats_dev) return "null kept / stats pointer";
  float* rows; int* header;
  if (const char* e = mb_peer_slot(p, p->rank, &rows, &header)) return e;
  peer::peer_header_kernel<<<1, 32, 0, st>>>(header, kept_dev, p->cap_rows, stats_dev);
  return nullptr;
}
const char* mb_peer_ack_launch(const mobody_peer_desc* p, unsigned int consumed, cudaStream_t st) {
  if (const char* e = check_peer(p)) return e;
  peer::AckArgs a{}; a.world = p->world; a.rank = p->rank; a.value = consumed;
  for (int r = 0; r < p->world; ++r) a.flags[r] = flags_of(p, r);
  peer::peer_ack_kernel<<<1, 32, 0, st>>>(a);
  return nullptr;
}
const char* mb_peer_wait_launch(const mobody_peer_desc* p, cudaStream_t st) {
  if (const char* e = check_peer(p)) return e;
  peer::peer_wait_kernel<<<1, 32, 0, st>>>(flags_of(p, p->rank), p->world, p->epoch);
  return nullptr;
}
const char* mb_peer_push_launch(const mobody_peer_desc* p, const int* kept_dev, const double* stats_dev, cudaStream_t st) {
  if (const char* e = check_peer(p)) return e;
  if (!kept_dev || !stats_dev) return "null kept / stats pointer";
  const peer::Layout L = peer::layout(p->world, p->cap_rows, p->W);
  const long long slot_off = (long long)(p->epoch & 1u) * L.half_floats + (long long)p->rank * L.slot_floats;
  peer::PushArgs a{};
  a.m_dev = kept_dev; a.m_cap = p->cap_rows; a.stats = stats_dev;
  a.world = p->world; a.rank = p->rank; a.epoch = p->epoch; a.cap_rows = p->cap_rows; a.W = p->W;
  for (int r = 0; r < p->world; ++r) {
    a.dst[r] = reinterpret_cast<float*>(p->base[r]) + slot_off;
    a.flags[r] = flags_of(p, r);
  }
  a.src = a.dst[p->rank];
  a.mc_dst = p->multicast ? reinterpret_cast<float*>(p->multicast) + slot_off : nullptr;
  // A fused step kernel leaves ~5 KB of shared memory and ~10 K registers per SM: a push CTA of 128 threads (40 registers,
  // 1 KB of shared memory) fits BESIDE a step tile, a 512-thread one has to wait for a tile to finish and then holds that
  // SM's next tile back for the length of the push (MOBODY_PUSH_THREADS / MOBODY_PUSH_CTAS for A/B runs).
  static const int threads = [] { const char* e = getenv("MOBODY_PUSH_THREADS"); const int v = e ? atoi(e) : 0; return (v >= 32 && v <= 512) ? (v & ~31) : 128; }();
  int ctas = p->ctas > 0 ? p->ctas : 64;        // measured: 128 x 64 vs 512 x 8: N=2 162.2 vs 160.7, N=8 591.4 vs 588.5 M transitions/s
  if (ctas > 592) ctas = 592;
  peer::peer_push_kernel<<<ctas, threads, 0, st>>>(a);
  return nullptr;
}
