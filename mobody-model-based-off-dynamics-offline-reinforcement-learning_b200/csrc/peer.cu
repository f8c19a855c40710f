// Multi-GPU assembly of the synthetic transitions over peer memory.
//
// The reference is single-GPU; when rollout start states are sharded over the GPUs of a box, every rank needs all
// transitions in its own (device-resident) fake buffer.  Instead of an NCCL all-gather of padded slabs, the PACK stage of
// the rollout writes the kept transitions straight into slot `rank` of every rank's receive buffer through peer-mapped
// pointers (NVLink stores), and raises a flag there.  No collective kernel competes with the step kernel for SMs, no
// padding travels, and the consumer waits on the device.
//
// Receive buffer of one rank (mapped by all):  [half 0 | half 1 | flag block]
//   half h    = world slots; slot r = [cap_rows][W] packed transitions of rank r + one header row
//               header (int32 words): [0] kept rows, [1..2] produced transitions (int64), [3..4] reward sum (double)
//   flag block = unsigned arrive[8] (arrive[r] = last epoch rank r delivered here), unsigned ack[8] (ack[r] = last epoch
//               rank r has finished consuming in ITS buffer), unsigned ticket
#include "common.cuh"
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
  const float *obss, *acts, *nexts, *rews, *pens; const unsigned char* terms; const int* pos; const int* m_dev; long long m_cap;
  const double* stats; int S, A;
  float* dst[MOBODY_MAX_PEERS];       // slot `rank` of half (epoch & 1) at every rank
  unsigned* flags[MOBODY_MAX_PEERS];  // flag block of every rank
  int world, rank; unsigned epoch; long long cap_rows; int W;
};

// One kernel = gather of the kept transitions (stable compaction order `pos`) + NVLink stores to every rank + header + flag.
__global__ void __launch_bounds__(512) rollout_pack_push_kernel(const PushArgs a) {
  __shared__ bool last;
  // a peer's half (epoch & 1) may be overwritten once that peer has consumed epoch - 2 (its ack lands in OUR flag block)
  if ((int)threadIdx.x < a.world) {
    const volatile unsigned* ack = a.flags[a.rank] + ACK;
    const unsigned need = a.epoch >= 2u ? a.epoch - 2u : 0u;
    while (ack[threadIdx.x] < need) __nanosleep(64);
  }
  __syncthreads();
  const int W = a.W, S = a.S, A = a.A;
  const long long m = min((long long)*a.m_dev, a.m_cap);
  const long long total = m * W, n4 = (total + 3) >> 2;
  for (long long t4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t4 < n4; t4 += (long long)gridDim.x * blockDim.x) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long t = t4 * 4 + k;
      float x = 0.f;
      if (t < total) {
        const long long j = t / W; const int c = (int)(t - j * W);
        const size_t p = (size_t)a.pos[j];
        if (c < S) x = a.obss[p * S + c];
        else if (c < S + A) x = a.acts[p * A + (c - S)];
        else if (c < 2 * S + A) x = a.nexts[p * S + (c - S - A)];
        else if (c == 2 * S + A) x = a.rews[p];
        else if (c == 2 * S + A + 1) x = (float)a.terms[p];
        else x = a.pens[p];
      }
      v[k] = x;
    }
    const float4 q = make_float4(v[0], v[1], v[2], v[3]);
    for (int r = 0; r < a.world; ++r) reinterpret_cast<float4*>(a.dst[r])[t4] = q;      // r == rank: the local copy
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
    int* h = reinterpret_cast<int*>(a.dst[r] + a.cap_rows * W);
    const long long produced = (long long)a.stats[1];
    h[0] = (int)m;
    h[1] = (int)(produced & 0xffffffffLL); h[2] = (int)(produced >> 32);
    h[3] = __double2loint(a.stats[0]); h[4] = __double2hiint(a.stats[0]);
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(a.flags[r] + a.rank) = a.epoch;
  }
  if (threadIdx.x == 0) *ticket = 0u;
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
const char* mb_rollout_push_launch(const mobody_rollout_desc* d, const mobody_peer_desc* p, cudaStream_t st) {
  if (const char* e = check_peer(p)) return e;
  const int S = d->step.S, A = d->step.A, T = d->T, B = d->step.B;
  if (p->W != 2 * S + A + 3) return "peer W does not match 2S+A+3";
  if ((long long)T * B > p->cap_rows) return "rollout capacity exceeds the peer slot";
  const peer::Layout L = peer::layout(p->world, p->cap_rows, p->W);
  peer::PushArgs a{};
  a.obss = d->obss; a.acts = d->acts; a.nexts = d->nexts; a.rews = d->rews; a.pens = d->pens; a.terms = d->terms; a.pos = d->pos;
  a.m_dev = d->counts + T + 1; a.m_cap = (long long)T * B; a.stats = d->stats; a.S = S; a.A = A;
  a.world = p->world; a.rank = p->rank; a.epoch = p->epoch; a.cap_rows = p->cap_rows; a.W = p->W;
  for (int r = 0; r < p->world; ++r) {
    a.dst[r] = reinterpret_cast<float*>(p->base[r]) + (long long)(p->epoch & 1u) * L.half_floats + (long long)p->rank * L.slot_floats;
    a.flags[r] = flags_of(p, r);
  }
  int ctas = p->ctas > 0 ? p->ctas : 24;
  if (ctas > 148) ctas = 148;
  peer::rollout_pack_push_kernel<<<ctas, 512, 0, st>>>(a);
  return nullptr;
}
