// Layout of the packed tensor-core weight blobs and the per-tile layer schedule.
// Shared by the pack kernels, the launcher and the rollout kernel (host + device).
//
// A weight layer W[K][N] (math: y = x W + b) is stored per 16-wide K step s, per split plane p
// (0 = bf16 hi, 1 = bf16 lo), as the exact shared-memory image the UMMA B operand wants:
//   K-major, no swizzle: [s][p][kgroup 0..1][n 0..Np)[8 consecutive k] bf16
// so one K step of all planes is one contiguous bulk copy of NS * Np * 32 bytes.
#pragma once
#include <stdint.h>

#define TC_MAX_LAYERS 80

enum TcRegion { REG_MAIN = 0, REG_OBS = 1, REG_SA = 2, REG_SAS = 3 };
enum TcKind {
  EPI_SWISH256 = 0,  // -> A_main (bf16 planes), swish           [zs1 zs2 transition1 transition2 reward_model1]
  EPI_RELU256 = 1,   // -> A_main, relu                          [policy 0, 2]
  EPI_ZS = 2,        // zs3 (mu half): keep zs, write [zs|act] operand
  EPI_G = 3,         // za1: swish -> 32-wide operand
  EPI_Z = 4,         // za2 (mu half): z = zs + za -> 16-wide operand
  EPI_MEAN = 5,      // transition3: + bias -> mean[e] (global)
  EPI_REWARD = 6,    // reward_model2: swish, dot with reward_model3[:,0]
  EPI_ACTION = 7     // policy 4: tanh * max_action -> act
};

struct TcLayer {
  uint32_t w_off;      // byte offset of the layer's weights from its blob base
  uint32_t b_off;      // float offset of the layer's bias from its blob's bias section
  uint16_t ksteps;     // Kp / 16
  uint16_t n;          // Np (multiple of 16, <= 256)
  uint8_t a_region;    // TcRegion of the A operand
  uint8_t a_wait;      // number of 32-column A chunks announced by the previous epilogue (0 = static)
  uint8_t kind;        // TcKind
  uint8_t blob;        // 0 = dynamics blob, 1 = policy blob
};

struct TcSched {
  int n_layers;
  int first_dyn;       // index of member 0's zs1 layer (3 with a policy, else 0)
  TcLayer L[TC_MAX_LAYERS];
};

static inline __host__ __device__ int tc_rup16(int x) { return (x + 15) & ~15; }

// ---- dynamics blob ----
enum { PK_ZS1 = 0, PK_ZS2, PK_ZS3, PK_ZASRC1, PK_ZASRC2, PK_ZATRG1, PK_ZATRG2, PK_T1, PK_T2, PK_T3, PK_R1, PK_R2, PK_COUNT };

struct TcGeom { int K, N, Kp, Np; };

static inline __host__ __device__ TcGeom tc_dyn_geom(int layer, int S, int A) {
  TcGeom g{0, 0, 0, 0};
  switch (layer) {
    case PK_ZS1: g.K = S; g.N = 256; break;
    case PK_ZS2: case PK_T2: case PK_R2: g.K = 256; g.N = 256; break;
    case PK_ZS3: g.K = 256; g.N = 16; break;                 // mu half of 32 columns
    case PK_ZASRC1: case PK_ZATRG1: g.K = 16 + A; g.N = 32; break;
    case PK_ZASRC2: case PK_ZATRG2: g.K = 32; g.N = 16; break; // mu half
    case PK_T1: g.K = 16; g.N = 256; break;
    case PK_T3: g.K = 256; g.N = S; break;
    case PK_R1: g.K = 2 * S + A; g.N = 256; break;
  }
  g.Kp = tc_rup16(g.K); g.Np = tc_rup16(g.N);
  return g;
}
static inline __host__ __device__ size_t tc_layer_bytes(const TcGeom& g, int ns) { return (size_t)(g.Kp / 16) * ns * g.Np * 32; }

// Swish layers are packed PRE-SCALED so the epilogue needs no multiply in front of the SFU op:
//   exact mode (ns = 2): accumulator holds t = -log2(e) x, epilogue returns t / (1 + 2^t) = -log2(e) swish(x)
//   loose mode (ns = 1): accumulator holds t = x / 2,      epilogue returns t + t tanh(t) = swish(x)
// `so` multiplies W and b of a layer whose output goes through swish, `si` multiplies W of a layer whose INPUT is
// such an epilogue result (si * so_epilogue_factor = 1: -ln 2 in exact mode, 1 in loose mode).
static inline __host__ __device__ void tc_swish_scales(int ns, double* so, double* si) {
  if (ns == 2) { *so = -1.4426950408889634; *si = -0.6931471805599453; } else { *so = 0.5; *si = 1.0; }
}
static inline __host__ __device__ bool tc_out_is_swish(int pk) {
  return pk == PK_ZS1 || pk == PK_ZS2 || pk == PK_ZASRC1 || pk == PK_ZATRG1 || pk == PK_T1 || pk == PK_T2 || pk == PK_R1 || pk == PK_R2;
}
static inline __host__ __device__ bool tc_in_is_swish(int pk) {
  return pk == PK_ZS2 || pk == PK_ZS3 || pk == PK_ZASRC2 || pk == PK_ZATRG2 || pk == PK_T2 || pk == PK_T3 || pk == PK_R2;
}

// Full (unpadded, both halves) shape of live parameter tensor `layer` (MbLayer order, common.cuh): weight [E][in][out], bias [E][1][out].
static inline __host__ __device__ void tc_dyn_full_dims(int layer, int S, int A, int* in, int* out) {
  const int I[13] = {S, 256, 256, 16 + A, 32, 16 + A, 32, 16, 256, 256, 2 * S + A, 256, 256};
  const int O[13] = {256, 256, 32, 32, 32, 32, 32, 256, 256, S, 256, 256, 2};
  *in = I[layer]; *out = O[layer];
}

#define TC_R3_FLOATS 272   // reward_model3 column 0 (256) + its bias at [256], padded

struct TcDynLayout {
  size_t w_off[PK_COUNT];       // within one member's weight block
  size_t member_w_bytes;
  uint32_t b_off[PK_COUNT + 1]; // float offsets within one member's bias block; [PK_COUNT] = reward_model3 vector
  uint32_t member_b_floats;
  size_t bias_base;             // byte offset of the bias section
  size_t total_bytes;
};

static inline __host__ __device__ TcDynLayout tc_dyn_layout(int S, int A, int ns) {
  TcDynLayout L{};
  size_t w = 0; uint32_t b = 0;
  for (int i = 0; i < PK_COUNT; ++i) {
    TcGeom g = tc_dyn_geom(i, S, A);
    L.w_off[i] = w; w += tc_layer_bytes(g, ns);
    L.b_off[i] = b; b += g.Np;
  }
  L.b_off[PK_COUNT] = b; b += TC_R3_FLOATS;
  L.member_w_bytes = w; L.member_b_floats = b;
  L.bias_base = 7 * w;
  L.total_bytes = L.bias_base + (size_t)7 * b * 4;
  return L;
}

// ---- policy blob: MLPNetwork S -> 256 -> 256 -> A ----
struct TcMlpLayout { size_t w_off[3]; uint32_t b_off[3]; TcGeom g[3]; size_t bias_base, total_bytes; };

static inline __host__ __device__ TcMlpLayout tc_mlp_layout(int din, int dout, int ns) {
  TcMlpLayout L{};
  int K[3] = {din, 256, 256}, N[3] = {256, 256, dout};
  size_t w = 0; uint32_t b = 0;
  for (int i = 0; i < 3; ++i) {
    L.g[i] = TcGeom{K[i], N[i], tc_rup16(K[i]), tc_rup16(N[i])};
    L.w_off[i] = w; w += tc_layer_bytes(L.g[i], ns);
    L.b_off[i] = b; b += L.g[i].Np;
  }
  L.bias_base = w; L.total_bytes = w + (size_t)b * 4;
  return L;
}
