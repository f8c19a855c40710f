// Termination predicates on next_obs (reference algo/mb_utils/terminal_funs.py:10-121).
// Kind ids are shared with include/mobody_b200.h and oracle.TERM_KINDS.  fp32 compares against
// the fp32-rounded thresholds: NumPy (both 1.x value-based casting and 2.x weak scalars) compares a
// float32 array with a Python float in float32, so the mask is reproduced bit for bit.
#pragma once
#include <math.h>

enum MbTerm { TERM_NEVER = 0, TERM_HALFCHEETAH = 1, TERM_HOPPER = 2, TERM_WALKER2D = 3, TERM_ANT = 4,
              TERM_HUMANOID = 5, TERM_PEN = 6 };

// x: one row of next_obs with element stride `stride` floats. Returns 1 if done.
__device__ __forceinline__ int mb_terminal(int kind, const float* x, int S, int stride = 1) {
  switch (kind) {
    case TERM_HALFCHEETAH: {            // :10-16  done = !(all x>-100 && all x<100); NaN -> done
      bool ok = true;
      for (int j = 0; j < S; ++j) { float v = x[j * stride]; ok = ok && (v > -100.0f) && (v < 100.0f); }
      return !ok;
    }
    case TERM_HOPPER: {                 // :18-30  no lower bound on dims>=1 (abs() is applied to the bool)
      bool ok = true;
      for (int j = 0; j < S; ++j) { float v = x[j * stride]; ok = ok && isfinite(v); if (j >= 1) ok = ok && (v < 100.0f); }
      float h = x[0], a = x[stride];
      ok = ok && (h > 0.7f) && (fabsf(a) < 0.2f);
      return !ok;
    }
    case TERM_WALKER2D: {               // :63-75
      bool ok = true;
      for (int j = 0; j < S; ++j) { float v = x[j * stride]; ok = ok && (v > -100.0f) && (v < 100.0f); }
      float h = x[0], a = x[stride];
      ok = ok && (h > 0.8f) && (h < 2.0f) && (a > -1.0f) && (a < 1.0f);
      return !ok;
    }
    case TERM_ANT: {                    // :39-61 (antangle and ant are identical)
      bool ok = true;
      for (int j = 0; j < S; ++j) ok = ok && isfinite(x[j * stride]);
      float h = x[0];
      ok = ok && (h >= 0.2f) && (h <= 1.0f);
      return !ok;
    }
    case TERM_HUMANOID: { float z = x[0]; return (z < 1.0f) || (z > 2.0f); }   // :98-104
    case TERM_PEN: return x[26 * stride] < 0.075f;                             // :106-113
    default: return 0;                                                         // never-terminating envs
  }
}
