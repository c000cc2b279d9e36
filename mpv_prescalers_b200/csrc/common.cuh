// Shared host/device helpers for libmpvp (sm_100a).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <type_traits>

#include "../../include/mpvp.h"

namespace mpvp {

// ---- error reporting ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define MPVP_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::mpvp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MPVP_E_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define MPVP_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::mpvp::set_error(__VA_ARGS__);    \
      return MPVP_E_INVALID;             \
    }                                    \
  } while (0)

// RAII device switch: launches run on the device that owns the weights.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int sm_count(int device);

// Test hook (mpvp_debug_set_grid_limit): caps the number of persistent CTAs of every launch so that small planes make
// each CTA / warpgroup walk many tiles -- the steady state of the tile loops (buffer hand-over, mbarrier phase flips,
// TMEM slot reuse) then runs at sizes the CPU oracle checks in seconds.  0 = no cap (the product setting).
extern std::atomic<int> g_grid_limit;
inline long long cap_grid(long long grid) {
  const int l = g_grid_limit.load(std::memory_order_relaxed);
  return (l > 0 && grid > l) ? (long long)l : grid;
}

// The sqrt/division-free key (key_from_abd_fast) reads the DERIVED tables l1_thr[] / coh_ratio[] of mpvp_key_params;
// a caller that filled only the shader constants must run mpvp_key_params_finalize() first.  MPVP_OK or MPVP_E_INVALID.
int check_fast_key(const mpvp_key_params* key);

}  // namespace mpvp

// Opaque handle behind mpvp_weights*.
struct mpvp_weights {
  int device = 0;
  int kind = 0;  // 0 = LUT, 1 = NNEDI3
  // LUT: rows x texels_per_row float4 (values already rounded to fp16 precision if requested)
  float* lut = nullptr;
  void* lut_half = nullptr;  // same texels as 4 x binary16 (exact when the LUT was rounded to fp16): halves L2 traffic
  cudaArray_t lut_arr = nullptr;        // ravu-zoom LUTs only: the binary16 texels as a 2-D array behind a texture object
  cudaTextureObject_t lut_tex = 0;      // FILTER LINEAR, clamp-to-edge, unnormalised coordinates
  int lut_w = 0, lut_h = 0;
  // ravu-zoom: per-geometry phase plans (member tables + pre-blended phase LUTs), owned by ravu_zoom.cu
  void* zoom_plans = nullptr;
  void (*zoom_plans_free)(void*) = nullptr;
  // NNEDI3
  int nns = 0, win_short = 0;
  int nn_group = 16;        // neurons per accumulator block in nn_b (16: 32-column blocks, 8: 16-column blocks)
  void* nn_b = nullptr;     // packed fp16 B operand(s) for tcgen05
  float* nn_bias = nullptr; // [2*nns] interleaved (b1*log2e, b2)
  float* nn_w = nullptr;    // fp32 copy [2*nns][K] (reference SIMT path / debugging)
  size_t nn_b_bytes = 0;
};

namespace mpvp {

// ---- RAVU key (structure tensor -> LUT row) --------------------------------------------------
// Every operation below is an explicitly rounded IEEE fp32 op in the operand order of the shader
// text (ravu-lite-ar-r3.hook:48-88, ravu-r3.hook:58-119): no FMA contraction, correctly rounded
// sqrt / division.  The >= 99.99 % bucket-agreement rule leaves no room for re-association
// (SURVEY.md section 7.4 item 2, App. H16).

constexpr float kEps = 1.192092896e-7f;
constexpr float kPi = 3.141592653589793f;

enum Stencil { STENCIL_LITE = 0, STENCIL_RAVU = 1 };

// Correctly rounded x / 12 in three FMA-pipe instructions instead of the generic IEEE division sequence:
// q = x * fl(1/12); r = x - 12 q (exact, one FMA); q' = q + r * fl(1/12).  Verified equal to the correctly
// rounded quotient for all 2^24 mantissas; the 4th-order stencil needs the *rounded* /12
// of the shader, a bare multiplication by 1/12 flips buckets.
__device__ __forceinline__ float div12_rn(float x) {
  const float z = 0.0833333358168602f;
  const float q = __fmul_rn(x, z);
  const float r = __fmaf_rn(-q, 12.0f, x);
  return __fmaf_rn(r, z, q);
}

// One finite difference along an axis; S(d) returns the sample at offset d along that axis.
template <int FAMILY, int N, class SF>
__device__ __forceinline__ float key_diff(int k, SF S) {
  if (FAMILY == STENCIL_RAVU && k - 2 >= 0 && k + 2 <= N - 1) {
    // (-s[+2] + 8.0*s[+1] - 8.0*s[-1] + s[-2]) / 12.0, left to right.  8.0*s is exact in binary floating point, so
    // fma(8, s[+1], -s[+2]) rounds exactly once, like the shader's add of the exact product, and so does the
    // second term: two FFMA replace two FMUL + two FADD with bit-identical results.
    float t = __fmaf_rn(8.0f, S(1), -S(2));
    t = __fmaf_rn(-8.0f, S(-1), t);
    t = __fadd_rn(t, S(-2));
    return div12_rn(t);
  }
  if (k - 1 >= 0 && k + 1 <= N - 1) return __fmul_rn(__fsub_rn(S(1), S(-1)), 0.5f);
  if (k - 1 < 0) return __fsub_rn(S(1), S(0));
  return __fsub_rn(S(0), S(-1));
}

// The same finite differences selected by an explicit stencil kind (0: 4th order, 1: central, 2: forward, 3: backward)
template <int N, class SF>
__device__ __forceinline__ float key_diff_kind(int kind, SF S) {
  if (kind == 0) {
    float t = __fmaf_rn(8.0f, S(1), -S(2));
    t = __fmaf_rn(-8.0f, S(-1), t);
    t = __fadd_rn(t, S(-2));
    return div12_rn(t);
  }
  if (kind == 1) return __fmul_rn(__fsub_rn(S(1), S(-1)), 0.5f);
  if (kind == 2) return __fsub_rn(S(1), S(0));
  return __fsub_rn(S(0), S(-1));
}

struct KeyOut {
  int row;
};

// a, b, d -> LUT row.  NTHR = number of strength thresholds (0 => log2 form).
__device__ __forceinline__ int key_from_abd(const mpvp_key_params& kp, float a, float b, float d) {
  const float T = __fadd_rn(a, d);
  const float D = __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, b));
  const float delta = __fsqrt_rn(fmaxf(__fsub_rn(__fmul_rn(__fmul_rn(T, T), 0.25f), D), 0.0f));
  const float halfT = __fmul_rn(T, 0.5f);
  const float L1 = __fadd_rn(halfT, delta);
  const float L2 = __fsub_rn(halfT, delta);
  const float sqrtL1 = __fsqrt_rn(L1);
  const float sqrtL2 = __fsqrt_rn(L2);  // NaN for a tiny negative L2, deliberately kept (App. D.2)
  float theta;
  if (fabsf(b) < kEps) {
    theta = 0.0f;
  } else {
    const float at = __fadd_rn(atan2f(__fsub_rn(L1, a), b), kPi);
    theta = __fsub_rn(at, __fmul_rn(kPi, floorf(__fdiv_rn(at, kPi))));
  }
  const float ssum = __fadd_rn(sqrtL1, sqrtL2);
  float mu = __fdiv_rn(__fsub_rn(sqrtL1, sqrtL2), ssum);
  if (ssum < kEps) mu = 0.0f;
  const float angle = floorf(__fdiv_rn(__fmul_rn(theta, 24.0f), kPi));
  float strength;
  if (kp.n_strength_thr > 0) {
    strength = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
      if (i < kp.n_strength_thr && sqrtL1 >= kp.strength_thr[i]) strength += 1.0f;
  } else {
    strength = floorf(log2f(__fadd_rn(__fmul_rn(sqrtL1, kp.strength_log2_scale), kEps)));
    strength = fminf(fmaxf(strength, 0.0f), (float)(kp.n_strength - 1));
  }
  const float coh = (mu >= kp.coherence_thr[0] ? 1.0f : 0.0f) + (mu >= kp.coherence_thr[1] ? 1.0f : 0.0f);
  float rowf = (angle * (float)kp.n_strength + strength) * 3.0f + coh;
  const int nrows = 24 * kp.n_strength * 3;
  if (!(rowf >= 0.0f)) rowf = 0.0f;  // also catches NaN
  int row = (int)rowf;
  return row > nrows - 1 ? nrows - 1 : row;
}

// Full key for an N x N window whose sample (i, j) [i <-> dx, j <-> dy] is W(i, j).
template <int FAMILY, int N, int G, class WF>
__device__ __forceinline__ int ravu_key(const mpvp_key_params& kp, WF W) {
  constexpr int O = (N - G) / 2;
  float a = 0.0f, b = 0.0f, d = 0.0f;
#pragma unroll
  for (int i = O; i < O + G; ++i) {
#pragma unroll
    for (int j = O; j < O + G; ++j) {
      const float gx = key_diff<FAMILY, N>(i, [&](int dd) { return W(i + dd, j); });
      const float gy = key_diff<FAMILY, N>(j, [&](int dd) { return W(i, j + dd); });
      const float g = kp.gauss[(i - O) * G + (j - O)];
      a = __fadd_rn(a, __fmul_rn(__fmul_rn(gx, gx), g));
      b = __fadd_rn(b, __fmul_rn(__fmul_rn(gx, gy), g));
      d = __fadd_rn(d, __fmul_rn(__fmul_rn(gy, gy), g));
    }
  }
  return key_from_abd(kp, a, b, d);
}

// Packed f32x2 multiply / add with explicit .rn rounding.  CAUTION (ptxas 12.9): a mul.rn.f32x2 feeding an
// add.rn.f32x2 IS contracted into one FFMA2 (unlike the scalar mul.rn.f32 / add.rn.f32 pair, and regardless of
// -fmad=false), which the bucket-parity rule forbids: keep the additions of the key sums scalar.
__device__ __forceinline__ uint64_t pack2(float2 v) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 unpack2(uint64_t r) {
  float2 v;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
  return v;
}
__device__ __forceinline__ float2 mul2_rn(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2(a)), "l"(pack2(b)));
  return unpack2(d);
}
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2(a)), "l"(pack2(b)));
  return unpack2(d);
}

// ---- the same key without sqrt / division / atan2 ------------------------------------------------------
// Up to (L1, L2) the arithmetic is the shader's, op for op.  The three quantisers are then evaluated on
// quantities that decide identically except within fp32 rounding noise of a bucket edge:
//   strength : L1 >= l1_thr[i]              (exactly equivalent: thresholds found by bisection on the host)
//   coherence: L1 >= L2 * ((1+c)/(1-c))^2   (<=> (sqrt L1 - sqrt L2)/(sqrt L1 + sqrt L2) >= c)
//   angle    : sector of the direction (b, L1 - a) by octant folding + 5 slope comparisons
//              (<=> floor(mod(atan(L1 - a, b) + pi, pi) * 24 / pi))
// Degenerate inputs keep the shader's behaviour: |b| < eps -> angle 0; sqrt(L2 < 0) = NaN -> coherence 0;
// sqrt L1 + sqrt L2 < eps -> coherence 0.
template <int NTHR>
__device__ __forceinline__ int key_from_abd_fast(const mpvp_key_params& kp, float a, float b, float d) {
  const float T = __fadd_rn(a, d);
  const float D = __fsub_rn(__fmul_rn(a, d), __fmul_rn(b, b));
  const float delta = __fsqrt_rn(fmaxf(__fsub_rn(__fmul_rn(__fmul_rn(T, T), 0.25f), D), 0.0f));
  const float halfT = __fmul_rn(T, 0.5f);
  const float L1 = __fadd_rn(halfT, delta);
  const float L2 = __fsub_rn(halfT, delta);
  int strength = 0;
#pragma unroll
  for (int i = 0; i < NTHR; ++i) strength += (L1 >= kp.l1_thr[i]) ? 1 : 0;
  int coh;
  if (L1 < 3.5e-15f || L2 < 0.0f) {
    coh = 0;  // sqrt(L1) + sqrt(L2) <= 2 sqrt(L1) < eps, or sqrt(L2) is NaN: both comparisons are false
  } else if (L1 < 1.5e-14f) {
    const float s1 = __fsqrt_rn(L1), s2 = __fsqrt_rn(L2);
    const float ssum = __fadd_rn(s1, s2);
    const float mu = ssum < kEps ? 0.0f : __fdiv_rn(__fsub_rn(s1, s2), ssum);
    coh = (mu >= kp.coherence_thr[0] ? 1 : 0) + (mu >= kp.coherence_thr[1] ? 1 : 0);
  } else {
    coh = (L1 >= L2 * kp.coh_ratio[0] ? 1 : 0) + (L1 >= L2 * kp.coh_ratio[1] ? 1 : 0);
  }
  int angle = 0;
  float Y = __fsub_rn(L1, a);
  // Y == 0 (L1 - a cancels completely on near axis-aligned edges): atan(0, b) is 0 or pi, and the shader's
  // fp32 mod(pi + pi, pi) wraps both to theta = 0
  if (!(fabsf(b) < kEps) && Y != 0.0f) {
    float X = b;
    if (Y < 0.0f) { X = -X; Y = -Y; }       // theta is taken mod pi
    const bool neg = X < 0.0f;               // mirror about the y axis: sector s -> 23 - s
    X = fabsf(X);
    const bool sw = Y > X;                   // reflect about 45 degrees: sector s -> 11 - s
    const float lo = sw ? X : Y, hi = sw ? Y : X;
    // tan(k * pi / 24), k = 1..5
    int s = (lo >= hi * 0.13165249758739586f ? 1 : 0) + (lo >= hi * 0.2679491924311227f ? 1 : 0) +
            (lo >= hi * 0.41421356237309503f ? 1 : 0) + (lo >= hi * 0.5773502691896257f ? 1 : 0) +
            (lo >= hi * 0.7673269879789602f ? 1 : 0);
    s = sw ? 11 - s : s;
    angle = neg ? 23 - s : s;
  }
  return (angle * (NTHR + 1) + strength) * 3 + coh;
}

// Full key for an N x N window; FAST selects the sqrt/atan-free quantisers.
template <int FAMILY, int N, int G, int NTHR, bool FAST, class WF>
__device__ __forceinline__ int ravu_key2(const mpvp_key_params& kp, WF W) {
  if constexpr (!FAST) {
    return ravu_key<FAMILY, N, G>(kp, W);
  } else {
    constexpr int O = (N - G) / 2;
    // (gx^2, gy^2) and their Gaussian weighting as packed FMUL2: each lane rounds exactly like the scalar multiply
    // of the shader, so the sums are bit-identical to the op-for-op form
    float2 ad = make_float2(0.0f, 0.0f);
    float b = 0.0f;
#pragma unroll
    for (int i = O; i < O + G; ++i) {
#pragma unroll
      for (int j = O; j < O + G; ++j) {
        const float gx = key_diff<FAMILY, N>(i, [&](int dd) { return W(i + dd, j); });
        const float gy = key_diff<FAMILY, N>(j, [&](int dd) { return W(i, j + dd); });
        const float g = kp.gauss[(i - O) * G + (j - O)];
        const float2 gxy = make_float2(gx, gy);
        const float2 t = mul2_rn(mul2_rn(gxy, gxy), make_float2(g, g));
        ad.x = __fadd_rn(ad.x, t.x);  // scalar adds: see the note at mul2_rn
        ad.y = __fadd_rn(ad.y, t.y);
        b = __fadd_rn(b, __fmul_rn(__fmul_rn(gx, gy), g));
      }
    }
    return key_from_abd_fast<NTHR>(kp, ad.x, b, ad.y);
  }
}

// ---- two RAVU keys whose windows are one column apart ---------------------------------------------------------
// The int10 / int01 keys of one pixel (ravu-r2.hook:115-316) read windows on the 45-degree lattice that are shifted
// by ONE step along i: sample (i, j) of the second = sample (i-1, j) of the first.  U(a, j), a in [0, N], j in [0, N),
// is the union window (first key: column a = i + 1, second key: a = i).  One sweep over the columns evaluates every
// j-direction stencil once for both keys, and the i-direction stencil once wherever both keys pick the same
// stencil form (the inner columns); each key's sums still see its own points in the shader's order (i outer, j
// inner) with individually rounded products, so both rows are bit-identical to two separate ravu_key2 calls.
__host__ __device__ constexpr int stencil_kind(int k, int n) {
  return (k - 2 >= 0 && k + 2 <= n - 1) ? 0 : ((k - 1 >= 0 && k + 1 <= n - 1) ? 1 : (k - 1 < 0 ? 2 : 3));
}
template <int N, int G, int NTHR, class UF>
__device__ __forceinline__ void ravu_key_pair(const mpvp_key_params& kp, UF U, int& row_first, int& row_second) {
  constexpr int O = (N - G) / 2;
  float2 ad0 = make_float2(0.0f, 0.0f), ad1 = ad0;
  float b0 = 0.0f, b1 = 0.0f;
#pragma unroll
  for (int a = O; a <= O + G; ++a) {
    const bool use0 = a - 1 >= O && a - 1 < O + G;   // first key: i = a - 1
    const bool use1 = a < O + G;                      // second key: i = a
#pragma unroll
    for (int j = O; j < O + G; ++j) {
      const float gy = key_diff<STENCIL_RAVU, N>(j, [&](int dd) { return U(a, j + dd); });
      float gx0 = 0.0f, gx1 = 0.0f;
      // in the union window the first key's column i sits at a = i + 1, hence the same S for both
      auto S = [&](int dd) { return U(a + dd, j); };
      if (use0) gx0 = key_diff_kind<N>(stencil_kind(a - 1, N), S);
      if (use1) gx1 = (use0 && stencil_kind(a, N) == stencil_kind(a - 1, N)) ? gx0 : key_diff_kind<N>(stencil_kind(a, N), S);
      if (use0) {
        const float g = kp.gauss[(a - 1 - O) * G + (j - O)];
        const float2 gxy = make_float2(gx0, gy);
        const float2 t = mul2_rn(mul2_rn(gxy, gxy), make_float2(g, g));
        ad0.x = __fadd_rn(ad0.x, t.x);
        ad0.y = __fadd_rn(ad0.y, t.y);
        b0 = __fadd_rn(b0, __fmul_rn(__fmul_rn(gx0, gy), g));
      }
      if (use1) {
        const float g = kp.gauss[(a - O) * G + (j - O)];
        const float2 gxy = make_float2(gx1, gy);
        const float2 t = mul2_rn(mul2_rn(gxy, gxy), make_float2(g, g));
        ad1.x = __fadd_rn(ad1.x, t.x);
        ad1.y = __fadd_rn(ad1.y, t.y);
        b1 = __fadd_rn(b1, __fmul_rn(__fmul_rn(gx1, gy), g));
      }
    }
  }
  row_first = key_from_abd_fast<NTHR>(kp, ad0.x, b0, ad0.y);
  row_second = key_from_abd_fast<NTHR>(kp, ad1.x, b1, ad1.y);
}

__device__ __forceinline__ float pow32(float c) {
  c *= c; c *= c; c *= c; c *= c; c *= c;
  return c;
}

// ---- plane formats (mpvp_io) --------------------------------------------------------------------------------
// One load site (tile staging, once per source pixel) and one store site per kernel go through these; the format
// is warp-uniform, so the switch costs a uniform branch.
struct IoFmt {
  int in_fmt = MPVP_FMT_F32, out_fmt = MPVP_FMT_F32;
  float in_max = 1.0f, out_max = 1.0f;
};

inline int parse_io(const mpvp_io* io, IoFmt& f) {
  f = IoFmt{};
  if (!io) return MPVP_OK;
  MPVP_REQUIRE(io->in_format >= MPVP_FMT_F32 && io->in_format <= MPVP_FMT_U16, "bad in_format %d", io->in_format);
  MPVP_REQUIRE(io->out_format >= MPVP_FMT_F32 && io->out_format <= MPVP_FMT_U16, "bad out_format %d", io->out_format);
  f.in_fmt = io->in_format; f.out_fmt = io->out_format;
  if (f.in_fmt >= MPVP_FMT_U8) {
    MPVP_REQUIRE(io->in_max >= 1.0f && io->in_max <= (f.in_fmt == MPVP_FMT_U8 ? 255.0f : 65535.0f), "bad in_max %g", io->in_max);
    f.in_max = io->in_max;
  }
  if (f.out_fmt >= MPVP_FMT_U8) {
    MPVP_REQUIRE(io->out_max >= 1.0f && io->out_max <= (f.out_fmt == MPVP_FMT_U8 ? 255.0f : 65535.0f), "bad out_max %g", io->out_max);
    f.out_max = io->out_max;
  }
  return MPVP_OK;
}
__host__ __device__ __forceinline__ int fmt_bytes(int fmt) { return fmt == MPVP_FMT_F32 ? 4 : (fmt == MPVP_FMT_U8 ? 1 : 2); }

// sample `off` (in elements) of a plane of format fmt, as the shader's HOOKED_tex() would return it
__device__ __forceinline__ float load_px(const void* __restrict__ p, int64_t off, int fmt, float in_max) {
  switch (fmt) {
    case MPVP_FMT_F32: return __ldg(static_cast<const float*>(p) + off);
    case MPVP_FMT_F16: return __half2float(__ldg(static_cast<const __half*>(p) + off));
    case MPVP_FMT_U8: return __fdiv_rn((float)__ldg(static_cast<const unsigned char*>(p) + off), in_max);
    default: return __fdiv_rn((float)__ldg(static_cast<const unsigned short*>(p) + off), in_max);
  }
}
// The staging loops hoist the format switch OUT of the loop (a switch inside it keeps the compiler from batching the
// loads of several iterations: ravu-r3-rgb 3.24 -> 3.84 ms): dispatch_in_fmt runs the loop body instantiated for the
// plane format, load_px_t is the monomorphic load.
template <int FMT>
__device__ __forceinline__ float load_px_t(const void* __restrict__ p, int64_t off, float in_max) {
  if constexpr (FMT == MPVP_FMT_F32) return __ldg(static_cast<const float*>(p) + off);
  else if constexpr (FMT == MPVP_FMT_F16) return __half2float(__ldg(static_cast<const __half*>(p) + off));
  else if constexpr (FMT == MPVP_FMT_U8) return __fdiv_rn((float)__ldg(static_cast<const unsigned char*>(p) + off), in_max);
  else return __fdiv_rn((float)__ldg(static_cast<const unsigned short*>(p) + off), in_max);
}
template <class F>
__device__ __forceinline__ void dispatch_in_fmt(int fmt, F&& body) {
  switch (fmt) {
    case MPVP_FMT_F32: body(std::integral_constant<int, MPVP_FMT_F32>{}); break;
    case MPVP_FMT_F16: body(std::integral_constant<int, MPVP_FMT_F16>{}); break;
    case MPVP_FMT_U8: body(std::integral_constant<int, MPVP_FMT_U8>{}); break;
    default: body(std::integral_constant<int, MPVP_FMT_U16>{}); break;
  }
}
__device__ __forceinline__ unsigned int quant_px(float v, float out_max) {
  return __float2uint_rn(fminf(fmaxf(v, 0.0f), 1.0f) * out_max);
}
// one pixel at element offset off (streaming store)
__device__ __forceinline__ void store_px(void* __restrict__ p, int64_t off, float v, int fmt, float out_max) {
  switch (fmt) {
    case MPVP_FMT_F32: __stcs(static_cast<float*>(p) + off, v); break;
    case MPVP_FMT_F16: __stcs(reinterpret_cast<unsigned short*>(p) + off, __half_as_ushort(__float2half_rn(v))); break;
    case MPVP_FMT_U8: __stcs(static_cast<unsigned char*>(p) + off, (unsigned char)quant_px(v, out_max)); break;
    default: __stcs(static_cast<unsigned short*>(p) + off, (unsigned short)quant_px(v, out_max)); break;
  }
}
// two horizontally adjacent pixels at an EVEN element offset (one vector store)
__device__ __forceinline__ void store_px2(void* __restrict__ p, int64_t off, float a, float b, int fmt, float out_max) {
  switch (fmt) {
    case MPVP_FMT_F32: __stcs(reinterpret_cast<float2*>(static_cast<float*>(p) + off), make_float2(a, b)); break;
    case MPVP_FMT_F16: {
      const __half2 h = __floats2half2_rn(a, b);
      __stcs(reinterpret_cast<unsigned int*>(static_cast<__half*>(p) + off), *reinterpret_cast<const unsigned int*>(&h));
      break;
    }
    case MPVP_FMT_U8:
      __stcs(reinterpret_cast<unsigned short*>(static_cast<unsigned char*>(p) + off),
             (unsigned short)(quant_px(a, out_max) | (quant_px(b, out_max) << 8)));
      break;
    default:
      __stcs(reinterpret_cast<unsigned int*>(static_cast<unsigned short*>(p) + off), quant_px(a, out_max) | (quant_px(b, out_max) << 16));
      break;
  }
}

// ---- persistent-CTA tile walk ----------------------------------------------------------------------------------
// tile = (f * tiles_y + tiy) * tiles_x + tix, visited as first, first + step, first + 2 step, ...  Decoding a 64-bit tile
// index costs three 64-bit divisions (~200 instructions) per tile and thread -- 5-10 % of a small-network NNEDI3 tile;
// the walk decodes `first` and `step` once (32-bit) and then advances with carries only.
struct TileWalk {
  int tix, tiy, f;       // current tile
  int sx, sy, sf;        // the step in the same mixed radix
  int tiles_x, tiles_y;
  __device__ __forceinline__ TileWalk(long long first, long long step, int tx, int ty) : tiles_x(tx), tiles_y(ty) {
    const unsigned a = (unsigned)first, b = (unsigned)step;   // the host guarantees total_tiles < 2^31
    unsigned q = a / (unsigned)tx;
    tix = (int)(a - q * (unsigned)tx);
    f = (int)(q / (unsigned)ty);
    tiy = (int)(q - (unsigned)f * (unsigned)ty);
    q = b / (unsigned)tx;
    sx = (int)(b - q * (unsigned)tx);
    sf = (int)(q / (unsigned)ty);
    sy = (int)(q - (unsigned)sf * (unsigned)ty);
  }
  __device__ __forceinline__ void next() {
    tix += sx;
    int c = tix >= tiles_x ? 1 : 0;
    tix -= c ? tiles_x : 0;
    tiy += sy + c;
    c = tiy >= tiles_y ? 1 : 0;
    tiy -= c ? tiles_y : 0;
    f += sf + c;
  }
};

__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace mpvp
