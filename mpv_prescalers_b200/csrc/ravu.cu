// RAVU (non-lite): the four reference passes in ONE kernel.
//
//   step1  int11 = value at (x+1/2, y+1/2)         ravu-r2.hook:15-114   (rgb: ravu-r2-rgb.hook:15-131)
//   step2  int10 = value at (x+1/2, y)             ravu-r2.hook:115-215  (45-degree rotated lattice)
//   step3  int01 = value at (x, y+1/2)             ravu-r2.hook:216-316
//   step4  merge into 2W x 2H, OFFSET -0.5 -0.5    ravu-r2.hook:317-338
//
// A CTA owns a 64x32 tile of input pixels.  Phase A computes int11 on the tile plus the halo that
// steps 2/3 will tap ((64+2r-1) x (32+2r-1) positions) into shared memory; positions outside the
// image replicate the border int11 value (clamp-to-edge on the *saved texture*),
// which is done by evaluating int11 at the clamped coordinate.  Phase B computes int10 and int01
// from the staged HOOKED tile and the int11 tile and writes the interleaved 2x2 block.  int11 never
// touches HBM.  -yuv / -rgb variants carry three colour planes plus the key plane.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "tma.cuh"

namespace mpvp {
namespace {

struct RavuArgs {
  const void* __restrict__ in;   // planes of format io.in_fmt
  void* __restrict__ out;        // planes of format io.out_fmt
  IoFmt io;
  const float4* __restrict__ lut;  // [648][LW]
  const uint2* __restrict__ lut_half;  // same texels as 4 x binary16 (null if the LUT was not rounded to fp16)
  int32_t* __restrict__ bucket;    // [n][3][h][w] or null
  int n, h, w;
  int64_t in_sn, in_sc, in_sy, out_sn, out_sc, out_sy;
  int tiles_x, tiles_y;
  long long total_tiles;
  mpvp_key_params key;
};

#ifndef MPVP_X_RAVU_TH1
#define MPVP_X_RAVU_TH1 64
#endif
#ifndef MPVP_X_R23_NT1
#define MPVP_X_R23_NT1 512
#endif
#ifndef MPVP_X_R4_NT1
#define MPVP_X_R4_NT1 256
#endif
#ifndef MPVP_X_R4_NT3
#define MPVP_X_R4_NT3 384
#endif
#ifndef MPVP_X_RAVU_TH3
#define MPVP_X_RAVU_TH3 48
#endif
constexpr int kTW = 64;
// tile height: the int11 halo (2r-1 rows and columns are recomputed per tile) costs (64+7)(TH+7)/(64 TH) of phase A:
// 1.35x at TH = 32, 1.23x at TH = 64 (ravu-r3 1.91 -> 1.66 ms, TH = 128: 1.63); three-channel tiles (4 planes of
// HOOKED + int11) fit shared memory up to TH = 56
template <int C>
struct TileH {
  static constexpr int v = C == 1 ? MPVP_X_RAVU_TH1 : MPVP_X_RAVU_TH3;
};
constexpr float kCp0 = 0.2126f, kCp1 = 0.7152f, kCp2 = 0.0722f;

__device__ __forceinline__ float rgb_luma(float r, float g, float b) {
  // dot(rgb, color_primary) evaluated left to right without contraction (ravu-r2-rgb.hook:24)
  return __fadd_rn(__fadd_rn(__fmul_rn(r, kCp0), __fmul_rn(g, kCp1)), __fmul_rn(b, kCp2));
}

template <int R, int C, bool LH, class KF, class CF>
__device__ __forceinline__ void ravu_apply(int row, const void* __restrict__ s_lut_raw, KF KS, CF CS, float (&res)[C]);

// One key + convolution.  KS(t) = key sample t, CS(c, t) = colour sample of channel c.
// LH: the LUT sits in shared memory as 4 x binary16 per texel (exact, the texels are binary16 values): the
// per-lane row gather is bank-conflict bound and 8-byte texels need half the wavefronts of 16-byte ones.
template <int R, int C, bool LH, class KF, class CF>
__device__ __forceinline__ int ravu_conv(const mpvp_key_params& kp, const void* __restrict__ s_lut_raw, KF KS, CF CS,
                                         float (&res)[C]) {
  constexpr int N = 2 * R, TAPS = N * N, G = (R == 4) ? 6 : 4;
  constexpr int LW = (TAPS / 2 + 3) / 4;
  constexpr int LWP = LW | 1;  // odd float4 pitch in shared memory: rows spread over all banks (r4: 8 -> 9)
  float ks[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) ks[t] = KS(t);
  const int row = ravu_key2<STENCIL_RAVU, N, G, 8, true>(kp, [&](int i, int j) { return ks[i * N + j]; });
  ravu_apply<R, C, LH>(row, s_lut_raw, [&](int t) { return ks[t]; }, CS, res);
  return row;
}

// The convolution of one pass with the weights of LUT row `row`: res = sum_k (s_k + s_{N-1-k}) * w_k, clamp(0, 1)
template <int R, int C, bool LH, class KF, class CF>
__device__ __forceinline__ void ravu_apply(int row, const void* __restrict__ s_lut_raw, KF KS, CF CS, float (&res)[C]) {
  constexpr int N = 2 * R, TAPS = N * N;
  constexpr int LW = (TAPS / 2 + 3) / 4;
  constexpr int LWP = LW | 1;
#pragma unroll
  for (int c = 0; c < C; ++c) res[c] = 0.f;
#pragma unroll
  for (int q = 0; q < LW; ++q) {
    float wv[4];
    if constexpr (LH) {
      const uint2 u = reinterpret_cast<const uint2*>(s_lut_raw)[row * LWP + q];
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      wv[0] = a.x; wv[1] = a.y; wv[2] = b.x; wv[3] = b.y;
    } else {
      const float4 w4 = reinterpret_cast<const float4*>(s_lut_raw)[row * LWP + q];
      wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = q * 4 + e;
      if (k < TAPS / 2) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float sa = (C == 1) ? KS(k) : CS(c, k);
          const float sb = (C == 1) ? KS(TAPS - 1 - k) : CS(c, TAPS - 1 - k);
          res[c] = fmaf(sa + sb, wv[e], res[c]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) res[c] = fminf(fmaxf(res[c], 0.f), 1.f);
}

// KEYMODE: 0 luma (C=1), 1 yuv (key = channel 0), 2 rgb (key = BT.709 luma)
// OF32: float32 output planes at compile time (the common case; false = any mpvp_io output format)
// TMA: the HOOKED tile + halo arrives by cp.async.bulk.tensor (float32 planes that meet the 16-byte rules, tma.cuh): the
// box starts XO = 4 / 8 texels left of the tile (16-byte aligned origin) and is 72 / 80 texels wide; luma kernels keep two
// tile buffers so that the next tile is in flight while the current one is computed (three-channel tiles fill shared
// memory: one buffer, issue and wait); border tiles are patched to clamp-to-edge after arrival.
// GP (r3 / r4 luma): step 1 takes its gradients from per-texel GRADIENT PLANES in shared memory.  A gradient stencil
// (4th-order where +-2 fits in the window, central at the window's rim; ravu-r4.hook:86-252) depends on the texel and on
// which of the two forms the window position selects, not on the window itself, so both forms are computed ONCE per
// texel -- (gx4, gy4) and (gxc, gyc) -- instead of once per window that contains the texel (36 times for r4).  Each
// value is the same expression on the same samples, so the key sums are bit-identical to the per-window evaluation.
#ifndef MPVP_X_RAVU_GP
#define MPVP_X_RAVU_GP 0   // measured: bit-identical, but not faster (ravu-r4 x16: 3.30 vs 2.87 ms at 256 threads, 2.90 at 512; ravu-r3 1.66 vs 1.55): the extra phase + barrier of a 1-CTA/SM kernel costs what the stencils save
#endif
template <int R, int C>
struct UseGP {
  static constexpr bool v = MPVP_X_RAVU_GP && C == 1 && (R == 3 || R == 4);
};
template <int R, int C, int KEYMODE, int NT, bool LH, bool OF32, bool TMA>
__global__ void __launch_bounds__(NT, 1) ravu_kernel(const __grid_constant__ RavuArgs A, const __grid_constant__ CUtensorMap tmap) {
  constexpr int N = 2 * R, TAPS = N * N;
  constexpr bool GP = UseGP<R, C>::v;
  constexpr int GG = (R == 4) ? 6 : 4;                       // gradient square side
  constexpr int GO = (N - GG) / 2;                           // first gradient index in the window (1 for r3 / r4)
  constexpr int GW = kTW + GG + 2 * R - 2, GH = TileH<C>::v + GG + 2 * R - 2;   // gradient-plane extent (texels)
  constexpr int kTH = TileH<C>::v;
  constexpr int LW = (TAPS / 2 + 3) / 4;
  constexpr int LWP = LW | 1;
  constexpr int HH = 2 * R - 1;              // HOOKED halo
  constexpr int XO = TMA ? ((HH + 3) & ~3) : HH;          // staged columns left of the tile
  constexpr int HW_ = TMA ? ((XO + kTW + HH + 3) & ~3) : (kTW + 2 * HH), HHt = kTH + 2 * HH;   // staged HOOKED tile
#ifndef MPVP_X_RAVU_NBUF_R4
#define MPVP_X_RAVU_NBUF_R4 1   // r4 luma: one buffer (the second buffer's per-iteration base costs more than the overlap wins: 3.29 vs 2.87 ms per 16 frames)
#endif
  constexpr int NBUF = (TMA && C == 1) ? (R == 4 ? MPVP_X_RAVU_NBUF_R4 : 2) : 1;
  constexpr int HPL = TMA ? ((HHt * HW_ + 31) & ~31) : HHt * HW_;   // plane pitch (TMA destinations are 128-byte aligned)
  constexpr int IW = kTW + 2 * R - 1, IH = kTH + 2 * R - 1;  // int11 tile: x' in [x0-R, x0+TW+R-2]
  constexpr int NP = (C == 1) ? 1 : ((KEYMODE == 2) ? 4 : 3);  // planes: colours (+ key plane for rgb)
  constexpr int KP = (C == 1) ? 0 : ((KEYMODE == 2) ? 3 : 0);  // index of the key plane

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_lut = reinterpret_cast<float4*>(smem_raw);
  constexpr int kLutBytes = (int)(LH ? sizeof(uint2) : sizeof(float4)) * 648 * LWP;
  constexpr int HBUF = (NP * HPL + 31) & ~31;                              // floats per HOOKED buffer (128-byte multiple)
  float* s_h0 = reinterpret_cast<float*>(smem_raw + ((kLutBytes + 127) & ~127));  // [NBUF][NP][HHt][HW_]
  float* s_i = s_h0 + NBUF * HBUF;                                               // [NP][IH][IW]
  float2* s_g4 = reinterpret_cast<float2*>(s_i + ((NP * IH * IW + 3) & ~3));     // [GH][GW] (gx, gy) 4th-order   (GP)
  float2* s_gc = s_g4 + GW * GH;                                                 // [GH][GW] (gx, gy) central     (GP)
  __shared__ __align__(8) uint64_t s_mbar[2];

  const int tid = threadIdx.x;
  if constexpr (TMA) {
    if (tid == 0) {
      mbar_init1(smem_addr(&s_mbar[0]));
      mbar_init1(smem_addr(&s_mbar[1]));
      mbar_init_fence();
    }
  }
  const uint64_t tmap_ptr = reinterpret_cast<uint64_t>(&tmap);
  // one thread asks the TMA engine for the colour planes of a tile (plane index = frame * C + channel)
  auto tma_issue = [&, tmap_ptr](const TileWalk& tw, int buf) {
    const uint32_t bar = smem_addr(&s_mbar[buf]);
    tma_expect(bar, C * HHt * HW_ * 4);
#pragma unroll
    for (int c = 0; c < C; ++c)
      tma_load_3d(smem_addr(s_h0 + buf * HBUF + c * HPL), tmap_ptr, tw.tix * kTW - XO, tw.tiy * kTH - HH, tw.f * C + c, bar);
  };
  if constexpr (LH) {
    for (int i = tid; i < 648 * LW; i += NT) reinterpret_cast<uint2*>(smem_raw)[(i / LW) * LWP + (i % LW)] = A.lut_half[i];
  } else {
    for (int i = tid; i < 648 * LW; i += NT) s_lut[(i / LW) * LWP + (i % LW)] = A.lut[i];
  }

  TileWalk walk(blockIdx.x, gridDim.x, A.tiles_x, A.tiles_y);
  TileWalk ahead = walk;
  __syncthreads();   // LUT and mbarriers are set up
  if constexpr (TMA && NBUF == 2) {
    if (tid < 32 && blockIdx.x < A.total_tiles) {
      if (elect_one()) tma_issue(ahead, 0);
    }
  }
  ahead.next();
  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, walk.next(), ahead.next(), ++it) {
    const int tix = walk.tix, tiy = walk.tiy, f = walk.f;
    const int x0 = tix * kTW, y0 = tiy * kTH;
    const int64_t src0 = (int64_t)f * A.in_sn;
    float* __restrict__ s_h = s_h0 + (NBUF == 2 ? (it & 1) * HBUF : 0);

    __syncthreads();   // previous tile fully consumed (its HOOKED buffer and the int11 tile are free)
    if constexpr (TMA) {
      if constexpr (NBUF == 2) {
        if (tid < 32 && tile + gridDim.x < A.total_tiles) {
          fence_proxy_async_smem();
          if (elect_one()) tma_issue(ahead, (it + 1) & 1);
        }
        mbar_wait_parity(smem_addr(&s_mbar[it & 1]), (it >> 1) & 1);
      } else {
        if (tid < 32) {
          fence_proxy_async_smem();
          if (elect_one()) tma_issue(walk, 0);
        }
        mbar_wait_parity(smem_addr(&s_mbar[0]), it & 1);
      }
      const bool edge = x0 - XO < 0 || y0 - HH < 0 || x0 - XO + HW_ > A.w || y0 - HH + HHt > A.h;
      if (edge) {   // CTA-uniform
#pragma unroll
        for (int c = 0; c < C; ++c)
          patch_clamp_to_edge(s_h + c * HPL, HW_, HW_, HHt, x0 - XO, y0 - HH, A.w, A.h, tid, NT, [] { __syncthreads(); });
      }
      if constexpr (KEYMODE == 2) {
        for (int i = tid; i < HW_ * HHt; i += NT)
          s_h[3 * HPL + i] = rgb_luma(s_h[i], s_h[HPL + i], s_h[2 * HPL + i]);
      }
    } else {
    // ---- stage HOOKED (clamp-to-edge) ---------------------------------------------------
    dispatch_in_fmt(A.io.in_fmt, [&](auto ftag) {
      constexpr int FMT = decltype(ftag)::value;
      for (int i = tid; i < HW_ * HHt; i += NT) {
        const int sy = i / HW_, sx = i - sy * HW_;
        const int gx = clampi(x0 + sx - XO, 0, A.w - 1);
        const int gy = clampi(y0 + sy - HH, 0, A.h - 1);
        const int64_t off = src0 + (int64_t)gy * A.in_sy + gx;
        if (C == 1) {
          s_h[i] = load_px_t<FMT>(A.in, off, A.io.in_max);
        } else {
          const float c0 = load_px_t<FMT>(A.in, off, A.io.in_max);
          const float c1 = load_px_t<FMT>(A.in, off + A.in_sc, A.io.in_max);
          const float c2 = load_px_t<FMT>(A.in, off + 2 * A.in_sc, A.io.in_max);
          s_h[i] = c0;
          s_h[HPL + i] = c1;
          s_h[2 * HPL + i] = c2;
          if (KEYMODE == 2) s_h[3 * HPL + i] = rgb_luma(c0, c1, c2);
        }
      }
    });
    }
    __syncthreads();

    // ---- gradient planes of the HOOKED tile (GP) ---------------------------------------------
    // plane texel (0, 0) is image texel (x0 - (2R-2), y0 - (2R-2)): the first gradient point of the left / top-most window
    if constexpr (GP) {
      for (int i = tid; i < GW * GH; i += NT) {
        const int gy_ = i / GW, gx_ = i - gy_ * GW;
        // staged coordinates of this texel
        const int sx = gx_ - (2 * R - 2) + XO, sy = gy_ - (2 * R - 2) + HH;
        const float* __restrict__ p = s_h + sy * HW_ + sx;
        float2 g4 = make_float2(0.f, 0.f), gc = g4;
        // a form whose stencil leaves the staged tile is never selected by any window (it would leave that window too)
        if (sx >= 2 && sx + 2 < HW_) {
          float t = __fmaf_rn(8.0f, p[1], -p[2]);
          t = __fmaf_rn(-8.0f, p[-1], t);
          g4.x = div12_rn(__fadd_rn(t, p[-2]));
        }
        if (sy >= 2 && sy + 2 < HHt) {
          float t = __fmaf_rn(8.0f, p[HW_], -p[2 * HW_]);
          t = __fmaf_rn(-8.0f, p[-HW_], t);
          g4.y = div12_rn(__fadd_rn(t, p[-2 * HW_]));
        }
        if (sx >= 1 && sx + 1 < HW_) gc.x = __fmul_rn(__fsub_rn(p[1], p[-1]), 0.5f);
        if (sy >= 1 && sy + 1 < HHt) gc.y = __fmul_rn(__fsub_rn(p[HW_], p[-HW_]), 0.5f);
        s_g4[i] = g4;
        s_gc[i] = gc;
      }
      __syncthreads();
    }

    // ---- phase A: int11 on the tile + halo ------------------------------------------------
    for (int i = tid; i < IW * IH; i += NT) {
      const int iy = i / IW, ix = i - iy * IW;
      const int px = x0 - R + ix, py = y0 - R + iy;          // int11 texel this slot stands for
      const int cx = clampi(px, 0, A.w - 1), cy = clampi(py, 0, A.h - 1);
      // staged coordinates of the window origin (tap offset -(R-1))
      const int bx = cx - (R - 1) - (x0 - XO), by = cy - (R - 1) - (y0 - HH);
      const float* __restrict__ kb = s_h + KP * HPL + by * HW_ + bx;
      const float* __restrict__ cb = s_h + by * HW_ + bx;
      float res[C];
      int row;
      if constexpr (GP) {
        // window point (i, j) is plane texel (cx - (R-1) + i - (x0 - (2R-2)), ...): same sums, same order as ravu_key2
        const int qx = cx - x0 + (R - 1), qy = cy - y0 + (R - 1);
        float2 ad = make_float2(0.0f, 0.0f);
        float b = 0.0f;
#pragma unroll
        for (int ii = GO; ii < GO + GG; ++ii) {
#pragma unroll
          for (int jj = GO; jj < GO + GG; ++jj) {
            const int q = (qy + jj) * GW + qx + ii;
            const bool x4 = ii - 2 >= 0 && ii + 2 <= N - 1, y4 = jj - 2 >= 0 && jj + 2 <= N - 1;
            float gx, gy;
            if (x4 && y4) { const float2 v = s_g4[q]; gx = v.x; gy = v.y; }
            else if (!x4 && !y4) { const float2 v = s_gc[q]; gx = v.x; gy = v.y; }
            else { gx = x4 ? s_g4[q].x : s_gc[q].x; gy = y4 ? s_g4[q].y : s_gc[q].y; }
            const float g = A.key.gauss[(ii - GO) * GG + (jj - GO)];
            const float2 gxy = make_float2(gx, gy);
            const float2 t = mul2_rn(mul2_rn(gxy, gxy), make_float2(g, g));
            ad.x = __fadd_rn(ad.x, t.x);
            ad.y = __fadd_rn(ad.y, t.y);
            b = __fadd_rn(b, __fmul_rn(__fmul_rn(gx, gy), g));
          }
        }
        row = key_from_abd_fast<8>(A.key, ad.x, b, ad.y);
        ravu_apply<R, C, LH>(row, s_lut, [&](int t) { return kb[(t % N) * HW_ + (t / N)]; },
                             [&](int c, int t) { return cb[c * HPL + (t % N) * HW_ + (t / N)]; }, res);
      } else {
        row = ravu_conv<R, C, LH>(
            A.key, s_lut, [&](int t) { return kb[(t % N) * HW_ + (t / N)]; },
            [&](int c, int t) { return cb[c * HPL + (t % N) * HW_ + (t / N)]; }, res);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) s_i[c * IH * IW + i] = res[c];
      if (KEYMODE == 2) s_i[3 * IH * IW + i] = rgb_luma(res[0], res[C > 1 ? 1 : 0], res[C > 2 ? 2 : 0]);
      if (A.bucket && px == cx && py == cy && px >= x0 && px < x0 + kTW && py >= y0 && py < y0 + kTH)
        A.bucket[(((int64_t)f * 3 + 0) * A.h + py) * A.w + px] = row;
    }
    __syncthreads();

    // ---- phase B: int10 / int01 + merge ---------------------------------------------------
    for (int i = tid; i < kTW * kTH; i += NT) {
      const int ly = i / kTW, lx = i - ly * kTW;
      const int x = x0 + lx, y = y0 + ly;
      if (x >= A.w || y >= A.h) continue;
      const float* __restrict__ hb = s_h + (ly + HH) * HW_ + (lx + XO);  // HOOKED(x, y)
      const float* __restrict__ ib = s_i + (ly + R) * IW + (lx + R);     // int11(x, y)
      float r10[C], r01[C];
      int rows[2];
      // sample (i, j) of a pass sits at twice-the-real-position (tx2 - (2R-1) + i + j, ty2 - i + j) relative to (2x, 2y),
      // with (tx2, ty2) = (1, 0) for int10 and (0, 1) for int01: even positions are HOOKED texels, odd ones int11.
      // The int01 window is the int10 window shifted by one along i (sample (i, j) of int01 = sample (i-1, j) of
      // int10), so both keys read ONE (N+1) x N register window U[a][j] = int10 sample (a-1, j): the samples are
      // loaded once, and every gradient / product the two keys have in common (all the j-direction stencils of the
      // shared columns, the 4th-order i-direction ones of the inner columns) is the SAME expression on the same
      // registers, which the compiler evaluates once.
      auto fetch0 = [&](int plane, int ii, int jj) -> float {
        const int px2 = 1 - (2 * R - 1) + ii + jj, py2 = -ii + jj;
        if ((px2 & 1) == 0) return hb[plane * HPL + (py2 / 2) * HW_ + (px2 / 2)];
        return ib[plane * IH * IW + ((py2 - 1) / 2) * IW + ((px2 - 1) / 2)];
      };
      float U[(N + 1) * N];
#pragma unroll
      for (int a = 0; a <= N; ++a)
#pragma unroll
        for (int jj = 0; jj < N; ++jj) U[a * N + jj] = fetch0(KP, a - 1, jj);
      if constexpr (R == 4 && C == 1) {
        // both keys in one sweep over the union window: ravu-r4 3.41 -> 2.88 ms per 16 frames.  (r2 / r3 share too few
        // stencils to pay for the second accumulator set, +1..3 %; the 3-channel kernels are bound by their colour
        // sample loads and do not gain.)
        ravu_key_pair<N, 6, 8>(A.key, [&](int a, int jj) { return U[a * N + jj]; }, rows[0], rows[1]);
        ravu_apply<R, C, LH>(rows[0], s_lut, [&](int t) { return U[(t / N + 1) * N + t % N]; },
                             [&](int c, int t) { return fetch0(c, t / N, t % N); }, r10);
        ravu_apply<R, C, LH>(rows[1], s_lut, [&](int t) { return U[(t / N) * N + t % N]; },
                             [&](int c, int t) { return fetch0(c, t / N - 1, t % N); }, r01);
      } else {
        rows[0] = ravu_conv<R, C, LH>(
            A.key, s_lut, [&](int t) { return U[(t / N + 1) * N + t % N]; },
            [&](int c, int t) { return fetch0(c, t / N, t % N); }, r10);
        rows[1] = ravu_conv<R, C, LH>(
            A.key, s_lut, [&](int t) { return U[(t / N) * N + t % N]; },
            [&](int c, int t) { return fetch0(c, t / N - 1, t % N); }, r01);
      }
      if (A.bucket) {
        A.bucket[(((int64_t)f * 3 + 1) * A.h + y) * A.w + x] = rows[0];
        A.bucket[(((int64_t)f * 3 + 2) * A.h + y) * A.w + x] = rows[1];
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int64_t o = (int64_t)f * A.out_sn + c * A.out_sc + (int64_t)(2 * y) * A.out_sy + 2 * x;
        // (2x,2y)=HOOKED (2x+1,2y)=int10 (2x,2y+1)=int01 (2x+1,2y+1)=int11   (ravu-r2.hook:327-338)
        const int ofmt = OF32 ? MPVP_FMT_F32 : A.io.out_fmt;
        store_px2(A.out, o, hb[c * HPL], r10[c], ofmt, A.io.out_max);
        store_px2(A.out, o + A.out_sy, r01[c], ib[c * IH * IW], ofmt, A.io.out_max);
      }
    }
  }
}

template <int R, int C, int KEYMODE, int NT, bool LH, bool OF32>
int launch_ravu_impl(const RavuArgs& a0, int device, cudaStream_t stream) {
  constexpr int N = 2 * R, TAPS = N * N, LW = ((TAPS / 2 + 3) / 4) | 1, HH = 2 * R - 1;  // LW: padded pitch
  constexpr int kTH = TileH<C>::v;
  constexpr int NP = (C == 1) ? 1 : ((KEYMODE == 2) ? 4 : 3);
  constexpr int XO_T = (HH + 3) & ~3, HW_T = (XO_T + kTW + HH + 3) & ~3, HHt = kTH + 2 * HH;
  RavuArgs a = a0;
  a.tiles_x = (a.w + kTW - 1) / kTW;
  a.tiles_y = (a.h + kTH - 1) / kTH;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * a.n;
  MPVP_REQUIRE(a.total_tiles < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", a.total_tiles);
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  // float32 planes whose channels are evenly spaced (plane index = frame * C + channel) can be fetched by TMA
  bool use_tma = false;
  if (a.io.in_fmt == MPVP_FMT_F32 && (C == 1 || a.in_sn == (int64_t)C * a.in_sc))
    use_tma = make_plane_tmap(&tmap, a.in, 4, a.w, a.h, a.n * C, a.in_sy, C == 1 ? a.in_sn : a.in_sc, HW_T, HHt);
  const int hw = use_tma ? HW_T : (kTW + 2 * HH), nbuf = (use_tma && C == 1) ? (R == 4 ? MPVP_X_RAVU_NBUF_R4 : 2) : 1;
  const size_t hpl = use_tma ? (((size_t)HHt * hw + 31) & ~(size_t)31) : (size_t)HHt * hw;
  const size_t hbuf = (NP * hpl + 31) & ~(size_t)31;
  constexpr int GG = (R == 4) ? 6 : 4;
  constexpr size_t kGradBytes = UseGP<R, C>::v ? 2 * sizeof(float2) * (size_t)(kTW + GG + 2 * R - 2) * (kTH + GG + 2 * R - 2) : 0;
  const size_t smem = ((((LH ? sizeof(uint2) : sizeof(float4)) * 648 * LW) + 127) & ~(size_t)127) +
                      sizeof(float) * (nbuf * hbuf + (((size_t)NP * (kTW + 2 * R - 1) * (kTH + 2 * R - 1) + 3) & ~(size_t)3)) + kGradBytes;
  auto kern = use_tma ? ravu_kernel<R, C, KEYMODE, NT, LH, OF32, true> : ravu_kernel<R, C, KEYMODE, NT, LH, OF32, false>;
  MPVP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MPVP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
  if (per_sm < 1) {
    set_error("ravu kernel does not fit on an SM (smem %zu B)", smem);
    return MPVP_E_UNSUPPORTED;
  }
  long long grid = (long long)sm_count(device) * per_sm;
  if (grid > a.total_tiles) grid = a.total_tiles;
  grid = cap_grid(grid);
  if (grid < 1) return MPVP_OK;
  kern<<<(unsigned)grid, NT, smem, stream>>>(a, tmap);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}

// MPVP_LUT_SMEM=fp32 keeps 16-byte texels in shared memory (A/B switch)
template <int R, int C, int KEYMODE, int NT>
int launch_ravu(const RavuArgs& a, int device, cudaStream_t stream) {
  static const bool half_ok = [] {
    const char* e = getenv("MPVP_LUT_SMEM");
    return !(e && e[0] == 'f' && e[2] == '3');
  }();
  if (a.lut_half && half_ok) {
    if (a.io.out_fmt == MPVP_FMT_F32) return launch_ravu_impl<R, C, KEYMODE, NT, true, true>(a, device, stream);
    return launch_ravu_impl<R, C, KEYMODE, NT, true, false>(a, device, stream);
  }
  return launch_ravu_impl<R, C, KEYMODE, NT, false, false>(a, device, stream);
}

}  // namespace
}  // namespace mpvp

using namespace mpvp;

extern "C" int mpvp_ravu_launch(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                                const float* in, float* out, int n, int h, int w, int64_t in_stride_n,
                                int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_c,
                                int64_t out_stride_y, int32_t* bucket_out, void* stream) {
  return mpvp_ravu_launch_io(lut, key, radius, key_mode, in, out, n, h, w, in_stride_n, in_stride_c, in_stride_y,
                             out_stride_n, out_stride_c, out_stride_y, bucket_out, nullptr, stream);
}

extern "C" int mpvp_ravu_launch_io(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                                   const void* in, void* out, int n, int h, int w, int64_t in_stride_n,
                                   int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n,
                                   int64_t out_stride_c, int64_t out_stride_y, int32_t* bucket_out, const mpvp_io* io,
                                   void* stream) {
  IoFmt iof;
  if (int rc0 = parse_io(io, iof)) return rc0;
  MPVP_REQUIRE(lut && lut->kind == 0 && lut->lut, "lut handle is null or not a LUT");
  MPVP_REQUIRE(key && in && out, "null argument");
  MPVP_REQUIRE(radius >= 2 && radius <= 4, "radius %d not in {2,3,4}", radius);
  MPVP_REQUIRE(key_mode >= 0 && key_mode <= 2, "key_mode %d", key_mode);
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad frame geometry n=%d h=%d w=%d", n, h, w);
  const int taps = 4 * radius * radius, g = radius == 4 ? 6 : 4;
  MPVP_REQUIRE(lut->lut_w == (taps / 2 + 3) / 4 && lut->lut_h == 648, "LUT is %dx%d, expected %dx648", lut->lut_w,
               lut->lut_h, (taps / 2 + 3) / 4);
  MPVP_REQUIRE(key->n_gauss == g * g && key->n_strength == 9 && key->n_strength_thr == 0,
               "key params do not describe a RAVU (log2-strength) hook");
  if (int rck = check_fast_key(key)) return rck;
  MPVP_REQUIRE((out_stride_y % 2) == 0 && (out_stride_n % 2) == 0 && (out_stride_c % 2) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) % (2 * fmt_bytes(iof.out_fmt))) == 0,
               "output rows must be aligned to a pixel pair (even strides, base aligned to two elements)");
  if (n == 0) return MPVP_OK;
  DeviceGuard guard(lut->device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", lut->device);
  RavuArgs a{};
  a.io = iof;
  a.in = in; a.out = out; a.lut = reinterpret_cast<const float4*>(lut->lut); a.lut_half = reinterpret_cast<const uint2*>(lut->lut_half); a.bucket = bucket_out;
  a.n = n; a.h = h; a.w = w;
  a.in_sn = in_stride_n; a.in_sc = in_stride_c; a.in_sy = in_stride_y;
  a.out_sn = out_stride_n; a.out_sc = out_stride_c; a.out_sy = out_stride_y;
  a.key = *key;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dev = lut->device;
  switch (radius * 3 + key_mode) {
    case 6: return launch_ravu<2, 1, 0, MPVP_X_R23_NT1>(a, dev, st);
    case 7: return launch_ravu<2, 3, 1, 512>(a, dev, st);
    case 8: return launch_ravu<2, 3, 2, 512>(a, dev, st);
    case 9: return launch_ravu<3, 1, 0, MPVP_X_R23_NT1>(a, dev, st);
    case 10: return launch_ravu<3, 3, 1, 512>(a, dev, st);
    case 11: return launch_ravu<3, 3, 2, 512>(a, dev, st);
    case 12: return launch_ravu<4, 1, 0, MPVP_X_R4_NT1>(a, dev, st);
    case 13: return launch_ravu<4, 3, 1, MPVP_X_R4_NT3>(a, dev, st);
    case 14: return launch_ravu<4, 3, 2, MPVP_X_R4_NT3>(a, dev, st);
  }
  return MPVP_E_INVALID;
}
