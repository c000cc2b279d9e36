#pragma once
// RAVU-Lite(-AR) and RAVU-3x: one fused kernel per call (kernel + launcher templates; instantiated by ravu_lite.cu
// for the 2x family and by ravu_3x.cu for RAVU-3x, two translation units so that they compile in parallel).
//
// Reference passes replaced (all of them in ONE launch, nothing round-trips HBM):
//   RAVU-Lite(-AR) step1 + step2        ravu-lite-ar-r3.hook:15-197   (compute form :22-170)
//   RAVU-3x                             compute/ravu-3x-r2.hook:15-115
//
// Design (B200): persistent CTAs (grid = SMs x resident CTAs) walk a (frame, tile) work list.  The
// whole LUT lives in shared memory for the lifetime of the CTA (r3: 288 x 13 float4 = 58.5 KB;
// values are exactly the fp16-rounded texels the reference's rgba16f texture holds).  A CTA stages
// one input tile + halo in shared memory -- by TMA (cp.async.bulk.tensor.3d into a double buffer, the
// next tile in flight while the current one is computed; out-of-image halo texels, which TMA zero-fills,
// are patched in shared memory to the reference's clamp-to-edge) when the plane is 16-byte aligned, else
// by plain clamped loads -- then every thread
// walks a vertical strip of P pixels keeping the (P + 2o) x n luma window in registers, so that
// gradients, and the (0.1+l)^32 / (1.1-l)^32 anti-ringing powers are computed once per source
// pixel and reused by every output pixel that taps them.  All 4 (or 9) sub-pixel phases are
// written interleaved with 8-byte coalesced streaming stores.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tma.cuh"

namespace mpvp {

struct LiteArgs {
  const void* __restrict__ in;   // planes of format io.in_fmt
  void* __restrict__ out;        // planes of format io.out_fmt
  IoFmt io;
  const float4* __restrict__ lut;  // [rows][LW]
  const uint2* __restrict__ lut_half;  // same texels as 4 x binary16 (present when the LUT was rounded to fp16)
  int32_t* __restrict__ bucket;
  int n, h, w;
  int64_t in_sn, in_sc, in_sy, out_sn, out_sc, out_sy;
  int tiles_x, tiles_y;
  long long total_tiles;
  float ar_strength;
  mpvp_key_params key;
};

// the -ar kernels of the 2x family are instantiated in their own translation unit (ravu_lite_ar.cu)
int ravu_lite_ar_dispatch(const LiteArgs& a, int radius, int device, cudaStream_t st);

namespace {


template <int R>
struct LiteGeom {
  static constexpr int N = 2 * R - 1;                 // window side
  static constexpr int O = R - 1;                     // halo
  static constexpr int G = (R == 4) ? 5 : 3;          // gradient square side
  static constexpr int TAPS = N * N;
  static constexpr int HALF = (TAPS - 1) / 2;
};

// Is window tap t (x-major) inside the anti-ringing diamond dx^2 + dy^2 <= 4 ?
template <int R>
__device__ __forceinline__ constexpr bool ar_tap(int t) {
  const int dx = t / (2 * R - 1) - (R - 1), dy = t % (2 * R - 1) - (R - 1);
  return dx * dx + dy * dy <= 4;
}

// experiment knobs (tools/build_variant.py): defaults are the measured best
#ifndef MPVP_X_PACKCONV_AR
#define MPVP_X_PACKCONV_AR 0   // convolution sums as packed FFMA2 in the -ar kernels (FMA-pipe bound: no gain, 3.46 -> 3.60 ms)
#endif
#ifndef MPVP_X_LITE_BLOCKS
#define MPVP_X_LITE_BLOCKS 2   // CTAs per SM of the plain (non -ar) r2 / r3 2x kernels
#endif
#ifndef MPVP_X_LITE_P
#define MPVP_X_LITE_P 4
#endif
#ifndef MPVP_X_LITE_STRIPS
#define MPVP_X_LITE_STRIPS 2
#endif
#ifndef MPVP_X_R4_BLOCKS
#define MPVP_X_R4_BLOCKS 1   // CTAs per SM for the r4 2x kernels when the LUT is binary16 (57.6 KB)
#endif
#ifndef MPVP_X_AR4_STRIPS
#define MPVP_X_AR4_STRIPS 4
#endif
#ifndef MPVP_X_AR3_BLOCKS
#define MPVP_X_AR3_BLOCKS 2
#endif
#ifndef MPVP_X_AR3_P
#define MPVP_X_AR3_P 2
#endif
#ifndef MPVP_X_AR3_STRIPS
#define MPVP_X_AR3_STRIPS 5
#endif

constexpr int kTW = 64;       // tile width  (input pixels)
constexpr int kThreads = 256; // 64 x 4 threads

// acc + v * s with s broadcast to both lanes: one FFMA2 (the scalar rides in the instruction's .F32 operand form)
__device__ __forceinline__ float2 fma2s(float2 v, float s, float2 acc) { return __ffma2_rn(v, make_float2(s, s), acc); }

__device__ __forceinline__ float rgb_luma709(float r, float g, float b) {
  // dot(rgb, color_primary), left to right, no contraction (compute/ravu-3x-r2-rgb.hook:23,33)
  return __fadd_rn(__fadd_rn(__fmul_rn(r, 0.2126f), __fmul_rn(g, 0.7152f)), __fmul_rn(b, 0.0722f));
}

// C = colour channels (1 or 3; 3 only for SCALE == 3), KEYMODE: 0 luma, 1 yuv (key = channel 0), 2 rgb
// LH: LUT kept in shared memory as 4 x binary16 per texel (exact: the texels ARE binary16 values, App. D.1).  A warp's
// 32 lanes gather 32 different LUT rows, so the gather is bank-conflict bound: 8-byte texels need half the
// shared-memory wavefronts of 16-byte ones (68 vs 133 per 13-texel row on the config-2 planes).
// OF32: the output planes are float32 at compile time (the runtime format switch at the store costs 2.8 % on the
// register-bound -ar kernel: 3.53 vs 3.43 ms); false = any mpvp_io output format.
// RAWB: bytes per element of the plane the TMA engine fetches (4: float32 straight into the tile; 1 / 2: uint8 / uint16
// video planes into a raw double buffer, converted to the float tile -- raw / in_max, one division per source
// pixel -- by a pass over the tile once the mbarrier fires).
template <int R, bool AR, int SCALE, int P, int STRIPS, int C, int KEYMODE, bool FASTKEY, bool TMA, bool LH, bool OF32, int RAWB = 4>
__global__ void __launch_bounds__(kThreads, (R == 4 ? ((LH && SCALE == 2) ? MPVP_X_R4_BLOCKS : 1) : ((AR && R == 3) ? MPVP_X_AR3_BLOCKS : ((!AR && SCALE == 2) ? MPVP_X_LITE_BLOCKS : 2))))
ravu_lite_kernel(const __grid_constant__ LiteArgs A, const __grid_constant__ CUtensorMap tmap) {
  static_assert(!TMA || C == 1, "TMA staging is implemented for single-plane inputs");
  static_assert(C == 1 || SCALE == 3, "3-channel planes exist only for RAVU-3x");
  using Gm = LiteGeom<R>;
  constexpr int N = Gm::N, O = Gm::O, G = Gm::G, TAPS = Gm::TAPS, HALF = Gm::HALF;
  constexpr int LW = (SCALE == 2) ? (TAPS + 1) / 2 : (TAPS + 1);
  constexpr int LWP = LW | 1;  // odd float4 pitch in shared memory (3x: 10/26/50 -> 11/27/51): no row-to-row bank aliasing
  constexpr int ROWS = (SCALE == 2) ? 288 : 216;
  constexpr int TR = kThreads / kTW;          // thread rows
  constexpr int TH = TR * P * STRIPS;         // tile height
  // TMA needs the box to START on a 16-byte boundary and its rows to be multiples of 16 bytes (an x origin of
  // x0 - O is rejected as an illegal instruction, tools/tma_min.cu): the TMA box starts 4 texels left of the tile
  constexpr int XO = TMA ? 4 : O;             // staged columns left of the tile
  constexpr int SW = TMA ? (kTW + 8) : ((kTW + 2 * O + 3) & ~3);  // staged row pitch
  constexpr int SH = TH + 2 * O;
  constexpr int PLANE = SW * SH;              // plane 0 = key plane, planes 1..3 = colours (C == 3)
  constexpr int TBUF = ((PLANE * 4 + 127) / 128) * 32;  // floats per TMA buffer (128-byte aligned)
  constexpr bool RAW = TMA && RAWB != 4;
  constexpr int XOR_ = 16 / RAWB;             // raw box: starts 16 bytes left of the tile ...
  constexpr int RW = ((XOR_ + kTW + O + XOR_ - 1) / XOR_) * XOR_;   // ... and spans a multiple of 16 bytes
  constexpr int RBUF = ((RW * SH * RAWB + 127) / 128) * 32;         // floats per raw buffer (128-byte aligned)
  static_assert(!RAW || (C == 1), "raw integer staging is single-plane");
  // anti-ringing: ((0.1+l)^32, (1.1-l)^32, (0.1+l)^33, (1.1-l)^33) of every source pixel the tile's diamonds tap,
  // computed ONCE per source pixel into shared memory (the shader recomputes them per output pixel and tap)
  constexpr int AO = O < 2 ? O : 2;           // diamond reach dx^2 + dy^2 <= 4
  constexpr int PW = kTW + 2 * AO, PH = TH + 2 * AO;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4* s_lut = reinterpret_cast<float4*>(smem_raw);
  uint2* s_luth = reinterpret_cast<uint2*>(smem_raw);
  constexpr int kLutBytes = ((int)(LH ? sizeof(uint2) : sizeof(float4)) * ROWS * LWP + 127) & ~127;
  float* s_tiles = reinterpret_cast<float*>(smem_raw + kLutBytes);
  // RAW: [raw buffer 0][raw buffer 1][one float tile]; float32 TMA: [tile 0][tile 1]; plain staging: [planes]
  float4* s_pow = reinterpret_cast<float4*>(s_tiles + (RAW ? 2 * RBUF + TBUF : (TMA ? 2 * TBUF : PLANE * (C == 1 ? 1 : 4))));
  __shared__ __align__(8) uint64_t s_mbar[2];

  const int tid = threadIdx.x;
  if constexpr (TMA) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_mbar[0])) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&s_mbar[1])) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  if constexpr (LH) {
    for (int i = tid; i < ROWS * LW; i += kThreads) s_luth[(i / LW) * LWP + (i % LW)] = A.lut_half[i];
  } else {
    for (int i = tid; i < ROWS * LW; i += kThreads) s_lut[(i / LW) * LWP + (i % LW)] = A.lut[i];
  }
  __syncthreads();

  const int tx = tid % kTW, tr = tid / kTW;

  // one thread asks the TMA engine for the (SW x SH) box of `tile` (origin may be negative: OOB -> 0)
  const uint64_t tmap_ptr = reinterpret_cast<uint64_t>(&tmap);  // address of the __grid_constant__ parameter itself
  auto tma_issue = [&, tmap_ptr](const TileWalk& tw, int buf) {
    const int tix = tw.tix, tiy = tw.tiy, f = tw.f;
    const uint32_t bar = smem_addr(&s_mbar[buf]);
    const uint32_t dst = smem_addr(s_tiles + buf * (RAW ? RBUF : TBUF));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(RAW ? RW * SH * RAWB : PLANE * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(tmap_ptr), "r"(tix * kTW - (RAW ? XOR_ : XO)), "r"(tiy * TH - O), "r"(f), "r"(bar)
        : "memory");
  };
  TileWalk walk(blockIdx.x, gridDim.x, A.tiles_x, A.tiles_y);   // the tile being computed
  TileWalk ahead = walk;                                        // the tile being fetched (one step ahead)
  if constexpr (TMA) {
    if (tid < 32 && blockIdx.x < A.total_tiles) {
      if (elect_one()) tma_issue(ahead, 0);
    }
  }
  ahead.next();

  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, ++it, walk.next(), ahead.next()) {
    const int tix = walk.tix, tiy = walk.tiy, f = walk.f;
    const int x0 = tix * kTW, y0 = tiy * TH;
    const int64_t src0 = (int64_t)f * A.in_sn;   // element offset of this frame
    float* __restrict__ s_tile = s_tiles + (RAW ? 2 * RBUF : (TMA ? (it & 1) * TBUF : 0));

    if constexpr (TMA) {
      // the other buffer was released by the barrier that ended the previous iteration
      if (tid < 32 && tile + gridDim.x < A.total_tiles) {
        if (elect_one()) tma_issue(ahead, (it + 1) & 1);
      }
      mbar_wait_parity(smem_addr(&s_mbar[it & 1]), (it >> 1) & 1);
      if constexpr (RAW) {
        // raw integer texels -> the float tile (same layout as the float32 TMA tile: XO = 4 columns left of the tile)
        const unsigned char* __restrict__ rb = reinterpret_cast<const unsigned char*>(s_tiles + (it & 1) * RBUF);
        for (int i = tid; i < PLANE; i += kThreads) {
          const int sy = i / SW, sx = i - sy * SW;
          const int ri = sy * RW + sx + (XOR_ - XO);
          const float rawv = RAWB == 1 ? (float)rb[ri] : (float)reinterpret_cast<const unsigned short*>(rb)[ri];
          s_tile[i] = __fdiv_rn(rawv, A.io.in_max);
        }
        __syncthreads();
      }
      const bool edge = x0 - XO < 0 || y0 - O < 0 || x0 - XO + SW > A.w || y0 - O + SH > A.h;
      if (edge) {  // CTA-uniform: replicate the border (clamp-to-edge) over the zero-filled texels
        for (int i = tid; i < PLANE; i += kThreads) {
          const int sy = i / SW, sx = i - sy * SW;
          const int gx = x0 - XO + sx, gy = y0 - O + sy;
          if (gy >= 0 && gy < A.h && (gx < 0 || gx >= A.w)) s_tile[i] = s_tile[sy * SW + clampi(gx, 0, A.w - 1) - (x0 - XO)];
        }
        __syncthreads();
        for (int i = tid; i < PLANE; i += kThreads) {
          const int sy = i / SW, sx = i - sy * SW;
          const int gy = y0 - O + sy;
          if (gy < 0 || gy >= A.h) s_tile[i] = s_tile[(clampi(gy, 0, A.h - 1) - (y0 - O)) * SW + sx];
        }
        __syncthreads();
      }
    } else {
    __syncthreads();  // previous tile fully consumed
    dispatch_in_fmt(A.io.in_fmt, [&](auto ftag) {
      constexpr int FMT = decltype(ftag)::value;
      for (int i = tid; i < SW * SH; i += kThreads) {
        const int sy = i / SW, sx = i - sy * SW;
        const int gx = clampi(x0 + sx - XO, 0, A.w - 1);
        const int gy = clampi(y0 + sy - O, 0, A.h - 1);
        const int64_t off = src0 + (int64_t)gy * A.in_sy + gx;
        if constexpr (C == 1) {
          s_tile[i] = load_px_t<FMT>(A.in, off, A.io.in_max);
        } else {
          const float c0 = load_px_t<FMT>(A.in, off, A.io.in_max);
          const float c1 = load_px_t<FMT>(A.in, off + A.in_sc, A.io.in_max);
          const float c2 = load_px_t<FMT>(A.in, off + 2 * A.in_sc, A.io.in_max);
          s_tile[i] = (KEYMODE == 2) ? rgb_luma709(c0, c1, c2) : c0;
          s_tile[PLANE + i] = c0;
          s_tile[2 * PLANE + i] = c1;
          s_tile[3 * PLANE + i] = c2;
        }
      }
    });
    __syncthreads();
    // the staging above is synchronous (no TMA on this path): pull the rows of the NEXT tile into L2 now, one
    // 128-byte line per thread, so that its staging loads find them there
    if (tile + gridDim.x < A.total_tiles) {
      const int eb = fmt_bytes(A.io.in_fmt);
      constexpr int LPR = (SW * 4 + 127) / 128 + 1;            // lines per staged row (upper bound)
      for (int i = tid; i < SH * LPR * C; i += kThreads) {
        const int c = i / (SH * LPR), r = i - c * (SH * LPR);
        const int sy = r / LPR, ln = r - sy * LPR;
        const int gy = clampi(ahead.tiy * TH + sy - O, 0, A.h - 1);
        const int gx = ahead.tix * kTW - XO + ln * (128 / eb);
        if (gx < A.w) {
          const char* ptr = static_cast<const char*>(A.in) +
                            ((int64_t)ahead.f * A.in_sn + c * A.in_sc + (int64_t)gy * A.in_sy + (gx < 0 ? 0 : gx)) * eb;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        }
      }
    }
    }

    if constexpr (AR) {
      for (int i = tid; i < PW * PH; i += kThreads) {
        const int sy = i / PW, sx = i - sy * PW;
        const float lv = s_tile[(sy + O - AO) * SW + sx + (XO - AO)];
        const float c = 0.1f + lv, dd = 1.1f - lv;
        const float pc = pow32(c), pd = pow32(dd);
        s_pow[i] = make_float4(pc, pd, pc * c, pd * dd);
      }
      __syncthreads();
    }

    const int x = x0 + tx;
#pragma unroll 1
    for (int s = 0; s < STRIPS; ++s) {
      const int ly0 = (s * TR + tr) * P;  // first tile row of this strip
      const int yb = y0 + ly0;
      if (x >= A.w || yb >= A.h) continue;

      // register window: l[yy][xx] = luma at (x + xx - O, yb + yy - O)
      float l[P + 2 * O][N];
#pragma unroll
      for (int yy = 0; yy < P + 2 * O; ++yy)
#pragma unroll
        for (int xx = 0; xx < N; ++xx) l[yy][xx] = s_tile[(ly0 + yy) * SW + tx + xx + (XO - O)];

#pragma unroll
      for (int p = 0; p < P; ++p) {
        const int y = yb + p;
        const bool live = y < A.h;  // rows past the image are computed (from clamped data) but not stored,
                                    // so the P pixels of a strip form one basic block for the scheduler
        // window sample (i, j) with i <-> dx, j <-> dy
        auto Wn = [&](int i, int j) { return l[p + j][i]; };
        const int row = ravu_key2<STENCIL_LITE, N, G, (SCALE == 2 ? 3 : 2), FASTKEY>(A.key, Wn);
        if (A.bucket && live) A.bucket[((int64_t)f * A.h + y) * A.w + x] = row;
        // texel t of this pixel's LUT row
        auto texel = [&](int t) -> float4 {
          if constexpr (LH) {
            const uint2 u = s_luth[row * LWP + t];
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
            return make_float4(a.x, a.y, b.x, b.y);
          } else {
            return s_lut[row * LWP + t];
          }
        };

        if constexpr (SCALE == 2) {
          // packed f32x2 accumulators (FFMA2 with a broadcast scalar operand): q01/q23 take the tap, the
          // mirrored tap uses the same texel reversed (w.wzyx), so it accumulates into (phase 3, 2) / (1, 0)
          float2 q01 = make_float2(0.f, 0.f), q23 = q01, m32 = q01, m10 = q01;
          // (hi, lo)[c] and (hi2, lo2)[c] as packed f32x2 accumulators: one FFMA2 updates both
          float2 hl[4], hl2[4];
          if constexpr (AR) {
#pragma unroll
            for (int c = 0; c < 4; ++c) hl[c] = hl2[c] = make_float2(0.f, 0.f);
          }
          [[maybe_unused]] const float4* __restrict__ prow = s_pow + (ly0 + p + AO) * PW + tx + AO;
#pragma unroll
          for (int t = 0; t <= HALF; ++t) {
            const float4 w = texel(t);
            const float la = Wn(t / N, t % N);
            if constexpr (AR && !MPVP_X_PACKCONV_AR) {
              q01.x = fmaf(la, w.x, q01.x); q01.y = fmaf(la, w.y, q01.y); q23.x = fmaf(la, w.z, q23.x); q23.y = fmaf(la, w.w, q23.y);
            } else {
              q01 = fma2s(make_float2(w.x, w.y), la, q01);
              q23 = fma2s(make_float2(w.z, w.w), la, q23);
            }
            if (t < HALF) {
              const float lb = Wn((TAPS - 1 - t) / N, (TAPS - 1 - t) % N);
              if constexpr (AR && !MPVP_X_PACKCONV_AR) {
                m32.x = fmaf(lb, w.x, m32.x); m32.y = fmaf(lb, w.y, m32.y); m10.x = fmaf(lb, w.z, m10.x); m10.y = fmaf(lb, w.w, m10.y);
              } else {
                m32 = fma2s(make_float2(w.x, w.y), lb, m32);
                m10 = fma2s(make_float2(w.z, w.w), lb, m10);
              }
            }
            if constexpr (AR) {
              if (ar_tap<R>(t)) {
                const float g0 = fmaxf(w.x, 0.f), g1 = fmaxf(w.y, 0.f), g2 = fmaxf(w.z, 0.f), g3 = fmaxf(w.w, 0.f);
                const float2 G[4] = {make_float2(g0, g0), make_float2(g1, g1), make_float2(g2, g2), make_float2(g3, g3)};
                {
                  const float4 pa = prow[(t % N - O) * PW + (t / N - O)];
                  const float2 a1 = make_float2(pa.x, pa.y), a2 = make_float2(pa.z, pa.w);
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    hl[c] = __ffma2_rn(a1, G[c], hl[c]);
                    hl2[c] = __ffma2_rn(a2, G[c], hl2[c]);
                  }
                }
                if (t < HALF) {
                  const float4 pb = prow[-(t % N - O) * PW - (t / N - O)];
                  const float2 b1 = make_float2(pb.x, pb.y), b2 = make_float2(pb.z, pb.w);
#pragma unroll
                  for (int c = 0; c < 4; ++c) {
                    hl[c] = __ffma2_rn(b1, G[3 - c], hl[c]);
                    hl2[c] = __ffma2_rn(b2, G[3 - c], hl2[c]);
                  }
                }
              }
            }
          }
          float res[4] = {q01.x + m10.y, q01.y + m10.x, q23.x + m32.y, q23.y + m32.x};
          if constexpr (AR) {
            const float st = A.ar_strength;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // (hi2/hi - 0.1, 1.1 - lo2/lo)
              const float2 q = __fmul2_rn(hl2[c], make_float2(__fdividef(1.0f, hl[c].x), __fdividef(1.0f, hl[c].y)));
              const float2 hv = __ffma2_rn(q, make_float2(1.0f, -1.0f), make_float2(-0.1f, 1.1f));
              const float cl = fminf(fmaxf(res[c], hv.y), hv.x);
              res[c] = res[c] * (1.0f - st) + cl * st;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) res[c] = fminf(fmaxf(res[c], 0.f), 1.f);
          }
          // phase c -> (2x + c/2, 2y + c%2)
          const int64_t o = (int64_t)f * A.out_sn + (int64_t)(2 * y) * A.out_sy + 2 * x;
          if (live) {
            const int ofmt = OF32 ? MPVP_FMT_F32 : A.io.out_fmt;
            store_px2(A.out, o, res[0], res[2], ofmt, A.io.out_max);
            store_px2(A.out, o + A.out_sy, res[1], res[3], ofmt, A.io.out_max);
          }
        } else {
          // RAVU-3x: two texels per tap, res0 -> phases 0..3, res1 -> phases 5..8, centre copied
#pragma unroll
          for (int c = 0; c < C; ++c) {
            // colour sample of channel c at window tap (i, j)
            auto Cn = [&](int i, int j) -> float {
              if constexpr (C == 1) return l[p + j][i];
              else return s_tile[(1 + c) * PLANE + (ly0 + p + j) * SW + tx + i + (XO - O)];
            };
            float2 a01 = make_float2(0.f, 0.f), a23 = a01, b01 = a01, b23 = a01, ar32 = a01, ar10 = a01, br32 = a01, br10 = a01;
#pragma unroll
            for (int t = 0; t <= HALF; ++t) {
              const float4 w0 = texel(2 * t), w1 = texel(2 * t + 1);
              const float la = Cn(t / N, t % N);
              a01 = fma2s(make_float2(w0.x, w0.y), la, a01); a23 = fma2s(make_float2(w0.z, w0.w), la, a23);
              b01 = fma2s(make_float2(w1.x, w1.y), la, b01); b23 = fma2s(make_float2(w1.z, w1.w), la, b23);
              if (t < HALF) {
                const float lb = Cn((TAPS - 1 - t) / N, (TAPS - 1 - t) % N);
                ar32 = fma2s(make_float2(w1.x, w1.y), lb, ar32); ar10 = fma2s(make_float2(w1.z, w1.w), lb, ar10);
                br32 = fma2s(make_float2(w0.x, w0.y), lb, br32); br10 = fma2s(make_float2(w0.z, w0.w), lb, br10);
              }
            }
            const float v[9] = {a01.x + ar10.y, a01.y + ar10.x, a23.x + ar32.y, a23.y + ar32.x, -1.f,
                                b01.x + br10.y, b01.y + br10.x, b23.x + br32.y, b23.y + br32.x};
            const int64_t o = (int64_t)f * A.out_sn + c * A.out_sc + (int64_t)(3 * y) * A.out_sy + 3 * x;
#pragma unroll
            for (int q = 0; q < 9; ++q) {
              const int i = q / 3, j = q % 3;  // imageStore(gid*3 + ivec2(i, j)): x offset i, y offset j
              const float val = (q == 4) ? Cn(O, O) : fminf(fmaxf(v[q], 0.f), 1.f);
              if (live) store_px(A.out, o + (int64_t)j * A.out_sy + i, val, OF32 ? MPVP_FMT_F32 : A.io.out_fmt, A.io.out_max);
            }
          }
        }
      }
    }
    if constexpr (TMA) {
      // this buffer may be refilled by the TMA (async proxy) issued at the top of the next iteration
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }
  }
}

template <int R, bool AR, int SCALE, int P, int STRIPS, int C, int KEYMODE, bool FASTKEY, bool LH, bool OF32>
int launch_lite_impl(const LiteArgs& a0, int device, cudaStream_t stream) {
  using Gm = LiteGeom<R>;
  constexpr int LW = ((SCALE == 2) ? (Gm::TAPS + 1) / 2 : (Gm::TAPS + 1)) | 1;  // padded pitch
  constexpr int ROWS = (SCALE == 2) ? 288 : 216;
  constexpr int TH = (kThreads / kTW) * P * STRIPS;
  constexpr int SW = (kTW + 2 * Gm::O + 3) & ~3, SH = TH + 2 * Gm::O;   // plain-load staging
  constexpr int SWT = kTW + 8;                                           // TMA box width (starts at x0 - 4)
  constexpr int TBUF = ((SWT * SH * 4 + 127) / 128) * 32;
  LiteArgs a = a0;
  a.tiles_x = (a.w + kTW - 1) / kTW;
  a.tiles_y = (a.h + TH - 1) / TH;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * a.n;
  MPVP_REQUIRE(a.total_tiles < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", a.total_tiles);
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  bool use_tma = false;
  int rawb = 4;   // bytes per element fetched by TMA
  if constexpr (C == 1) {
    if (a.io.in_fmt == MPVP_FMT_F32) {
      use_tma = make_plane_tmap(&tmap, a.in, 4, a.w, a.h, a.n, a.in_sy, a.in_sn, SWT, SH);
    } else if constexpr (FASTKEY && LH && !OF32) {   // integer video planes: raw TMA fetch + conversion pass
      if (a.io.in_fmt == MPVP_FMT_U8 || a.io.in_fmt == MPVP_FMT_U16) {
        rawb = a.io.in_fmt == MPVP_FMT_U8 ? 1 : 2;
        const int xor_ = 16 / rawb, rw = ((xor_ + kTW + Gm::O + xor_ - 1) / xor_) * xor_;
        use_tma = make_plane_tmap(&tmap, a.in, rawb, a.w, a.h, a.n, a.in_sy, a.in_sn, rw, SH);
        if (!use_tma) rawb = 4;
      }
    }
  }
  constexpr int AO = Gm::O < 2 ? Gm::O : 2;
  constexpr size_t kPow = AR ? sizeof(float4) * (kTW + 2 * AO) * (TH + 2 * AO) : 0;  // anti-ringing power tile
  size_t tiles_bytes = sizeof(float) * SW * SH * (C == 1 ? 1 : 4);
  if (use_tma) {
    tiles_bytes = sizeof(float) * 2 * TBUF;
    if (rawb != 4) {
      const int xor_ = 16 / rawb, rw = ((xor_ + kTW + Gm::O + xor_ - 1) / xor_) * xor_;
      tiles_bytes = sizeof(float) * (2 * (size_t)(((rw * SH * rawb + 127) / 128) * 32) + TBUF);
    }
  }
  const size_t smem = ((((LH ? sizeof(uint2) : sizeof(float4)) * ROWS * LW) + 127) & ~(size_t)127) + tiles_bytes + kPow;
  auto kern = ravu_lite_kernel<R, AR, SCALE, P, STRIPS, C, KEYMODE, FASTKEY, false, LH, OF32>;
  if constexpr (C == 1) {
    if (use_tma) kern = ravu_lite_kernel<R, AR, SCALE, P, STRIPS, C, KEYMODE, FASTKEY, true, LH, OF32>;
    if constexpr (FASTKEY && LH && !OF32) {
      if (use_tma && rawb == 1) kern = ravu_lite_kernel<R, AR, SCALE, P, STRIPS, C, KEYMODE, FASTKEY, true, LH, OF32, 1>;
      if (use_tma && rawb == 2) kern = ravu_lite_kernel<R, AR, SCALE, P, STRIPS, C, KEYMODE, FASTKEY, true, LH, OF32, 2>;
    }
  }
  MPVP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MPVP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) {
    set_error("ravu_lite kernel does not fit on an SM (smem %zu B)", smem);
    return MPVP_E_UNSUPPORTED;
  }
  long long grid = (long long)sm_count(device) * per_sm;
  if (grid > a.total_tiles) grid = a.total_tiles;
  grid = cap_grid(grid);
  if (grid < 1) return MPVP_OK;
  kern<<<(unsigned)grid, kThreads, smem, stream>>>(a, tmap);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}

// MPVP_KEY=exact selects the op-for-op key (sqrt, division, atan2f) instead of the equivalent
// comparison form; used to A/B the two on the device.
bool exact_key() {
  static const bool v = [] {
    const char* e = getenv("MPVP_KEY");
    return e && e[0] == 'e';
  }();
  return v;
}

// MPVP_LUT_SMEM=fp32 keeps 16-byte texels in shared memory (A/B switch)
bool half_lut_enabled() {
  static const bool v = [] {
    const char* e = getenv("MPVP_LUT_SMEM");
    return !(e && e[0] == 'f' && e[2] == '3');
  }();
  return v;
}

template <int R, bool AR, int SCALE, int P, int STRIPS, int C = 1, int KEYMODE = 0>
int launch_lite(const LiteArgs& a, int device, cudaStream_t stream) {
  if (exact_key()) return launch_lite_impl<R, AR, SCALE, P, STRIPS, C, KEYMODE, false, false, false>(a, device, stream);
  if (a.lut_half && half_lut_enabled()) {
    if (a.io.out_fmt == MPVP_FMT_F32) return launch_lite_impl<R, AR, SCALE, P, STRIPS, C, KEYMODE, true, true, true>(a, device, stream);
    return launch_lite_impl<R, AR, SCALE, P, STRIPS, C, KEYMODE, true, true, false>(a, device, stream);
  }
  return launch_lite_impl<R, AR, SCALE, P, STRIPS, C, KEYMODE, true, false, false>(a, device, stream);
}

int check_common(const mpvp_weights* lut, const mpvp_key_params* key, int radius, const void* in, const void* out,
                 int n, int h, int w, int want_w, int want_h, int want_gauss) {
  MPVP_REQUIRE(lut && lut->kind == 0 && lut->lut, "lut handle is null or not a LUT");
  MPVP_REQUIRE(key, "key params are null");
  MPVP_REQUIRE(radius >= 2 && radius <= 4, "radius %d not in {2,3,4}", radius);
  MPVP_REQUIRE(in && out, "null frame pointer");
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad frame geometry n=%d h=%d w=%d", n, h, w);
  MPVP_REQUIRE(lut->lut_w == want_w && lut->lut_h == want_h, "LUT is %dx%d, expected %dx%d", lut->lut_w, lut->lut_h,
               want_w, want_h);
  MPVP_REQUIRE(key->n_gauss == want_gauss, "key params carry %d Gaussian weights, expected %d", key->n_gauss,
               want_gauss);
  return MPVP_OK;
}

}  // namespace
}  // namespace mpvp

