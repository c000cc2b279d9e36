// RAVU-Zoom(-AR): arbitrary-ratio upscale.
//
//   RAVU-Zoom      ravu-zoom-r2.hook:15-134, ravu-zoom-r3.hook:15-180
//   RAVU-Zoom-AR   ravu-zoom-ar-r2.hook:15-208  (3-channel mat4x3 form: ravu-zoom-ar-r2-rgb.hook:153-223)
//
// What the shader does per OUTPUT pixel: pos = HOOKED_pos * HOOKED_size; subpix = fract(pos - 0.5); window of (2r)^2 source
// texels around floor(pos - 0.5); key -> LUT row; weights = LINEAR-filtered fetches of a [288*9][B*9] LUT at
// (blk/B + LUTPOS(subpix.x)/B, row/288 + LUTPOS(subpix.y)/288) for the first half of the taps and at the mirrored
// sub-pixel position for the second half (ravu-zoom-r3.hook:23-31,128-177).
//
// Two device paths, same arithmetic contract:
//
//  * PHASE path (rational ratios with few sub-pixel phases: 2x, 3x, 3/2, 4/3 ...).  The sub-pixel phase of an output
//    column / row takes only a handful of values (up to fp32 rounding of `pos`), so the bilinear LUT blend is done ONCE
//    per (phase class pair, LUT row, tap) by a setup kernel into a phase LUT (cached per geometry on the weight handle)
//    instead of 10 four-texel gathers per output pixel.  Launch = a key pre-pass (one key per SOURCE cell -> uint16 map;
//    every output pixel whose floor(pos - 0.5) coincides shares window and bucket) + a convolution kernel whose
//    persistent CTAs each own a contiguous run of (class pair, member tile) work items with the class pair's phase LUT
//    ([288][taps] float32) resident in shared memory -- the structure of the RAVU-Lite / RAVU-3x kernels.
//  * GENERAL path (any ratio): one thread per output pixel, explicit fp32 blend of the four LUT texels.
//
// Position arithmetic follows SURVEY.md App. D.6 exactly (pos = ((o + 0.5) / O) * I in fp32, then subpix = fract(pos - 0.5)):
// at integer ratios a 1-ulp difference moves the window.  The LUT coordinate arithmetic also follows the shader's fp32
// operation order (coordinate = blk/B + LUTPOS/B resp. row/288 + LUTPOS/288, then * texture size - 0.5 as the LINEAR
// sampler does): the anti-ringing soft-min/max raises tap values to the 32nd power, which makes its result hypersensitive
// to blend weights that are exactly zero in the reference (a 3e-8 weight on a tap 2x brighter than the rest moves the
// clamp bound by 1e-2), so "equal up to rounding" is not good enough there.
// The texture-unit fetch (8-bit blend weights, what a GL driver does) is kept behind MPVP_ZOOM_TEX=1 only: measured
// 1.16e-3 max abs error against the oracle on a 1280x720 -> 3840x2160 frame, outside the 1e-3 parity bound.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace mpvp {
namespace {

struct ClassPair {
  int xsoff, ysoff;            // first x / y member segment of the two classes in the plan's segment arrays
  int tiles_x, tiles_y;        // member tiles per frame = segments of the x class  x  segments of the y class
  int tile_start;              // first work item of this class pair PER FRAME (work items are class-pair major; the
                               // class pair's items of an n-frame batch start at tile_start * n)
  int pad;
};

struct ZoomArgs {
  const void* __restrict__ in;   // planes of format io.in_fmt
  void* __restrict__ out;        // planes of format io.out_fmt
  IoFmt io;
  const void* __restrict__ lut;     // float4 texels, or 4 x binary16 texels (LUTH)
  const void* __restrict__ lut_ar;
  cudaTextureObject_t tex, tex_ar;  // the same LUTs behind LINEAR-filtering texture objects (TEXF kernels)
  int32_t* __restrict__ bucket;  // [n][oh][ow] or null
  int n, h, w, oh, ow;
  int64_t in_sn, in_sc, in_sy, out_sn, out_sc, out_sy;
  int tiles_x, tiles_y;
  long long total_tiles;
  float ar_strength;
  mpvp_key_params key;
  // phase path
  const int* __restrict__ xo; const int* __restrict__ xb;   // output coordinate / base texel of every class member
  const int* __restrict__ yo; const int* __restrict__ yb;
  const int2* __restrict__ xseg; const int2* __restrict__ yseg;   // member segments (first member, count): one tile side each
  const ClassPair* __restrict__ cps;
  int ncp;
  int tiles_per_frame, fgroup;      // member tiles of one frame (all class pairs); frames per L2-resident group
  const long long* __restrict__ cta_begin;   // [grid + 1] first work item of every CTA (cost-balanced), or null
  const float* __restrict__ plut;   // [ncp][288][PL]
  unsigned short* __restrict__ kmap;  // [n][h + 1][w + 1] LUT row of the source cell with base texel (x - 1, y - 1)
  int sw, sh;                       // staged source rectangle (pitch, rows) the plan needs
};

constexpr int kTOW = 32, kTOH = 32, kNT = 256;  // general path: output tile; each thread owns one column and kTOH/8 rows

// Canonical position arithmetic: base texel index and sub-pixel phase of output coordinate o.
__device__ __forceinline__ void zoom_pos(int o, int O, int I, int& base, float& sub) {
  const float pos = __fmul_rn(__fdiv_rn(__fadd_rn((float)o, 0.5f), (float)O), (float)I);
  const float t = __fsub_rn(pos, 0.5f);
  sub = __fsub_rn(t, floorf(t));
  base = (int)floorf(__fsub_rn(pos, sub));
}
// the same on the host (x86-64 SSE arithmetic is IEEE fp32 op by op; volatile keeps every rounding)
void zoom_pos_host(int o, int O, int I, int& base, float& sub) {
  volatile float a = (float)o + 0.5f;
  volatile float q = a / (float)O;
  volatile float pos = q * (float)I;
  volatile float t = pos - 0.5f;
  volatile float fl = floorf(t);
  volatile float s = t - fl;
  volatile float d = pos - s;
  sub = s;
  base = (int)floorf(d);
}

// LUTPOS(x, 9) = mix(0.5/9, 1 - 0.5/9, x) = a*(1-x) + b*x   (ravu-zoom-r2.hook:23)
__device__ __forceinline__ float lutpos9(float x) {
  const float a = __fdiv_rn(0.5f, 9.0f);
  const float b = __fsub_rn(1.0f, a);
  return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, x)), __fmul_rn(b, x));
}

// The LINEAR sampler at normalised coordinate c of an axis with `size` texels: texel index floor(c*size - 0.5) and the
// blend weight of its successor, in fp32 exactly as the shader + sampler arithmetic of the oracle evaluates them.
__device__ __forceinline__ void lin_coord(float c, int size, int& i0, float& f) {
  const float u = __fsub_rn(__fmul_rn(c, (float)size), 0.5f);
  const float u0 = floorf(u);
  i0 = (int)u0;
  f = __fsub_rn(u, u0);
}

// block offset literal of the shader: vec2(0.2, coord_y) etc. = float(blk / B)
template <int B>
__device__ __forceinline__ float blk_off(int blk) {
  return (float)((double)blk / (double)B);
}

// one LUT texel as float4, from fp32 or binary16 storage
template <bool LUTH>
__device__ __forceinline__ float4 lut_texel(const void* __restrict__ lut, int idx) {
  if constexpr (LUTH) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(lut) + idx);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(a.x, a.y, b.x, b.y);
  } else {
    return __ldg(reinterpret_cast<const float4*>(lut) + idx);
  }
}

// two-stage lerp of the sampler: (t00 (1-fu) + t10 fu) (1-fv) + (t01 (1-fu) + t11 fu) fv
#ifndef MPVP_X_ZOOM_FMA_LERP
#define MPVP_X_ZOOM_FMA_LERP 1   // second product of every blend fused into the add: -6 % on the general path
#endif
__device__ __forceinline__ float lerp2(float t00, float t10, float t01, float t11, float fu, float fv) {
  const float gu = __fsub_rn(1.0f, fu), gv = __fsub_rn(1.0f, fv);
#if MPVP_X_ZOOM_FMA_LERP
  // a blend weight of exactly 0 still returns the other texel exactly (the property the anti-ringing LUT needs)
  const float top = __fmaf_rn(t10, fu, __fmul_rn(t00, gu));
  const float bot = __fmaf_rn(t11, fu, __fmul_rn(t01, gu));
  return __fmaf_rn(bot, fv, __fmul_rn(top, gv));
#else
  const float top = __fadd_rn(__fmul_rn(t00, gu), __fmul_rn(t10, fu));
  const float bot = __fadd_rn(__fmul_rn(t01, gu), __fmul_rn(t11, fu));
  return __fadd_rn(__fmul_rn(top, gv), __fmul_rn(bot, fv));
#endif
}

__device__ __forceinline__ float key_luma709(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(r, 0.2126f), __fmul_rn(g, 0.7152f)), __fmul_rn(b, 0.0722f));
}

// =====================================================================================================================
// GENERAL path: one thread per output pixel
// =====================================================================================================================
// TEXF (opt-in): the LUT fetch is ONE texture instruction (hardware bilinear blend of the four binary16 texels with 8-bit
// blend weights) instead of four gathered loads and the fp32 blend.
#ifndef MPVP_X_ZOOM_POWT
#define MPVP_X_ZOOM_POWT 0   // general path, anti-ringing: powers of the staged tile kept as float4 in shared memory (measured)
#endif
template <int R, int C, int KEYMODE, bool AR, bool LUTH, bool TEXF>
__global__ void __launch_bounds__(kNT) ravu_zoom_kernel(const __grid_constant__ ZoomArgs A) {
  constexpr int N = 2 * R, TAPS = N * N, G = 4;
  constexpr int B = (TAPS / 2 + 3) / 4;   // LUT blocks per row group (2 for r2, 5 for r3)
  constexpr int LWt = B * 9, LHt = 288 * 9;
  constexpr int CW = kTOW + 2, CH = kTOH + 2;          // source cells a tile can touch (the hook only upscales)
  constexpr int SWt = CW + N - 1, SHt = CH + N - 1;    // staged source rectangle
  constexpr int PLANE = SWt * SHt;
  constexpr int NP = (C == 1) ? 1 : 4;                 // plane 0 = key plane, 1..3 = colours
  constexpr int RPT = kTOH / (kNT / kTOW);             // rows per thread

  constexpr bool POWT = MPVP_X_ZOOM_POWT && AR && C == 1;   // anti-ringing powers once per staged source pixel instead of per output pixel and tap
  __shared__ float s_src[NP * PLANE];
  __shared__ float4 s_pow[POWT ? PLANE : 1];
  __shared__ int s_key[CW * CH];
  __shared__ int s_bx[kTOW], s_by[kTOH];               // base texel - first base of the tile
  __shared__ int s_xi[kTOW][2 * B];                    // LUT texel column of (mirrored, blk)
  __shared__ float s_xf[kTOW][2 * B];                  // blend weight of its successor
  __shared__ float s_spy[kTOH][2];                     // LUTPOS(subpix.y) / 288: direct, mirrored

  const int tid = threadIdx.x;
  static_assert(kTOW == 32 && kTOH == 32 && kNT == 256, "the patch mapping below assumes 32x32 tiles and 8 warps");

  TileWalk walk(blockIdx.x, gridDim.x, A.tiles_x, A.tiles_y);
  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, walk.next()) {
    const int tix = walk.tix, tiy = walk.tiy, f = walk.f;
    const int ox0 = tix * kTOW, oy0 = tiy * kTOH;
    const int64_t src0 = (int64_t)f * A.in_sn;

    int bx_first, by_first, bx_last, by_last;
    float dummy;
    zoom_pos(ox0, A.ow, A.w, bx_first, dummy);
    zoom_pos(oy0, A.oh, A.h, by_first, dummy);
    zoom_pos(min(ox0 + kTOW, A.ow) - 1, A.ow, A.w, bx_last, dummy);
    zoom_pos(min(oy0 + kTOH, A.oh) - 1, A.oh, A.h, by_last, dummy);
    const int ncx = bx_last - bx_first + 1, ncy = by_last - by_first + 1;   // cells touched (<= CW, CH)
    const int sx0 = bx_first - (R - 1), sy0 = by_first - (R - 1);

    __syncthreads();
    // ---- per-axis tables -------------------------------------------------------------------------
    if (tid < kTOW) {
      int base;
      float sub;
      zoom_pos(min(ox0 + tid, A.ow - 1), A.ow, A.w, base, sub);
      s_bx[tid] = base - bx_first;
      const float p = lutpos9(sub), ip = __fsub_rn(1.0f, p);
      const float sp[2] = {__fdiv_rn(p, (float)B), __fdiv_rn(ip, (float)B)};
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int blk = 0; blk < B; ++blk) {
          int i0;
          float fr;
          lin_coord(__fadd_rn(blk_off<B>(blk), sp[m]), LWt, i0, fr);
          s_xi[tid][m * B + blk] = i0;
          s_xf[tid][m * B + blk] = fr;
        }
    } else if (tid < kTOW + kTOH) {
      const int j = tid - kTOW;
      int base;
      float sub;
      zoom_pos(min(oy0 + j, A.oh - 1), A.oh, A.h, base, sub);
      s_by[j] = base - by_first;
      const float p = lutpos9(sub), ip = __fsub_rn(1.0f, p);
      s_spy[j][0] = __fdiv_rn(p, 288.0f);
      s_spy[j][1] = __fdiv_rn(ip, 288.0f);
    }
    // ---- stage the source rectangle (clamp-to-edge) ------------------------------------------------
    const int need_w = ncx + N - 1, need_h = ncy + N - 1;
    dispatch_in_fmt(A.io.in_fmt, [&](auto ftag) {
      constexpr int FMT = decltype(ftag)::value;
      for (int i = tid; i < need_w * need_h; i += kNT) {
        const int sy = i / need_w, sx = i - sy * need_w;
        const int gx = clampi(sx0 + sx, 0, A.w - 1), gy = clampi(sy0 + sy, 0, A.h - 1);
        const int64_t off = src0 + (int64_t)gy * A.in_sy + gx;
        const int d = sy * SWt + sx;
        if constexpr (C == 1) {
          const float v = load_px_t<FMT>(A.in, off, A.io.in_max);
          s_src[d] = v;
          if constexpr (POWT) {
            const float cc = 0.1f + v, dd = 1.1f - v;
            const float pc = pow32(cc), pd = pow32(dd);
            s_pow[d] = make_float4(pc, pd, pc * cc, pd * dd);
          }
        } else {
          const float c0 = load_px_t<FMT>(A.in, off, A.io.in_max);
          const float c1 = load_px_t<FMT>(A.in, off + A.in_sc, A.io.in_max);
          const float c2 = load_px_t<FMT>(A.in, off + 2 * A.in_sc, A.io.in_max);
          s_src[d] = (KEYMODE == 2) ? key_luma709(c0, c1, c2) : c0;
          s_src[PLANE + d] = c0;
          s_src[2 * PLANE + d] = c1;
          s_src[3 * PLANE + d] = c2;
        }
      }
    });
    __syncthreads();
    // ---- one key per source cell: every output pixel whose base texel coincides shares window and bucket ----
    for (int i = tid; i < ncx * ncy; i += kNT) {
      const int cy = i / ncx, cx = i - cy * ncx;
      const float* __restrict__ kb = s_src + cy * SWt + cx;
      float ks[TAPS];
#pragma unroll
      for (int t = 0; t < TAPS; ++t) ks[t] = kb[(t % N) * SWt + (t / N)];
      s_key[cy * CW + cx] = ravu_key2<STENCIL_RAVU, N, G, 3, true>(A.key, [&](int ii, int jj) { return ks[ii * N + jj]; });
    }
    __syncthreads();

    // A warp covers an 8x4 patch of output pixels, not a 32x1 row segment: a compact patch spans 3-4 times fewer
    // source cells (hence LUT row groups) than a row segment does.
#pragma unroll 1
    for (int rr = 0; rr < RPT; ++rr) {
      const int patch = rr * (kNT / 32) + (tid >> 5);           // 4 patches across, 8 down
      const int lx = (patch & 3) * 8 + (tid & 7), ly = (patch >> 2) * 4 + ((tid >> 3) & 3);
      const int ox = ox0 + lx, oy = oy0 + ly;
      if (ox >= A.ow || oy >= A.oh) continue;
      const int cbx = s_bx[lx], cby = s_by[ly];
      const int row = s_key[cby * CW + cbx];
      if (A.bucket) A.bucket[((int64_t)f * A.oh + oy) * A.ow + ox] = row;
      const float* __restrict__ kb = s_src + cby * SWt + cbx;  // window tap (0,0)
      const float coord_y = __fdiv_rn((float)row, 288.0f);

      float res[C];
      float hi[C], lo[C], hi2[C], lo2[C];
#pragma unroll
      for (int c = 0; c < C; ++c) res[c] = hi[c] = lo[c] = hi2[c] = lo2[c] = 0.f;

#pragma unroll
      for (int m = 0; m < 2; ++m) {
        // LUT rows: clamp-to-edge on the whole texture like the GL sampler
        int yi;
        float fv;
        lin_coord(__fadd_rn(coord_y, s_spy[ly][m]), LHt, yi, fv);
        const int y0 = clampi(yi, 0, LHt - 1), y1 = clampi(yi + 1, 0, LHt - 1);
#pragma unroll
        for (int blk = 0; blk < B; ++blk) {
          const int xi = s_xi[lx][m * B + blk];
          const float fu = s_xf[lx][m * B + blk];
          const int x0 = clampi(xi, 0, LWt - 1), x1 = clampi(xi + 1, 0, LWt - 1);
          auto fetch = [&](const void* __restrict__ lut) {
            const float4 t00 = lut_texel<LUTH>(lut, y0 * LWt + x0), t10 = lut_texel<LUTH>(lut, y0 * LWt + x1);
            const float4 t01 = lut_texel<LUTH>(lut, y1 * LWt + x0), t11 = lut_texel<LUTH>(lut, y1 * LWt + x1);
            float4 r;
            r.x = lerp2(t00.x, t10.x, t01.x, t11.x, fu, fv);
            r.y = lerp2(t00.y, t10.y, t01.y, t11.y, fu, fv);
            r.z = lerp2(t00.z, t10.z, t01.z, t11.z, fu, fv);
            r.w = lerp2(t00.w, t10.w, t01.w, t11.w, fu, fv);
            return r;
          };
          float4 w4;
          float av[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (TEXF) {
            const float X = (float)xi + fu + 0.5f, Y = (float)yi + fv + 0.5f;
            w4 = tex2D<float4>(A.tex, X, Y);
            if constexpr (AR) {
              const float4 a4 = tex2D<float4>(A.tex_ar, X, Y);
              av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
            }
          } else {
            w4 = fetch(A.lut);
            if constexpr (AR) {
              const float4 a4 = fetch(A.lut_ar);
              av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
            }
          }
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int k = blk * 4 + e;
            if (k < TAPS / 2) {
              const int t = m ? (TAPS - 1 - k) : k;
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const float s = kb[(C == 1 ? 0 : (1 + c)) * PLANE + (t % N) * SWt + (t / N)];
                res[c] = fmaf(s, wv[e], res[c]);
                if constexpr (AR) {
                  float4 pw;
                  if constexpr (POWT) {
                    pw = s_pow[cby * SWt + cbx + (t % N) * SWt + (t / N)];
                  } else {
                    const float cc = 0.1f + s, dd = 1.1f - s;
                    const float pc = pow32(cc), pd = pow32(dd);
                    pw = make_float4(pc, pd, pc * cc, pd * dd);
                  }
                  hi[c] = fmaf(pw.x, av[e], hi[c]);
                  lo[c] = fmaf(pw.y, av[e], lo[c]);
                  hi2[c] = fmaf(pw.z, av[e], hi2[c]);
                  lo2[c] = fmaf(pw.w, av[e], lo2[c]);
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float r = res[c];
        if constexpr (AR) {
          const float hiv = __fdividef(hi2[c], hi[c]) - 0.1f;
          const float lov = 1.1f - __fdividef(lo2[c], lo[c]);
          const float cl = fminf(fmaxf(r, lov), hiv);
          r = r * (1.0f - A.ar_strength) + cl * A.ar_strength;
        } else {
          r = fminf(fmaxf(r, 0.f), 1.f);
        }
        store_px(A.out, (int64_t)f * A.out_sn + c * A.out_sc + (int64_t)oy * A.out_sy + ox, r, A.io.out_fmt, A.io.out_max);
      }
    }
  }
}

template <int R, int C, int KEYMODE, bool AR, bool LUTH, bool TEXF>
int launch_zoom_impl(const ZoomArgs& a0, int device, cudaStream_t stream) {
  ZoomArgs a = a0;
  a.tiles_x = (a.ow + kTOW - 1) / kTOW;
  a.tiles_y = (a.oh + kTOH - 1) / kTOH;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * a.n;
  MPVP_REQUIRE(a.total_tiles < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", a.total_tiles);
  auto kern = ravu_zoom_kernel<R, C, KEYMODE, AR, LUTH, TEXF>;
  int per_sm = 0;
  MPVP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kNT, 0));
  if (per_sm < 1) per_sm = 1;
  long long grid = (long long)sm_count(device) * per_sm;
  if (grid > a.total_tiles) grid = a.total_tiles;
  grid = cap_grid(grid);
  if (grid < 1) return MPVP_OK;
  kern<<<(unsigned)grid, kNT, 0, stream>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}

long long env_int(const char* name, long long dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoll(e) : dflt;
}
bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] != '0';
}

template <int R, int C, int KEYMODE, bool AR>
int launch_zoom_general(const ZoomArgs& a, int device, cudaStream_t stream, bool half_lut) {
  // MPVP_ZOOM_TEX=1: the texture unit does the bilinear blend (8-bit weights: outside the 1e-3 parity bound, opt-in only)
  if (half_lut && env_flag("MPVP_ZOOM_TEX", false) && a.tex && (!AR || a.tex_ar))
    return launch_zoom_impl<R, C, KEYMODE, AR, true, true>(a, device, stream);
  if (half_lut) return launch_zoom_impl<R, C, KEYMODE, AR, true, false>(a, device, stream);
  return launch_zoom_impl<R, C, KEYMODE, AR, false, false>(a, device, stream);
}

// =====================================================================================================================
// PHASE path
// =====================================================================================================================
#ifndef MPVP_X_ZOOM_SPT
#define MPVP_X_ZOOM_SPT 2   // strips per thread and tile: member tiles of 64 x 32, the per-tile set-up amortised over 8 pixels (1 / 2 / 4: 0.933 / 0.873 / 0.882 ms, zoom-r3 x8)
#endif
constexpr int kSPT = MPVP_X_ZOOM_SPT;
constexpr int kPTW = 64, kPTH = 16 * kSPT;   // member tile: 64 x 16 (32) output pixels of one class pair
constexpr int kMaxClasses = 8;                    // per axis
constexpr int kMaxClassesExact = 32;              // per axis when the phases around a LUT node are kept apart (anti-ringing)
constexpr int kMaxStep = 3;                       // base-texel distance of neighbouring class members
constexpr float kClusterGap = 2.5e-4f;            // sub-pixel phases closer than this belong to one class ...
constexpr float kMaxSpread = 2.6e-4f;             // ... whose total spread must stay below this (fp32 noise of `pos`)

// ---- key pre-pass: LUT row of every source cell (base texel (x-1, y-1), x in [0, w], y in [0, h]) -------------------
constexpr int kKTW = 64, kKTH = 16;
// TMA (luma float32 planes): the window tile arrives by cp.async.bulk.tensor into a double buffer, the next tile in flight
// while the keys of the current one are computed; the box starts 4 texels left of the first cell (16-byte aligned origin).
template <int R, int C, int KEYMODE, bool TMA>
__global__ void __launch_bounds__(256) zoom_key_kernel(const __grid_constant__ ZoomArgs A, const __grid_constant__ CUtensorMap tmap) {
  constexpr int N = 2 * R, TAPS = N * N, G = 4;
  constexpr int XO = TMA ? 4 : R;   // staged columns left of the first cell's base texel + 1
  constexpr int SW = TMA ? 72 : (kKTW + N - 1), SH = kKTH + N - 1;
  static_assert(!TMA || (4 + kKTW - 1 + R) <= 72, "TMA box too narrow");
  constexpr int KBUF = (SW * SH + 31) & ~31;
  __shared__ __align__(128) float s_kb[(TMA ? 2 : 1) * KBUF];
  __shared__ __align__(8) uint64_t s_mbar[2];
  const int tid = threadIdx.x;
  const int cw = A.w + 1, ch = A.h + 1;
  const uint64_t tmap_ptr = reinterpret_cast<uint64_t>(&tmap);
  auto tma_issue = [&, tmap_ptr](const TileWalk& tw, int buf) {
    const uint32_t bar = smem_addr(&s_mbar[buf]);
    tma_expect(bar, SW * SH * 4);
    tma_load_3d(smem_addr(s_kb + buf * KBUF), tmap_ptr, tw.tix * kKTW - XO, tw.tiy * kKTH - R, tw.f, bar);
  };
  TileWalk walk(blockIdx.x, gridDim.x, A.tiles_x, A.tiles_y);
  TileWalk ahead = walk;
  if constexpr (TMA) {
    if (tid == 0) {
      mbar_init1(smem_addr(&s_mbar[0]));
      mbar_init1(smem_addr(&s_mbar[1]));
      mbar_init_fence();
    }
    __syncthreads();
    if (tid < 32 && blockIdx.x < A.total_tiles) {
      if (elect_one()) tma_issue(ahead, 0);
    }
  }
  ahead.next();
  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, walk.next(), ahead.next(), ++it) {
    const int f = walk.f;
    const int cx0 = walk.tix * kKTW, cy0 = walk.tiy * kKTH;   // cell tile origin (cell index = base + 1)
    const int64_t src0 = (int64_t)f * A.in_sn;
    float* __restrict__ s_k = s_kb + (TMA ? (it & 1) * KBUF : 0);
    __syncthreads();   // the previous tile's keys are done: its buffer (the one `ahead` lands in) is free
    if constexpr (TMA) {
      if (tid < 32 && tile + gridDim.x < A.total_tiles) {
        fence_proxy_async_smem();
        if (elect_one()) tma_issue(ahead, (it + 1) & 1);
      }
      mbar_wait_parity(smem_addr(&s_mbar[it & 1]), (it >> 1) & 1);
      const bool edge = cx0 - XO < 0 || cy0 - R < 0 || cx0 - XO + SW > A.w || cy0 - R + SH > A.h;
      if (edge) patch_clamp_to_edge(s_k, SW, SW, SH, cx0 - XO, cy0 - R, A.w, A.h, tid, 256, [] { __syncthreads(); });
    } else {
    dispatch_in_fmt(A.io.in_fmt, [&](auto ftag) {
      constexpr int FMT = decltype(ftag)::value;
      for (int i = tid; i < SW * SH; i += 256) {
        const int sy = i / SW, sx = i - sy * SW;
        // window tap (0, 0) of cell (cx, cy) is source texel (cx - 1 - (R-1), cy - 1 - (R-1))
        const int gx = clampi(cx0 - R + sx, 0, A.w - 1), gy = clampi(cy0 - R + sy, 0, A.h - 1);
        const int64_t off = src0 + (int64_t)gy * A.in_sy + gx;
        if constexpr (C == 1) {
          s_k[i] = load_px_t<FMT>(A.in, off, A.io.in_max);
        } else if constexpr (KEYMODE == 2) {
          s_k[i] = key_luma709(load_px_t<FMT>(A.in, off, A.io.in_max), load_px_t<FMT>(A.in, off + A.in_sc, A.io.in_max),
                               load_px_t<FMT>(A.in, off + 2 * A.in_sc, A.io.in_max));
        } else {
          s_k[i] = load_px_t<FMT>(A.in, off, A.io.in_max);
        }
      }
    });
    __syncthreads();
    }
    for (int i = tid; i < kKTW * kKTH; i += 256) {
      const int ly = i / kKTW, lx = i - ly * kKTW;
      const int cx = cx0 + lx, cy = cy0 + ly;
      if (cx >= cw || cy >= ch) continue;
      const float* __restrict__ kb = s_k + ly * SW + lx + (XO - R);
      float ks[TAPS];
#pragma unroll
      for (int t = 0; t < TAPS; ++t) ks[t] = kb[(t % N) * SW + (t / N)];
      const int row = ravu_key2<STENCIL_RAVU, N, G, 3, true>(A.key, [&](int ii, int jj) { return ks[ii * N + jj]; });
      A.kmap[((int64_t)f * ch + cy) * cw + cx] = (unsigned short)row;
    }
  }
}

// ---- phase LUT builder: plut[cp][row][t] = the sampler's blend at the class pair's representative sub-pixel phase ----
struct BuildArgs {
  const float* lut;      // [2592][LWt][4] float32 texels (already rounded to binary16 precision under the rgba16f policy)
  const float* lut_ar;   // or null
  const float* rep_x;    // [ncx] representative sub-pixel phase of every x class
  const float* rep_y;    // [ncy]
  float* plut;           // [ncy * ncx][288][PL]
  int ncx, ncy;
};
template <int R, bool AR>
__global__ void zoom_build_plut_kernel(const BuildArgs A) {
  constexpr int N = 2 * R, TAPS = N * N;
  constexpr int B = (TAPS / 2 + 3) / 4, LWt = B * 9, LHt = 288 * 9;
  constexpr int PL = TAPS * (AR ? 2 : 1);
  const int cp = blockIdx.y, row = blockIdx.x;
  const int cx = cp % A.ncx, cy = cp / A.ncx;
  for (int t = threadIdx.x; t < TAPS; t += blockDim.x) {
    const bool m = t >= TAPS / 2;
    const int k = m ? TAPS - 1 - t : t;
    const int blk = k / 4, comp = k % 4;
    const float px = lutpos9(A.rep_x[cx]), py = lutpos9(A.rep_y[cy]);
    const float spx = __fdiv_rn(m ? __fsub_rn(1.0f, px) : px, (float)B);
    const float spy = __fdiv_rn(m ? __fsub_rn(1.0f, py) : py, 288.0f);
    int xi, yi;
    float fu, fv;
    lin_coord(__fadd_rn(blk_off<B>(blk), spx), LWt, xi, fu);
    lin_coord(__fadd_rn(__fdiv_rn((float)row, 288.0f), spy), LHt, yi, fv);
    const int x0 = clampi(xi, 0, LWt - 1), x1 = clampi(xi + 1, 0, LWt - 1);
    const int y0 = clampi(yi, 0, LHt - 1), y1 = clampi(yi + 1, 0, LHt - 1);
    auto blend = [&](const float* __restrict__ lut) {
      return lerp2(lut[((size_t)y0 * LWt + x0) * 4 + comp], lut[((size_t)y0 * LWt + x1) * 4 + comp],
                   lut[((size_t)y1 * LWt + x0) * 4 + comp], lut[((size_t)y1 * LWt + x1) * 4 + comp], fu, fv);
    };
    float* dst = A.plut + ((size_t)cp * 288 + row) * PL;
    dst[t] = blend(A.lut);
    if constexpr (AR) dst[TAPS + t] = blend(A.lut_ar);
  }
}

// ---- convolution of one class pair's members ------------------------------------------------------------------------
// The gather of 32 different phase-LUT rows per warp is what bounds this kernel (shared-memory wavefronts), so the rows are
// kept compact.  MIX = mixed-precision rows: the four central taps (|w| up to 1.45) stay float32, the others (|w| <= 0.36,
// mostly << 0.1) are stored as binary16 -- an absolute rounding error <= 1.2e-4 on the largest of them, ~1e-4 on the
// output in the worst case (the anti-ringing LUT, whose weights feed 32nd powers, always stays float32).  Row layout:
//   MIX : [c0 c1 c2 c3 : float32][the other TAPS - 4 taps in tap order : binary16][pad]      r3: 80 B, r2: 48 B
//   !MIX: [TAPS float32]([TAPS float32 anti-ringing weights])
// The pitch in 16-byte units is odd, so the rows spread over all bank groups.
template <int R>
__host__ __device__ constexpr bool central_tap(int t) {
  const int N = 2 * R, i = t / N, j = t % N;
  return (i == R - 1 || i == R) && (j == R - 1 || j == R);
}
template <int R, bool AR, bool MIX>
struct PhaseGeom {
  static constexpr int TAPS = 4 * R * R;
  static constexpr int PL = TAPS * (AR ? 2 : 1);                       // floats per row of the global phase LUT
  // anti-ringing rows in shared memory: TAPS float32 main weights + TAPS binary16 anti-ringing weights scaled by 2^13 (they are
  // all >= 0 and <= 1.01, only RATIOS of sums weighted by them are used, so the scale cancels and small weights keep 11 bits)
  static constexpr int QUADS = MIX ? (16 + 2 * (TAPS - 4) + 15) / 16 : (AR ? (TAPS * 4 + TAPS * 2 + 15) / 16 : PL / 4);   // 16-byte units read per row
  static constexpr int PITCH = (QUADS | 1) * 4;                        // floats per row in shared memory
};

__device__ __forceinline__ void store_px_wb(void* __restrict__ p, int64_t off, float v, int fmt, float out_max) {
  // write-back (not streaming) stores: a class pair writes every P-th pixel of a row, the other classes fill in the
  // rest of each 32-byte sector a little later -- the sectors should wait in L2 for that instead of being evicted first
  switch (fmt) {
    case MPVP_FMT_F32: static_cast<float*>(p)[off] = v; break;
    case MPVP_FMT_F16: reinterpret_cast<unsigned short*>(p)[off] = __half_as_ushort(__float2half_rn(v)); break;
    case MPVP_FMT_U8: static_cast<unsigned char*>(p)[off] = (unsigned char)quant_px(v, out_max); break;
    default: static_cast<unsigned short*>(p)[off] = (unsigned short)quant_px(v, out_max); break;
  }
}

// SWT: compile-time pitch of the staged source tile (0 = runtime A.sw): immediate offsets for the window loads.
// TMA (luma float32 planes, SWT > 0): the source tile arrives by cp.async.bulk.tensor; its box starts at the 16-byte
// aligned texel at or left of the tile's first texel (the window offsets absorb the 0..3 texel difference).
// STRIP: member rows per thread (the CTA has kPTW * kPTH / STRIP threads).
#ifndef MPVP_X_ZOOM_STRIP
#define MPVP_X_ZOOM_STRIP 4
#endif
constexpr int kStrip = MPVP_X_ZOOM_STRIP;
#ifndef MPVP_X_ZOOM_AR_MINB
#define MPVP_X_ZOOM_AR_MINB 3   // resident CTAs per SM of the anti-ringing phase kernel (76 registers, no spills; 2: 1.66 ms, 3: 1.47 ms)
#endif
#ifndef MPVP_X_ZOOM_AR_STRIP
#define MPVP_X_ZOOM_AR_STRIP 2   // member rows per strip of the anti-ringing kernel (5 accumulators per pixel: 4 rows need 148 registers)
#endif
constexpr int kPNT2 = kPTW * kPTH / (kStrip * kSPT);
#ifndef MPVP_X_ZOOM_MINB
#define MPVP_X_ZOOM_MINB 3   // resident CTAs per SM of the luma phase kernel (4 = 64 registers: measured)
#endif
template <int R, int C, bool AR, int SWT, bool MIX, bool TMA = false>
__global__ void __launch_bounds__(kPNT2, (C == 1 && !AR) ? (kStrip <= 2 ? 2 : MPVP_X_ZOOM_MINB) : ((C == 1 && MPVP_X_ZOOM_AR_STRIP < 4) ? MPVP_X_ZOOM_AR_MINB : 1)) zoom_phase_kernel(const __grid_constant__ ZoomArgs A, const __grid_constant__ CUtensorMap tmap) {
  static_assert(!TMA || (C == 1 && !AR && SWT > 0), "TMA staging: luma, no anti-ringing power tile, compile-time pitch");
  constexpr int kPNT = kPNT2;
  constexpr int N = 2 * R, TAPS = N * N;
  using PG = PhaseGeom<R, AR, MIX>;
  constexpr int PL = PG::PL, PITCH = PG::PITCH;
  constexpr bool POWT = AR && C == 1;   // anti-ringing powers once per staged source pixel (luma); 3-channel: on the fly
  constexpr int STRIP = (AR && C == 1) ? MPVP_X_ZOOM_AR_STRIP : kStrip;
  constexpr int SPT = kPTH / ((kPNT / kPTW) * STRIP);   // strips per thread and tile
  static_assert(!MIX || !AR, "the anti-ringing weights stay float32");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* s_lut = reinterpret_cast<float*>(smem_raw);            // [288][PITCH]
  float* s_src = s_lut + 288 * PITCH;                           // [C][sh][sw]
  // [sh][sw] (POWT): ((0.1 + l)^32, (1.1 - l)^32) of every staged texel; the 33rd powers are one multiplication away and cost
  // less than the second half of a 16-byte shared-memory read per tap (the tile reads are 2/3 of the kernel's wavefronts)
  float2* s_pow = reinterpret_cast<float2*>(s_src + (((size_t)C * A.sh * A.sw + 3) & ~(size_t)3));
  __shared__ int s_mxo[kPTW], s_mxb[kPTW], s_myo[kPTH], s_myb[kPTH];
  __shared__ __align__(8) uint64_t s_mbar;

  const int tid = threadIdx.x;
  if constexpr (TMA) {
    if (tid == 0) {
      mbar_init1(smem_addr(&s_mbar));
      mbar_init_fence();
    }
  }
  uint32_t tma_phase = 0;
  const int SW = SWT ? SWT : A.sw;
  const int PLANE = SW * A.sh;
  const int cw = A.w + 1, ch = A.h + 1;
  // contiguous run of work items (class-pair major): at most a couple of phase-LUT reloads per CTA
  const long long T = A.total_tiles;
  // member tiles differ in cost (partial segments, the small exact classes of the anti-ringing plans): the host cuts the
  // item list into ranges of equal estimated cost instead of equal length
  const long long w_begin = A.cta_begin ? A.cta_begin[blockIdx.x] : T * blockIdx.x / gridDim.x;
  const long long w_end = A.cta_begin ? A.cta_begin[blockIdx.x + 1] : T * (blockIdx.x + 1) / gridDim.x;
  // Frames are taken in GROUPS whose output stays in L2 (A.fgroup frames): the class pairs of a geometry each write every
  // P-th pixel of a row, i.e. a fraction of every 32-byte sector, and with all frames of a class pair done before the next
  // pair starts the partially written sectors are evicted and re-fetched once per class pair.  Measured (ncu, zoom-r3,
  // 8 x 2160p, 30 MB of input and 265 MB of output): 988 MB of DRAM reads + 763 MB of writes with all frames in one group,
  // 696 + 673 MB with one frame per group; the step time is the same (the kernel is not DRAM-bound), so this only lowers
  // the traffic.  Order: group, class pair, frame, tile.
  int cur_cp = -1;
  int cpi = 0;
  long long cur_group = -1;
  const long long per_group = (long long)A.tiles_per_frame * A.fgroup;
  for (long long item = w_begin; item < w_end; ++item) {
    const long long group = item / per_group;
    const int gi = (int)(item - group * per_group);
    const int gn = min(A.fgroup, A.n - (int)group * A.fgroup);      // frames of this group
    if (group != cur_group) {
      cur_group = group;
      cpi = 0;
    }
    while (cpi + 1 < A.ncp && gi >= A.cps[cpi + 1].tile_start * gn) ++cpi;
    const ClassPair cp = A.cps[cpi];
    const int local = gi - cp.tile_start * gn;
    const int per_frame = cp.tiles_x * cp.tiles_y;
    const int f = (int)group * A.fgroup + local / per_frame;
    const int tl = local % per_frame;
    const int tiy = tl / cp.tiles_x, tix = tl - tiy * cp.tiles_x;
    const int2 sgx = A.xseg[cp.xsoff + tix], sgy = A.yseg[cp.ysoff + tiy];
    const int mx0 = sgx.x, my0 = sgy.x;     // first member (index into xo / xb, yo / yb)
    const int nmx = sgx.y, nmy = sgy.y;     // members of this tile: <= kPTW, kPTH

    __syncthreads();   // previous tile fully consumed (tables, source tile, possibly the LUT slice)
    if (cpi != cur_cp) {
      cur_cp = cpi;
      const float* __restrict__ src = A.plut + (size_t)cpi * 288 * PL;
      if constexpr (MIX) {
        for (int r = tid; r < 288; r += kPNT) {
          const float* __restrict__ g = src + (size_t)r * PL;
          float* d = s_lut + r * PITCH;
          __half* dh = reinterpret_cast<__half*>(d + 4);
          int nc = 0, nh = 0;
#pragma unroll
          for (int t = 0; t < TAPS; ++t) {
            const float wv = __ldg(g + t);
            if (central_tap<R>(t)) d[nc++] = wv;
            else dh[nh++] = __float2half_rn(wv);
          }
        }
      } else if constexpr (AR) {
        for (int i = tid; i < 288 * (TAPS / 4); i += kPNT) {
          const int r = i / (TAPS / 4), q = i - r * (TAPS / 4);
          const float4* __restrict__ g = reinterpret_cast<const float4*>(src + (size_t)r * PL);
          *reinterpret_cast<float4*>(s_lut + r * PITCH + 4 * q) = __ldg(g + q);
          const float4 a = __ldg(g + TAPS / 4 + q);
          const __half2 h0 = __floats2half2_rn(a.x * 8192.0f, a.y * 8192.0f), h1 = __floats2half2_rn(a.z * 8192.0f, a.w * 8192.0f);
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned int*>(&h0);
          pk.y = *reinterpret_cast<const unsigned int*>(&h1);
          *reinterpret_cast<uint2*>(s_lut + r * PITCH + TAPS + 2 * q) = pk;
        }
      } else {
        for (int i = tid; i < 288 * (PL / 4); i += kPNT) {
          const int r = i / (PL / 4), q = i - r * (PL / 4);
          *reinterpret_cast<float4*>(s_lut + r * PITCH + 4 * q) = __ldg(reinterpret_cast<const float4*>(src + (size_t)r * PL) + q);
        }
      }
    }
    if (tid < kPTW) {
      const int m = mx0 + min(tid, nmx - 1);
      s_mxo[tid] = A.xo[m];
      s_mxb[tid] = A.xb[m];
    } else if (tid < kPTW + kPTH) {
      const int j = tid - kPTW;
      const int m = my0 + min(j, nmy - 1);
      s_myo[j] = A.yo[m];
      s_myb[j] = A.yb[m];
    }
    const int xb_true = A.xb[mx0], xb_last = A.xb[mx0 + nmx - 1];
    const int yb_first = A.yb[my0], yb_last = A.yb[my0 + nmy - 1];
    // TMA: the staged tile starts at the aligned texel ax0 <= xb_true - (R-1); in terms of the window arithmetic below
    // that is a tile whose "first base" is ax0 + (R-1)
    const int ax0 = TMA ? ((xb_true - (R - 1)) & ~3) : (xb_true - (R - 1));
    const int xb_first = ax0 + (R - 1);
    const int sx0 = ax0, sy0 = yb_first - (R - 1);
    const int need_w = xb_last - xb_first + N, need_h = yb_last - yb_first + N;   // <= sw, sh by construction of the plan
    const int64_t src0 = (int64_t)f * A.in_sn;
    if constexpr (TMA) {
      if (tid < 32) {
        fence_proxy_async_smem();
        if (elect_one()) {
          const uint32_t bar = smem_addr(&s_mbar);
          tma_expect(bar, SWT * A.sh * 4);
          tma_load_3d(smem_addr(s_src), reinterpret_cast<uint64_t>(&tmap), sx0, sy0, f, bar);
        }
      }
      mbar_wait_parity(smem_addr(&s_mbar), tma_phase);
      tma_phase ^= 1;
      const bool edge = sx0 < 0 || sy0 < 0 || sx0 + SWT > A.w || sy0 + A.sh > A.h;
      if (edge) patch_clamp_to_edge(s_src, SWT, SWT, A.sh, sx0, sy0, A.w, A.h, tid, kPNT, [] { __syncthreads(); });
    } else
    dispatch_in_fmt(A.io.in_fmt, [&](auto ftag) {
      constexpr int FMT = decltype(ftag)::value;
      for (int i = tid; i < need_w * need_h; i += kPNT) {
        const int sy = i / need_w, sx = i - sy * need_w;
        const int gx = clampi(sx0 + sx, 0, A.w - 1), gy = clampi(sy0 + sy, 0, A.h - 1);
        const int64_t off = src0 + (int64_t)gy * A.in_sy + gx;
        const int d = sy * SW + sx;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float v = load_px_t<FMT>(A.in, off + c * A.in_sc, A.io.in_max);
          s_src[c * PLANE + d] = v;
          if constexpr (POWT) {
            const float cc = 0.1f + v, dd = 1.1f - v;
            const float pc = pow32(cc), pd = pow32(dd);
            s_pow[d] = make_float2(pc, pd);
          }
        }
      }
    });
    __syncthreads();

    const int lx = tid & (kPTW - 1);
#pragma unroll 1
    for (int sp = 0; sp < SPT; ++sp) {
    const int ly0 = (tid / kPTW + sp * (kPNT / kPTW)) * STRIP;   // this thread's member rows: ly0 .. ly0 + STRIP - 1
    if (lx >= nmx || ly0 >= nmy) continue;
    const int cnt = min(STRIP, nmy - ly0);
    const int ox = s_mxo[lx], bx = s_mxb[lx];
    const int by0 = s_myb[ly0];

    // the LUT rows of the strip's pixels come from the key map in global memory (L2): all loads issued up front
    int rows[STRIP];
#pragma unroll
    for (int k = 0; k < STRIP; ++k)
      rows[k] = A.kmap[((int64_t)f * ch + (s_myb[ly0 + (k < cnt ? k : 0)] + 1)) * cw + (bx + 1)];

    // weights of LUT row `row` applied to the window whose tap (i, j) is Wn(c, i, j); finishes and stores one output pixel
    auto pixel = [&](int k, int row, auto Wn) {
      const int oy = s_myo[ly0 + k];
      [[maybe_unused]] const int by = s_myb[ly0 + k];
      if (A.bucket) A.bucket[((int64_t)f * A.oh + oy) * A.ow + ox] = row;
      const float4* __restrict__ wr = reinterpret_cast<const float4*>(s_lut + row * PITCH);
      float res[C];
      float hi[C], lo[C], hi2[C], lo2[C];
#pragma unroll
      for (int c = 0; c < C; ++c) res[c] = hi[c] = lo[c] = hi2[c] = lo2[c] = 0.f;
      if constexpr (MIX) {
        const float4 wc = wr[0];
        const float wcv[4] = {wc.x, wc.y, wc.z, wc.w};
        int nc = 0, nh = 0;
        uint4 hq = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int t = 0; t < TAPS; ++t) {
          float wt;
          if (central_tap<R>(t)) {
            wt = wcv[nc++];
          } else {
            if ((nh & 7) == 0) hq = *reinterpret_cast<const uint4*>(&wr[1 + nh / 8]);
            const unsigned int pr = (nh & 7) < 2 ? hq.x : ((nh & 7) < 4 ? hq.y : ((nh & 7) < 6 ? hq.z : hq.w));
            const __half2 h2 = *reinterpret_cast<const __half2*>(&pr);
            wt = (nh & 1) ? __high2float(h2) : __low2float(h2);
            ++nh;
          }
#pragma unroll
          for (int c = 0; c < C; ++c) res[c] = fmaf(Wn(c, t / N, t % N), wt, res[c]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < TAPS / 4; ++q) {
          const float4 w4 = wr[q];
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
          [[maybe_unused]] float av[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (AR) {
            const uint2 ah = *reinterpret_cast<const uint2*>(reinterpret_cast<const float*>(wr) + TAPS + 2 * q);
            const float2 a01 = __half22float2(*reinterpret_cast<const __half2*>(&ah.x));
            const float2 a23 = __half22float2(*reinterpret_cast<const __half2*>(&ah.y));
            av[0] = a01.x; av[1] = a01.y; av[2] = a23.x; av[3] = a23.y;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int t = q * 4 + e;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float sv = Wn(c, t / N, t % N);
              res[c] = fmaf(sv, wv[e], res[c]);
              if constexpr (AR) {
                float4 pw;
                if constexpr (POWT) {
                  const float2 p2 = s_pow[(by - yb_first + t % N) * SW + (bx - xb_first) + t / N];
                  pw = make_float4(p2.x, p2.y, p2.x * (0.1f + sv), p2.y * (1.1f - sv));
                } else {
                  const float cc = 0.1f + sv, dd = 1.1f - sv;
                  const float pc = pow32(cc), pd = pow32(dd);
                  pw = make_float4(pc, pd, pc * cc, pd * dd);
                }
                hi[c] = fmaf(pw.x, av[e], hi[c]);
                lo[c] = fmaf(pw.y, av[e], lo[c]);
                hi2[c] = fmaf(pw.z, av[e], hi2[c]);
                lo2[c] = fmaf(pw.w, av[e], lo2[c]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float r = res[c];
        if constexpr (AR) {
          const float hiv = __fdividef(hi2[c], hi[c]) - 0.1f;
          const float lov = 1.1f - __fdividef(lo2[c], lo[c]);
          const float cl = fminf(fmaxf(r, lov), hiv);
          r = r * (1.0f - A.ar_strength) + cl * A.ar_strength;
        } else {
          r = fminf(fmaxf(r, 0.f), 1.f);
        }
        store_px_wb(A.out, (int64_t)f * A.out_sn + c * A.out_sc + (int64_t)oy * A.out_sy + ox, r, A.io.out_fmt, A.io.out_max);
      }
    };

    // The member rows of a strip usually sit on consecutive base texels (integer ratios): their windows overlap in all but
    // one row, so ONE (N + cnt - 1) x N register window serves the whole strip (warp-uniform test: a warp shares ly0).
    bool consec = (C == 1);
    for (int k = 1; k < cnt; ++k) consec = consec && (s_myb[ly0 + k] == by0 + k);
    if (consec) {
      if constexpr (C == 1) {
        const float* __restrict__ kb = s_src + (by0 - yb_first) * SW + (bx - xb_first);
        float win[N + STRIP - 1][N];
#pragma unroll
        for (int j = 0; j < N + STRIP - 1; ++j)
#pragma unroll
          for (int i = 0; i < N; ++i) win[j][i] = (j < N + cnt - 1) ? kb[j * SW + i] : 0.f;
#pragma unroll
        for (int k = 0; k < STRIP; ++k)
          if (k < cnt) pixel(k, rows[k], [&](int, int i, int j) { return win[k + j][i]; });
      }
    } else {
#pragma unroll
      for (int k = 0; k < STRIP; ++k) {
        if (k >= cnt) break;
        const float* __restrict__ kb = s_src + (s_myb[ly0 + k] - yb_first) * SW + (bx - xb_first);
        pixel(k, rows[k], [&](int c, int i, int j) { return kb[c * PLANE + j * SW + i]; });
      }
    }
    }   // strips of this thread
  }
}

// ---- the plan: phase classes of a geometry, cached on the weight handle -------------------------------------------------
struct AxisPlan {
  std::vector<int> off;      // [ncls + 1] member offsets
  std::vector<int> o, b;     // members: output coordinate, base texel
  std::vector<float> rep;    // representative sub-pixel phase of every class
  std::vector<int> seg_off;  // [ncls + 1] segment offsets
  std::vector<int2> seg;     // member segments (first member, count): the side of a member tile
  int max_need = 0;          // largest staged extent a segment needs: b_last - b_first + N
  bool ok = false;
  bool node_spread = false;  // some class holds SEVERAL phases within fp32 noise of a LUT node (8 * sub ~ integer)
};

// Classes = runs of distinct sub-pixel phases closer than kClusterGap; never joined across the 0 / 1 wrap (at odd integer
// ratios `pos - 0.5` is mathematically an integer for every s-th output coordinate and fp32 rounding puts some of them
// at sub = 0.99999 of base b - 1 and the others at sub = 0.00001 of base b: different windows, hence two classes whose
// members interleave irregularly).  Members of a class are cut into segments of at most `tile` members whose base
// texels span at most tile * q + N source texels (q = the typical member distance), so a sparse class simply yields
// shorter segments.
// exact_nodes (the anti-ringing kernels): a run of phases within fp32 noise of a LUT node (8 * sub ~ integer) is NOT merged.
// There the sampler's blend weight of the neighbouring node is proportional to the distance from the node, the
// anti-ringing weights feed 32nd powers, and one representative cannot stand for phases 0, 1.2e-7, 3e-5 ...: every
// distinct value becomes a class of its own (exact 3x: 640 rows at phase 0 and eleven small classes of 1-20 rows).
AxisPlan build_axis(int O, int I, int tile, int N, bool exact_nodes = false) {
  AxisPlan ap;
  std::vector<float> sub(O);
  std::vector<int> base(O);
  for (int o = 0; o < O; ++o) zoom_pos_host(o, O, I, base[o], sub[o]);
  std::vector<float> u(sub);
  std::sort(u.begin(), u.end());
  u.erase(std::unique(u.begin(), u.end()), u.end());
  std::vector<float> lo, hi;
  for (size_t i = 0; i < u.size(); ++i) {
    if (i == 0 || u[i] - u[i - 1] > kClusterGap) {
      lo.push_back(u[i]);
      hi.push_back(u[i]);
    } else {
      hi.back() = u[i];
    }
  }
  if (exact_nodes) {
    // The blend weight of the FARTHER node is 8 * (distance of the phase to the nearer node): a class whose spread is not
    // small against that distance (on a node the distance is 0) cannot be represented by one phase when the weights feed
    // 32nd powers -- a relative error e of one weight moves the soft min / max by up to e / 4 of the local contrast.
    // Such a class is dissolved into its distinct values (spread <= 4e-3 of the distance keeps the shift below 1e-3).
    std::vector<float> lo2, hi2;
    for (size_t c = 0; c < lo.size(); ++c) {
      const float mid = 0.5f * (lo[c] + hi[c]);
      const float node = rintf(8.0f * mid) / 8.0f;
      const float dist = fminf(fabsf(lo[c] - node), fabsf(hi[c] - node));
      const bool straddles = lo[c] <= node && node <= hi[c];
      if (lo[c] != hi[c] && (straddles || hi[c] - lo[c] > 4e-3f * dist)) {
        for (float v : u)
          if (v >= lo[c] && v <= hi[c]) {
            lo2.push_back(v);
            hi2.push_back(v);
          }
      } else {
        lo2.push_back(lo[c]);
        hi2.push_back(hi[c]);
      }
    }
    lo.swap(lo2);
    hi.swap(hi2);
  }
  const int ncls = (int)lo.size();
  if (ncls > (exact_nodes ? kMaxClassesExact : kMaxClasses)) return ap;
  for (int c = 0; c < ncls; ++c)
    if (hi[c] - lo[c] > kMaxSpread) return ap;
  ap.off.assign(ncls + 1, 0);
  std::vector<int> cls(O);
  for (int o = 0; o < O; ++o) {
    int c = 0;
    while (c + 1 < ncls && sub[o] > hi[c]) ++c;
    cls[o] = c;
    ap.off[c + 1]++;
  }
  for (int c = 0; c < ncls; ++c) ap.off[c + 1] += ap.off[c];
  ap.o.resize(O);
  ap.b.resize(O);
  std::vector<int> fill(ap.off.begin(), ap.off.end() - 1);
  for (int o = 0; o < O; ++o) {
    ap.o[fill[cls[o]]] = o;
    ap.b[fill[cls[o]]++] = base[o];
  }
  // typical base distance of neighbouring members (the median over all classes)
  std::vector<int> steps;
  for (int c = 0; c < ncls; ++c)
    for (int m = ap.off[c] + 1; m < ap.off[c + 1]; ++m) steps.push_back(ap.b[m] - ap.b[m - 1]);
  int q = 1;
  if (!steps.empty()) {
    std::nth_element(steps.begin(), steps.begin() + steps.size() / 2, steps.end());
    q = std::max(1, steps[steps.size() / 2]);
  }
  if (q > kMaxStep) return ap;
  const int limit = tile * q + N;
  ap.seg_off.assign(1, 0);
  for (int c = 0; c < ncls; ++c) {
    // mid-range representative: halves the worst-case distance to a member; exact when the class is a single value
    ap.rep.push_back(lo[c] == hi[c] ? lo[c] : 0.5f * (lo[c] + hi[c]));
    // At a LUT node the sampler's blend weight of the neighbouring node is ~0 and PROPORTIONAL to the distance from the
    // node: the members of such a class have weights 0 ... 2e-4 there, and one representative cannot stand for them.
    // Harmless for the convolution (continuous in the weights) but not for the anti-ringing LUT, whose weights feed
    // 32nd powers (see the file header): the -AR variants take the general path for such geometries.
    const float u8 = 8.0f * ap.rep.back();
    if (lo[c] != hi[c] && fabsf(u8 - rintf(u8)) < 4e-3f) ap.node_spread = true;
    const int m1 = ap.off[c + 1];
    for (int m = ap.off[c]; m < m1;) {
      int cnt = 1;
      while (cnt < tile && m + cnt < m1 && ap.b[m + cnt] - ap.b[m] + N <= limit) ++cnt;
      ap.seg.push_back(make_int2(m, cnt));
      ap.max_need = std::max(ap.max_need, ap.b[m + cnt - 1] - ap.b[m] + N);
      m += cnt;
    }
    ap.seg_off.push_back((int)ap.seg.size());
  }
  ap.ok = true;
  return ap;
}

struct ZoomPlan {
  int h = 0, w = 0, oh = 0, ow = 0;
  const mpvp_weights* lut_ar = nullptr;
  bool usable = false;
  int ncp = 0, total_tiles_per_frame = 0, sw = 0, sh = 0;
  std::vector<ClassPair> cps_host;
  int *xo = nullptr, *xb = nullptr, *yo = nullptr, *yb = nullptr;
  int2 *xseg = nullptr, *yseg = nullptr;
  ClassPair* cps = nullptr;
  float* plut = nullptr;
  // key-map scratch of the pre-pass, owned by the plan and grown on demand: launches allocate nothing in the steady
  // state (a stream-ordered cudaMallocAsync would be re-allocated after every synchronisation of the caller: the
  // default pool trims to zero).  `kmap_ev` orders a launch after the previous user of the buffer, whatever its stream.
  // cost model of the work items (one prefix array per class pair over its tiles of ONE frame) and the cached CTA ranges
  std::vector<std::vector<long long>> cp_prefix;
  struct Ranges { int n, grid, fgroup; long long* dev; };
  std::vector<Ranges> ranges;
  std::mutex kmu;
  unsigned short* kmap = nullptr;
  size_t kmap_cap = 0;
  cudaEvent_t kmap_ev = nullptr;
  ~ZoomPlan() {
    cudaFree(xo); cudaFree(xb); cudaFree(yo); cudaFree(yb); cudaFree(xseg); cudaFree(yseg); cudaFree(cps); cudaFree(plut);
    if (kmap_ev) { cudaEventSynchronize(kmap_ev); cudaEventDestroy(kmap_ev); }
    cudaFree(kmap);
    for (auto& r : ranges) cudaFree(r.dev);
  }
};
struct ZoomPlanCache {
  std::mutex mu;
  std::vector<ZoomPlan*> plans;
};
void free_plan_cache(void* p) {
  ZoomPlanCache* c = static_cast<ZoomPlanCache*>(p);
  for (ZoomPlan* z : c->plans) delete z;
  delete c;
}

template <class T>
cudaError_t upload(T*& dst, const std::vector<T>& v) {
  cudaError_t e = cudaMalloc(&dst, std::max<size_t>(v.size(), 1) * sizeof(T));
  if (e == cudaSuccess && !v.empty()) e = cudaMemcpy(dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return e;
}

// Build (or find) the plan of a geometry.  Returns null with no error set when the geometry does not qualify for the phase
// path; a CUDA failure also falls back to the general path (the plan is marked unusable).
template <int R, bool AR>
ZoomPlan* get_plan(const mpvp_weights* lut, const mpvp_weights* lut_ar, int h, int w, int oh, int ow, int C, cudaStream_t stream) {
  mpvp_weights* L = const_cast<mpvp_weights*>(lut);
  static std::mutex create_mu;
  {
    std::lock_guard<std::mutex> lk(create_mu);
    if (!L->zoom_plans) {
      L->zoom_plans = new ZoomPlanCache();
      L->zoom_plans_free = free_plan_cache;
    }
  }
  ZoomPlanCache* cache = static_cast<ZoomPlanCache*>(L->zoom_plans);
  std::lock_guard<std::mutex> lk(cache->mu);
  for (ZoomPlan* z : cache->plans)
    if (z->h == h && z->w == w && z->oh == oh && z->ow == ow && z->lut_ar == lut_ar) return z->usable ? z : nullptr;
  ZoomPlan* z = new ZoomPlan();
  z->h = h; z->w = w; z->oh = oh; z->ow = ow; z->lut_ar = lut_ar;
  cache->plans.push_back(z);
  constexpr int N = 2 * R;
  const AxisPlan ax = build_axis(ow, w, kPTW, N, AR), ay = build_axis(oh, h, kPTH, N, AR);
  if (env_flag("MPVP_DEBUG_ZOOM", false))
    fprintf(stderr, "[mpvp] zoom plan %dx%d -> %dx%d: x %s (%zu classes, %zu segments, need %d), y %s (%zu classes, %zu segments, need %d)\n",
            w, h, ow, oh, ax.ok ? "ok" : "no", ax.rep.size(), ax.seg.size(), ax.max_need, ay.ok ? "ok" : "no", ay.rep.size(),
            ay.seg.size(), ay.max_need);
  if (!ax.ok || !ay.ok) return nullptr;
#ifndef MPVP_X_ZOOM_AR_SPREAD_OK
#define MPVP_X_ZOOM_AR_SPREAD_OK 0   // timing experiments only: wrong results on the rows / columns whose phase scatters around a node
#endif
  if (AR && (ax.node_spread || ay.node_spread) && !MPVP_X_ZOOM_AR_SPREAD_OK) return nullptr;
  const int ncx = (int)ax.rep.size(), ncy = (int)ay.rep.size();
  z->ncp = ncx * ncy;
  z->sw = ax.max_need | 1;   // odd pitch: member windows one or more texels apart spread over the banks
  z->sh = ay.max_need;
  constexpr int PL = PhaseGeom<R, AR, false>::PL;
  const size_t smem = sizeof(float) * (288 * PhaseGeom<R, AR, false>::PITCH + (((size_t)C * z->sw * z->sh + 3) & ~(size_t)3)) +
                      ((AR && C == 1) ? sizeof(float2) * (size_t)z->sw * z->sh : 0);
  if (smem > 200 * 1024) return nullptr;
  int start = 0;
  for (int cy = 0; cy < ncy; ++cy)
    for (int cx = 0; cx < ncx; ++cx) {
      ClassPair cp{};
      cp.xsoff = ax.seg_off[cx]; cp.tiles_x = ax.seg_off[cx + 1] - ax.seg_off[cx];
      cp.ysoff = ay.seg_off[cy]; cp.tiles_y = ay.seg_off[cy + 1] - ay.seg_off[cy];
      cp.tile_start = start;   // per frame; scaled by n at launch
      start += cp.tiles_x * cp.tiles_y;
      z->cps_host.push_back(cp);
    }
  z->total_tiles_per_frame = start;
  {
    // estimated cost of a member tile: a fixed part (staging, barriers, table loads, in units of one pixel) + its pixels.
    // Measured (config 4, 8 frames; equal-length ranges 0.874 / 1.660 ms for r3 / ar-r2): fixed = 128: 1.32 / 2.14, 384:
    // 1.08 / 1.86, 768: 0.92 / 1.60, 1536: 0.851 / 1.471, 2048: 0.848 / 1.498, 4096: 0.855 / 1.549 -- the set-up of a
    // tile costs about as much as the pixels of a full 64 x 32 tile.
    const long long fixed = env_int("MPVP_ZOOM_COST_FIXED", AR ? 1536 : 2048);
    for (const ClassPair& cp : z->cps_host) {
      std::vector<long long> pre(1, 0);
      for (int ty = 0; ty < cp.tiles_y; ++ty)
        for (int tx = 0; tx < cp.tiles_x; ++tx)
          pre.push_back(pre.back() + fixed + (long long)ax.seg[cp.xsoff + tx].y * ay.seg[cp.ysoff + ty].y);
      z->cp_prefix.push_back(std::move(pre));
    }
  }
  float *rep_x = nullptr, *rep_y = nullptr;
  cudaError_t e = upload(z->xo, ax.o);
  if (e == cudaSuccess) e = upload(z->xb, ax.b);
  if (e == cudaSuccess) e = upload(z->yo, ay.o);
  if (e == cudaSuccess) e = upload(z->yb, ay.b);
  if (e == cudaSuccess) e = upload(z->xseg, ax.seg);
  if (e == cudaSuccess) e = upload(z->yseg, ay.seg);
  if (e == cudaSuccess) e = upload(z->cps, z->cps_host);
  if (e == cudaSuccess) e = upload(rep_x, ax.rep);
  if (e == cudaSuccess) e = upload(rep_y, ay.rep);
  if (e == cudaSuccess) e = cudaMalloc(&z->plut, sizeof(float) * (size_t)z->ncp * 288 * PL);
  if (e == cudaSuccess) {
    BuildArgs b{};
    b.lut = lut->lut; b.lut_ar = lut_ar ? lut_ar->lut : nullptr;
    b.rep_x = rep_x; b.rep_y = rep_y; b.plut = z->plut; b.ncx = ncx; b.ncy = ncy;
    zoom_build_plut_kernel<R, AR><<<dim3(288, z->ncp), 64, 0, stream>>>(b);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);   // rep_x / rep_y are freed below; once per geometry
  }
  cudaFree(rep_x);
  cudaFree(rep_y);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  z->usable = true;
  return z;
}

// First work item of every CTA for an n-frame launch on `grid` CTAs: ranges of equal estimated cost.  The item order is
// (group of fgroup frames, class pair, frame, tile); the cost of an item depends on (class pair, tile) only, so the item at a
// given cumulative cost is found by division and one binary search.  Cached per (n, grid, fgroup) on the plan (caller holds
// z->kmu); null if the device allocation fails (the kernel then falls back to ranges of equal length).
const long long* cta_ranges(ZoomPlan* z, int n, int grid, int fgroup) {
  for (const auto& r : z->ranges)
    if (r.n == n && r.grid == grid && r.fgroup == fgroup) return r.dev;
  const int ncp = (int)z->cps_host.size();
  long long frame_cost = 0;
  for (const auto& pre : z->cp_prefix) frame_cost += pre.back();
  const long long tpf = z->total_tiles_per_frame;
  const long long total_cost = frame_cost * n, total_items = tpf * n;
  const int ngroups = (n + fgroup - 1) / fgroup;
  std::vector<long long> b((size_t)grid + 1);
  for (int k = 0; k <= grid; ++k) {
    if (k == grid) { b[k] = total_items; break; }
    long long t = total_cost / grid * k + total_cost % grid * k / grid;      // = floor(total_cost * k / grid) without overflow
    int g = (int)std::min<long long>(t / (frame_cost * fgroup), ngroups - 1);
    t -= (long long)g * frame_cost * fgroup;
    const int gn = std::min(fgroup, n - g * fgroup);
    int cp = 0;
    while (cp + 1 < ncp && t >= z->cp_prefix[cp].back() * gn) { t -= z->cp_prefix[cp].back() * gn; ++cp; }
    const auto& pre = z->cp_prefix[cp];
    const long long f = std::min<long long>(t / pre.back(), gn - 1);
    t -= f * pre.back();
    const long long tl = std::min<long long>((long long)(std::upper_bound(pre.begin(), pre.end(), t) - pre.begin()) - 1, (long long)pre.size() - 2);
    b[k] = (long long)g * fgroup * tpf + (long long)z->cps_host[cp].tile_start * gn + f * ((long long)pre.size() - 1) + std::max<long long>(tl, 0);
  }
  for (int k = 1; k <= grid; ++k) b[k] = std::max(b[k], b[k - 1]);
  long long* dev = nullptr;
  if (cudaMalloc(&dev, b.size() * sizeof(long long)) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  if (cudaMemcpy(dev, b.data(), b.size() * sizeof(long long), cudaMemcpyHostToDevice) != cudaSuccess) {
    (void)cudaGetLastError();
    cudaFree(dev);
    return nullptr;
  }
  if (z->ranges.size() >= 16) { cudaFree(z->ranges.front().dev); z->ranges.erase(z->ranges.begin()); }   // (launches are ordered by kmap_ev)
  z->ranges.push_back({n, grid, fgroup, dev});
  return dev;
}

template <int R, int C, int KEYMODE, bool AR>
int launch_zoom_phase(ZoomArgs a, ZoomPlan* z, int device, cudaStream_t stream) {
  // work items are class-pair major and cover all frames of a class pair before the next one starts
  const long long total = (long long)z->total_tiles_per_frame * a.n;
  MPVP_REQUIRE(total < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", total);
  const size_t kbytes = sizeof(unsigned short) * (size_t)a.n * (a.h + 1) * (a.w + 1);
  std::lock_guard<std::mutex> klk(z->kmu);
  if (!z->kmap_ev) MPVP_CUDA_OK(cudaEventCreateWithFlags(&z->kmap_ev, cudaEventDisableTiming));
  if (z->kmap_cap < kbytes) {   // first use of this geometry, or a larger batch than before
    if (z->kmap) {
      MPVP_CUDA_OK(cudaEventSynchronize(z->kmap_ev));
      cudaFree(z->kmap);
      z->kmap = nullptr;
      z->kmap_cap = 0;
    }
    MPVP_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&z->kmap), kbytes));
    z->kmap_cap = kbytes;
  } else {
    MPVP_CUDA_OK(cudaStreamWaitEvent(stream, z->kmap_ev, 0));   // the previous launch that used the buffer (any stream)
  }
  cudaError_t e = cudaSuccess;
  a.kmap = z->kmap;
  a.xo = z->xo; a.xb = z->xb; a.yo = z->yo; a.yb = z->yb; a.xseg = z->xseg; a.yseg = z->yseg; a.cps = z->cps; a.ncp = z->ncp; a.plut = z->plut;
  a.tiles_per_frame = z->total_tiles_per_frame;
  {
    // frames per group: their output (every class pair writes a fraction of each of its sectors) should stay in L2
    const size_t out_frame = (size_t)a.oh * a.ow * C * fmt_bytes(a.io.out_fmt);
    long long g = (long long)((size_t)48 << 20) / (long long)(out_frame ? out_frame : 1);
    if (const char* e = getenv("MPVP_ZOOM_FGROUP")) g = atoll(e);     // A/B switch (0 = all frames in one group)
    if (g < 1 || g > a.n) g = (g == 0 && getenv("MPVP_ZOOM_FGROUP")) ? a.n : (g < 1 ? 1 : a.n);
    a.fgroup = (int)g;
  }
  a.sw = z->sw; a.sh = z->sh;
  int rc = MPVP_OK;
  {
    ZoomArgs k = a;
    k.tiles_x = (a.w + 1 + kKTW - 1) / kKTW;
    k.tiles_y = (a.h + 1 + kKTH - 1) / kKTH;
    k.total_tiles = (long long)k.tiles_x * k.tiles_y * a.n;
    alignas(64) CUtensorMap ktm;
    memset(&ktm, 0, sizeof(ktm));
    bool ktma = false;
    if constexpr (C == 1) ktma = a.io.in_fmt == MPVP_FMT_F32 && make_plane_tmap(&ktm, a.in, 4, a.w, a.h, a.n, a.in_sy, a.in_sn, 72, kKTH + 2 * R - 1);
    auto kern = ktma ? zoom_key_kernel<R, C, KEYMODE, C == 1> : zoom_key_kernel<R, C, KEYMODE, false>;
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0);
    long long grid = (long long)sm_count(device) * (per_sm < 1 ? 1 : per_sm);
    if (grid > k.total_tiles) grid = k.total_tiles;
    grid = cap_grid(grid);
    kern<<<(unsigned)grid, 256, 0, stream>>>(k, ktm);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  {
    a.total_tiles = total;
    // MPVP_ZOOM_MIX=0: all phase-LUT weights stay float32 in shared memory (A/B switch for the binary16 outer taps)
    const bool mix = !AR && env_flag("MPVP_ZOOM_MIX", true);
    // compile-time tile pitch when neighbouring members sit one texel apart: tile + window + up to 3 texels of slack for
    // the 16-byte aligned TMA origin, a multiple of 4 texels (TMA box rows are multiples of 16 bytes)
    constexpr int kSWT = (kPTW + 2 * R + 3 + 3) & ~3;
    const bool fixed_pitch = a.sw + 3 <= kSWT;
    if (fixed_pitch) a.sw = kSWT;
    alignas(64) CUtensorMap ptm;
    memset(&ptm, 0, sizeof(ptm));
    bool ptma = false;
    if constexpr (C == 1 && !AR)
      ptma = fixed_pitch && a.io.in_fmt == MPVP_FMT_F32 && make_plane_tmap(&ptm, a.in, 4, a.w, a.h, a.n, a.in_sy, a.in_sn, kSWT, a.sh);
    auto launch = [&](auto kern, int pitch) {
      const size_t smem = sizeof(float) * (288 * (size_t)pitch + (((size_t)C * a.sw * a.sh + 3) & ~(size_t)3)) +
                          ((AR && C == 1) ? sizeof(float2) * (size_t)a.sw * a.sh : 0);
      cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      int per_sm = 0;
      if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPNT2, smem);
      if (e2 != cudaSuccess || per_sm < 1) {
        set_error("ravu-zoom phase kernel does not fit on an SM (smem %zu B): %s", smem, cudaGetErrorString(e2));
        rc = MPVP_E_UNSUPPORTED;
        return;
      }
      long long grid = (long long)sm_count(device) * per_sm;
      if (grid > total) grid = total;
      grid = cap_grid(grid);
      a.cta_begin = env_flag("MPVP_ZOOM_BALANCE", true) ? cta_ranges(z, a.n, (int)grid, a.fgroup) : nullptr;
      kern<<<(unsigned)grid, kPNT2, smem, stream>>>(a, ptm);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    };
    if constexpr (AR) {
      if (fixed_pitch) launch(zoom_phase_kernel<R, C, AR, kSWT, false>, PhaseGeom<R, AR, false>::PITCH);
      else launch(zoom_phase_kernel<R, C, AR, 0, false>, PhaseGeom<R, AR, false>::PITCH);
    } else {
      if constexpr (C == 1) {
        if (ptma && mix) launch(zoom_phase_kernel<R, C, AR, kSWT, true, true>, PhaseGeom<R, AR, true>::PITCH);
        else if (ptma) launch(zoom_phase_kernel<R, C, AR, kSWT, false, true>, PhaseGeom<R, AR, false>::PITCH);
      }
      if (!ptma) {
        if (mix && fixed_pitch) launch(zoom_phase_kernel<R, C, AR, kSWT, true>, PhaseGeom<R, AR, true>::PITCH);
        else if (mix) launch(zoom_phase_kernel<R, C, AR, 0, true>, PhaseGeom<R, AR, true>::PITCH);
        else if (fixed_pitch) launch(zoom_phase_kernel<R, C, AR, kSWT, false>, PhaseGeom<R, AR, false>::PITCH);
        else launch(zoom_phase_kernel<R, C, AR, 0, false>, PhaseGeom<R, AR, false>::PITCH);
      }
    }
  }
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaEventRecord(z->kmap_ev, stream);
  if (rc != MPVP_OK) return rc;
  MPVP_CUDA_OK(e);
  return MPVP_OK;
}

template <int R, int C, int KEYMODE, bool AR>
int launch_zoom(const ZoomArgs& a, const mpvp_weights* lut, const mpvp_weights* lut_ar, int device, cudaStream_t stream, bool half_lut) {
  // MPVP_ZOOM_PHASE=0 forces the general per-pixel path (A/B switch and cross-check).  Three-channel planes take the
  // general path by default (MPVP_ZOOM_PHASE=3 forces the phase path): without the shared register window of the luma
  // strips the phase kernel is slower there (zoom-r2-yuv 720p->2160p x4: 1.02 vs 0.82 ms).
  const char* pe = getenv("MPVP_ZOOM_PHASE");
  const bool phase_ok = C == 1 ? env_flag("MPVP_ZOOM_PHASE", true) : (pe && pe[0] == '3');
  if (phase_ok && !env_flag("MPVP_ZOOM_TEX", false)) {
    if (ZoomPlan* z = get_plan<R, AR>(lut, lut_ar, a.h, a.w, a.oh, a.ow, C, stream))
      return launch_zoom_phase<R, C, KEYMODE, AR>(a, z, device, stream);
  }
  return launch_zoom_general<R, C, KEYMODE, AR>(a, device, stream, half_lut);
}

}  // namespace
}  // namespace mpvp

using namespace mpvp;

extern "C" int mpvp_ravu_zoom_launch(const mpvp_weights* lut, const mpvp_weights* lut_ar, const mpvp_key_params* key,
                                     int radius, int key_mode, float ar_strength, const float* in, float* out, int n,
                                     int h, int w, int out_h, int out_w, int64_t in_stride_n, int64_t in_stride_c,
                                     int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_c,
                                     int64_t out_stride_y, int32_t* bucket_out, void* stream) {
  return mpvp_ravu_zoom_launch_io(lut, lut_ar, key, radius, key_mode, ar_strength, in, out, n, h, w, out_h, out_w,
                                  in_stride_n, in_stride_c, in_stride_y, out_stride_n, out_stride_c, out_stride_y,
                                  bucket_out, nullptr, stream);
}

extern "C" int mpvp_ravu_zoom_launch_io(const mpvp_weights* lut, const mpvp_weights* lut_ar, const mpvp_key_params* key,
                                        int radius, int key_mode, float ar_strength, const void* in, void* out, int n,
                                        int h, int w, int out_h, int out_w, int64_t in_stride_n, int64_t in_stride_c,
                                        int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_c,
                                        int64_t out_stride_y, int32_t* bucket_out, const mpvp_io* io, void* stream) {
  IoFmt iof;
  if (int rc0 = parse_io(io, iof)) return rc0;
  MPVP_REQUIRE(lut && lut->kind == 0 && lut->lut, "lut handle is null or not a LUT");
  MPVP_REQUIRE(!lut_ar || (lut_ar->kind == 0 && lut_ar->lut && lut_ar->device == lut->device), "bad lut_ar handle");
  MPVP_REQUIRE(key && in && out, "null argument");
  MPVP_REQUIRE(radius == 2 || radius == 3, "ravu-zoom radius %d not in {2,3}", radius);
  MPVP_REQUIRE(key_mode >= 0 && key_mode <= 2, "key_mode %d", key_mode);
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1 && out_h >= 1 && out_w >= 1, "bad geometry");
  MPVP_REQUIRE(out_h >= h && out_w >= w, "ravu-zoom only upscales (%dx%d -> %dx%d)", w, h, out_w, out_h);
  const int B = (2 * radius * radius + 3) / 4;
  MPVP_REQUIRE(lut->lut_w == B * 9 && lut->lut_h == 2592, "LUT is %dx%d, expected %dx2592", lut->lut_w, lut->lut_h,
               B * 9);
  MPVP_REQUIRE(!lut_ar || (lut_ar->lut_w == lut->lut_w && lut_ar->lut_h == lut->lut_h), "lut_ar geometry differs");
  MPVP_REQUIRE(!lut_ar || radius == 2, "RAVU-Zoom-AR exists for radius 2 only (the reference ships no r3 anti-ringing LUT)");
  MPVP_REQUIRE(key->n_gauss == 16 && key->n_strength == 4 && key->n_strength_thr == 3,
               "key params do not describe a RAVU-Zoom hook");
  if (int rck = check_fast_key(key)) return rck;
  if (n == 0) return MPVP_OK;
  DeviceGuard guard(lut->device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", lut->device);
  ZoomArgs a{};
  a.io = iof;
  a.in = in; a.out = out; a.bucket = bucket_out;
  // binary16 texels are exact whenever the LUT was created with round_to_fp16 (the rgba16f policy)
  const bool half_lut = lut->lut_half && (!lut_ar || lut_ar->lut_half);
  a.lut = half_lut ? lut->lut_half : static_cast<const void*>(lut->lut);
  a.lut_ar = lut_ar ? (half_lut ? lut_ar->lut_half : static_cast<const void*>(lut_ar->lut)) : nullptr;
  a.tex = lut->lut_tex;
  a.tex_ar = lut_ar ? lut_ar->lut_tex : 0;
  a.n = n; a.h = h; a.w = w; a.oh = out_h; a.ow = out_w;
  a.in_sn = in_stride_n; a.in_sc = in_stride_c; a.in_sy = in_stride_y;
  a.out_sn = out_stride_n; a.out_sc = out_stride_c; a.out_sy = out_stride_y;
  a.ar_strength = ar_strength;
  a.key = *key;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dev = lut->device;
  const int ar = lut_ar ? 1 : 0;
  switch ((radius * 3 + key_mode) * 2 + ar) {
    case 12: return launch_zoom<2, 1, 0, false>(a, lut, lut_ar, dev, st, half_lut);
    case 13: return launch_zoom<2, 1, 0, true>(a, lut, lut_ar, dev, st, half_lut);
    case 14: return launch_zoom<2, 3, 1, false>(a, lut, lut_ar, dev, st, half_lut);
    case 15: return launch_zoom<2, 3, 1, true>(a, lut, lut_ar, dev, st, half_lut);
    case 16: return launch_zoom<2, 3, 2, false>(a, lut, lut_ar, dev, st, half_lut);
    case 17: return launch_zoom<2, 3, 2, true>(a, lut, lut_ar, dev, st, half_lut);
    case 18: return launch_zoom<3, 1, 0, false>(a, lut, lut_ar, dev, st, half_lut);
    case 20: return launch_zoom<3, 3, 1, false>(a, lut, lut_ar, dev, st, half_lut);
    case 22: return launch_zoom<3, 3, 2, false>(a, lut, lut_ar, dev, st, half_lut);
  }
  return MPVP_E_INVALID;
}
