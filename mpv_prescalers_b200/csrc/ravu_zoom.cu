// RAVU-Zoom(-AR): arbitrary-ratio upscale, one thread per OUTPUT pixel.
//
//   RAVU-Zoom      ravu-zoom-r2.hook:15-134, ravu-zoom-r3.hook:15-180
//   RAVU-Zoom-AR   ravu-zoom-ar-r2.hook:15-208  (3-channel mat4x3 form: ravu-zoom-ar-r2-rgb.hook:153-223)
//
// Position arithmetic follows SURVEY.md App. D.6 exactly (pos = ((o + 0.5) / O) * I in fp32, then
// subpix = fract(pos - 0.5)): at integer ratios a 1-ulp difference moves the 4x4 / 6x6 window.
// A CTA owns a 32x8 tile of output pixels; the source rectangle it taps (at most tile + 2r + 2,
// because the hook only runs when upscaling) is staged in shared memory with clamp-to-edge.  The LUT
// ([288*9][B*9] float4, FILTER LINEAR) stays in global memory / L2 and is fetched with an explicit
// fp32 bilinear blend of four texels, at the same texel coordinates the GL sampler would use.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace mpvp {
namespace {

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
};

#ifndef MPVP_X_ZOOM_MANUAL
#define MPVP_X_ZOOM_MANUAL 0
#endif
constexpr int kTOW = 32, kTOH = 32, kNT = 256;  // output tile; each thread owns one column and kTOH/8 rows

// Canonical position arithmetic: base texel index and sub-pixel phase of output coordinate o.
__device__ __forceinline__ void zoom_pos(int o, int O, int I, int& base, float& sub) {
  const float pos = __fmul_rn(__fdiv_rn(__fadd_rn((float)o, 0.5f), (float)O), (float)I);
  const float t = __fsub_rn(pos, 0.5f);
  sub = __fsub_rn(t, floorf(t));
  base = (int)floorf(__fsub_rn(pos, sub));
}

// LUTPOS(x, 9) = mix(0.5/9, 1 - 0.5/9, x) = a*(1-x) + b*x
__device__ __forceinline__ float lutpos9(float x) {
  const float a = __fdiv_rn(0.5f, 9.0f);
  const float b = __fsub_rn(1.0f, a);
  return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, x)), __fmul_rn(b, x));
}

// Per output column (or row): everything that depends on one coordinate only.  The GL LINEAR fetch at
// normalised coordinate c of an axis with `size` texels reads texels floor(c*size - 0.5) and +1 with weight
// frac(c*size - 0.5); inside a 9-texel LUT block that is i0 = floor(8 s) and f = frac(8 s) for the direct
// half of the taps (sub-pixel phase s) and the same at 1 - s for the mirrored half (ravu-zoom-r2.hook:24-31).
struct AxisEntry {
  int base;        // source texel index of window tap 0 minus the tile origin (filled by the caller)
  int i0, i0m;     // first LUT texel inside the 9-texel block: direct / mirrored
  float f, fm;     // blend weight of texel i0+1: direct / mirrored
  float u, um;     // the same as continuous texel coordinates inside the block (texel centres at k + 0.5)
};

__device__ __forceinline__ AxisEntry axis_entry(int o, int O, int I, int groups) {
  // groups = number of 9-texel blocks along this LUT axis (B for x, 288 rows for y)
  AxisEntry e;
  float sub;
  zoom_pos(o, O, I, e.base, sub);
  const float p = lutpos9(sub), ip = __fsub_rn(1.0f, p);
  // the shader divides by `groups` to normalise and the sampler multiplies by groups*9 again
  const float u = __fsub_rn(__fmul_rn(__fdiv_rn(p, (float)groups), (float)(groups * 9)), 0.5f);
  const float um = __fsub_rn(__fmul_rn(__fdiv_rn(ip, (float)groups), (float)(groups * 9)), 0.5f);
  const float u0 = floorf(u), um0 = floorf(um);
  e.i0 = (int)u0; e.f = __fsub_rn(u, u0);
  e.i0m = (int)um0; e.fm = __fsub_rn(um, um0);
  e.u = u + 0.5f; e.um = um + 0.5f;
  return e;
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

// TEXF: the LUT fetch is ONE texture instruction (hardware bilinear blend of the four binary16 texels, exactly what
// the reference's FILTER LINEAR sampler does, 8-bit blend weights: |error| <= 9e-5) instead of
// four gathered loads and twelve FMAs.
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

  __shared__ float s_src[NP * PLANE];
  __shared__ int s_key[CW * CH];
  __shared__ AxisEntry s_ax[kTOW], s_ay[kTOH];

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
      AxisEntry e = axis_entry(min(ox0 + tid, A.ow - 1), A.ow, A.w, B);
      e.base -= bx_first;
      s_ax[tid] = e;
    } else if (tid < kTOW + kTOH) {
      AxisEntry e = axis_entry(min(oy0 + tid - kTOW, A.oh - 1), A.oh, A.h, 288);
      e.base -= by_first;
      s_ay[tid - kTOW] = e;
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
          s_src[d] = load_px_t<FMT>(A.in, off, A.io.in_max);
        } else {
          const float c0 = load_px_t<FMT>(A.in, off, A.io.in_max);
          const float c1 = load_px_t<FMT>(A.in, off + A.in_sc, A.io.in_max);
          const float c2 = load_px_t<FMT>(A.in, off + 2 * A.in_sc, A.io.in_max);
          s_src[d] = (KEYMODE == 2) ? __fadd_rn(__fadd_rn(__fmul_rn(c0, 0.2126f), __fmul_rn(c1, 0.7152f)), __fmul_rn(c2, 0.0722f)) : c0;
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

    // A warp covers an 8x4 patch of output pixels, not a 32x1 row segment: the LUT fetch is bound by the number of
    // distinct texel neighbourhoods per texture instruction, and a compact patch spans 3-4 times fewer source cells
    // (hence LUT row groups) than a row segment does.
#pragma unroll 1
    for (int rr = 0; rr < RPT; ++rr) {
      const int patch = rr * (kNT / 32) + (tid >> 5);           // 4 patches across, 8 down
      const int lx = (patch & 3) * 8 + (tid & 7), ly = (patch >> 2) * 4 + ((tid >> 3) & 3);
      const int ox = ox0 + lx, oy = oy0 + ly;
      if (ox >= A.ow || oy >= A.oh) continue;
      const AxisEntry ex = s_ax[lx];
      const AxisEntry ey = s_ay[ly];
      const int row = s_key[ey.base * CW + ex.base];
      if (A.bucket) A.bucket[((int64_t)f * A.oh + oy) * A.ow + ox] = row;
      const float* __restrict__ kb = s_src + ey.base * SWt + ex.base;  // window tap (0,0)

      float res[C];
      float hi[C], lo[C], hi2[C], lo2[C];
#pragma unroll
      for (int c = 0; c < C; ++c) res[c] = hi[c] = lo[c] = hi2[c] = lo2[c] = 0.f;

#pragma unroll
      for (int m = 0; m < 2; ++m) {
        // LUT rows: row*9 + i0 (+1), clamp-to-edge on the whole texture like the GL sampler
        const int yi = row * 9 + (m ? ey.i0m : ey.i0);
        const int y0 = clampi(yi, 0, LHt - 1), y1 = clampi(yi + 1, 0, LHt - 1);
        const float fv = m ? ey.fm : ey.f, fu = m ? ex.fm : ex.f;
        const float w00 = (1.0f - fu) * (1.0f - fv), w10 = fu * (1.0f - fv), w01 = (1.0f - fu) * fv, w11 = fu * fv;
        const int xi = m ? ex.i0m : ex.i0;
#pragma unroll
        for (int blk = 0; blk < B; ++blk) {
          const int x0 = clampi(blk * 9 + xi, 0, LWt - 1), x1 = clampi(blk * 9 + xi + 1, 0, LWt - 1);
          auto fetch = [&](const void* __restrict__ lut) {
            const float4 t00 = lut_texel<LUTH>(lut, y0 * LWt + x0), t10 = lut_texel<LUTH>(lut, y0 * LWt + x1);
            const float4 t01 = lut_texel<LUTH>(lut, y1 * LWt + x0), t11 = lut_texel<LUTH>(lut, y1 * LWt + x1);
            float4 r;
            r.x = t00.x * w00 + t10.x * w10 + t01.x * w01 + t11.x * w11;
            r.y = t00.y * w00 + t10.y * w10 + t01.y * w01 + t11.y * w11;
            r.z = t00.z * w00 + t10.z * w10 + t01.z * w01 + t11.z * w11;
            r.w = t00.w * w00 + t10.w * w10 + t01.w * w01 + t11.w * w11;
            return r;
          };
          float4 w4;
          float av[4] = {0.f, 0.f, 0.f, 0.f};
          // MPVP_X_ZOOM_MANUAL of the B blocks are fetched by explicit loads + fp32 blend even on the texture path, to
          // take load off the texture data pipe (the binding unit)
          if (TEXF && blk >= MPVP_X_ZOOM_MANUAL) {
            const float X = (float)(blk * 9) + (m ? ex.um : ex.u), Y = (float)(row * 9) + (m ? ey.um : ey.u);
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
                  const float cc = 0.1f + s, dd = 1.1f - s;
                  const float pc = pow32(cc), pd = pow32(dd);
                  hi[c] = fmaf(pc, av[e], hi[c]);
                  lo[c] = fmaf(pd, av[e], lo[c]);
                  hi2[c] = fmaf(pc * cc, av[e], hi2[c]);
                  lo2[c] = fmaf(pd * dd, av[e], lo2[c]);
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
  if (grid < 1) return MPVP_OK;
  kern<<<(unsigned)grid, kNT, 0, stream>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}

template <int R, int C, int KEYMODE, bool AR>
int launch_zoom(const ZoomArgs& a, int device, cudaStream_t stream, bool half_lut) {
  // MPVP_ZOOM_TEX=0: explicit fp32 blend of four loaded texels instead of the texture unit (A/B switch)
  static const bool tex_ok = [] {
    const char* e = getenv("MPVP_ZOOM_TEX");
    return !(e && e[0] == '0');
  }();
  if (half_lut && tex_ok && a.tex && (!AR || a.tex_ar)) return launch_zoom_impl<R, C, KEYMODE, AR, true, true>(a, device, stream);
  if (half_lut) return launch_zoom_impl<R, C, KEYMODE, AR, true, false>(a, device, stream);
  return launch_zoom_impl<R, C, KEYMODE, AR, false, false>(a, device, stream);
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
  MPVP_REQUIRE(key->n_gauss == 16 && key->n_strength == 4 && key->n_strength_thr == 3,
               "key params do not describe a RAVU-Zoom hook");
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
    case 12: return launch_zoom<2, 1, 0, false>(a, dev, st, half_lut);
    case 13: return launch_zoom<2, 1, 0, true>(a, dev, st, half_lut);
    case 14: return launch_zoom<2, 3, 1, false>(a, dev, st, half_lut);
    case 15: return launch_zoom<2, 3, 1, true>(a, dev, st, half_lut);
    case 16: return launch_zoom<2, 3, 2, false>(a, dev, st, half_lut);
    case 17: return launch_zoom<2, 3, 2, true>(a, dev, st, half_lut);
    case 18: return launch_zoom<3, 1, 0, false>(a, dev, st, half_lut);
    case 19: return launch_zoom<3, 1, 0, true>(a, dev, st, half_lut);
    case 20: return launch_zoom<3, 3, 1, false>(a, dev, st, half_lut);
    case 21: return launch_zoom<3, 3, 1, true>(a, dev, st, half_lut);
    case 22: return launch_zoom<3, 3, 2, false>(a, dev, st, half_lut);
    case 23: return launch_zoom<3, 3, 2, true>(a, dev, st, half_lut);
  }
  return MPVP_E_INVALID;
}
