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
#include "common.cuh"

namespace mpvp {
namespace {

struct ZoomArgs {
  const float* __restrict__ in;
  float* __restrict__ out;
  const float4* __restrict__ lut;
  const float4* __restrict__ lut_ar;
  int32_t* __restrict__ bucket;  // [n][oh][ow] or null
  int n, h, w, oh, ow;
  int64_t in_sn, in_sc, in_sy, out_sn, out_sc, out_sy;
  int tiles_x, tiles_y;
  long long total_tiles;
  float ar_strength;
  mpvp_key_params key;
};

constexpr int kTOW = 32, kTOH = 8, kNT = kTOW * kTOH;

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

struct Bilerp {
  int x0, x1, y0, y1;
  float fu, fv;
};

// GL LINEAR + clamp-to-edge at normalised coordinate (cx, cy) of a (w x h) texture.
__device__ __forceinline__ Bilerp make_bilerp(float cx, float cy, int w, int h) {
  const float u = __fsub_rn(__fmul_rn(cx, (float)w), 0.5f);
  const float v = __fsub_rn(__fmul_rn(cy, (float)h), 0.5f);
  const float u0 = floorf(u), v0 = floorf(v);
  Bilerp b;
  b.fu = __fsub_rn(u, u0);
  b.fv = __fsub_rn(v, v0);
  b.x0 = clampi((int)u0, 0, w - 1);
  b.x1 = clampi((int)u0 + 1, 0, w - 1);
  b.y0 = clampi((int)v0, 0, h - 1);
  b.y1 = clampi((int)v0 + 1, 0, h - 1);
  return b;
}

__device__ __forceinline__ float4 fetch_bilerp(const float4* __restrict__ lut, int lw, const Bilerp& b) {
  const float4 t00 = __ldg(lut + (int64_t)b.y0 * lw + b.x0), t10 = __ldg(lut + (int64_t)b.y0 * lw + b.x1);
  const float4 t01 = __ldg(lut + (int64_t)b.y1 * lw + b.x0), t11 = __ldg(lut + (int64_t)b.y1 * lw + b.x1);
  const float gu = 1.0f - b.fu, gv = 1.0f - b.fv;
  float4 r;
  r.x = (t00.x * gu + t10.x * b.fu) * gv + (t01.x * gu + t11.x * b.fu) * b.fv;
  r.y = (t00.y * gu + t10.y * b.fu) * gv + (t01.y * gu + t11.y * b.fu) * b.fv;
  r.z = (t00.z * gu + t10.z * b.fu) * gv + (t01.z * gu + t11.z * b.fu) * b.fv;
  r.w = (t00.w * gu + t10.w * b.fu) * gv + (t01.w * gu + t11.w * b.fu) * b.fv;
  return r;
}

template <int R, int C, int KEYMODE, bool AR>
__global__ void __launch_bounds__(kNT) ravu_zoom_kernel(const __grid_constant__ ZoomArgs A) {
  constexpr int N = 2 * R, TAPS = N * N, G = 4;
  constexpr int B = (TAPS / 2 + 3) / 4;   // LUT blocks per row group (2 for r2, 5 for r3)
  constexpr int LWt = B * 9, LHt = 288 * 9;
  constexpr int SWt = kTOW + 2 * R + 2, SHt = kTOH + 2 * R + 2;
  constexpr int PLANE = SWt * SHt;
  constexpr int NP = (C == 1) ? 1 : 4;  // plane 0 = key plane, 1..3 = colours

  __shared__ float s_src[NP * PLANE];

  const int tid = threadIdx.x;
  const int tx = tid % kTOW, ty = tid / kTOW;

  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x) {
    const int tix = (int)(tile % A.tiles_x);
    const int tiy = (int)((tile / A.tiles_x) % A.tiles_y);
    const int f = (int)(tile / ((long long)A.tiles_x * A.tiles_y));
    const int ox0 = tix * kTOW, oy0 = tiy * kTOH;
    const float* __restrict__ src = A.in + (int64_t)f * A.in_sn;

    int bx_first, by_first;
    float dummy;
    zoom_pos(ox0, A.ow, A.w, bx_first, dummy);
    zoom_pos(oy0, A.oh, A.h, by_first, dummy);
    const int sx0 = bx_first - (R - 1), sy0 = by_first - (R - 1);

    __syncthreads();
    for (int i = tid; i < PLANE; i += kNT) {
      const int sy = i / SWt, sx = i - sy * SWt;
      const int gx = clampi(sx0 + sx, 0, A.w - 1), gy = clampi(sy0 + sy, 0, A.h - 1);
      const int64_t off = (int64_t)gy * A.in_sy + gx;
      if constexpr (C == 1) {
        s_src[i] = __ldg(src + off);
      } else {
        const float c0 = __ldg(src + off), c1 = __ldg(src + A.in_sc + off), c2 = __ldg(src + 2 * A.in_sc + off);
        s_src[i] = (KEYMODE == 2) ? __fadd_rn(__fadd_rn(__fmul_rn(c0, 0.2126f), __fmul_rn(c1, 0.7152f)), __fmul_rn(c2, 0.0722f)) : c0;
        s_src[PLANE + i] = c0;
        s_src[2 * PLANE + i] = c1;
        s_src[3 * PLANE + i] = c2;
      }
    }
    __syncthreads();

    const int ox = ox0 + tx, oy = oy0 + ty;
    if (ox >= A.ow || oy >= A.oh) continue;
    int bx, by;
    float subx, suby;
    zoom_pos(ox, A.ow, A.w, bx, subx);
    zoom_pos(oy, A.oh, A.h, by, suby);
    const float* __restrict__ kb = s_src + (by - by_first) * SWt + (bx - bx_first);  // tap (0,0)

    float ks[TAPS];
#pragma unroll
    for (int t = 0; t < TAPS; ++t) ks[t] = kb[(t % N) * SWt + (t / N)];
    const int row = ravu_key2<STENCIL_RAVU, N, G, 3, true>(A.key, [&](int i, int j) { return ks[i * N + j]; });
    if (A.bucket) A.bucket[((int64_t)f * A.oh + oy) * A.ow + ox] = row;

    // LUT coordinates exactly as the shader forms them (ravu-zoom-r2.hook:24-31,109-112)
    const float px = lutpos9(subx), py = lutpos9(suby);
    const float ipx = __fsub_rn(1.0f, px), ipy = __fsub_rn(1.0f, py);
    const float spx = __fdiv_rn(px, (float)B), sipx = __fdiv_rn(ipx, (float)B);
    const float spy = __fdiv_rn(py, 288.0f), sipy = __fdiv_rn(ipy, 288.0f);
    const float coord_y = __fdiv_rn((float)row, 288.0f);

    float res[C];
    float hi[C], lo[C], hi2[C], lo2[C];
#pragma unroll
    for (int c = 0; c < C; ++c) res[c] = hi[c] = lo[c] = hi2[c] = lo2[c] = 0.f;

    auto sample = [&](int c, int t) -> float {
      if constexpr (C == 1) return ks[t];
      else return kb[(1 + c) * PLANE + (t % N) * SWt + (t / N)];
    };

#pragma unroll
    for (int m = 0; m < 2; ++m) {
#pragma unroll
      for (int blk = 0; blk < B; ++blk) {
        const float blkx = (float)((double)blk / (double)B);  // the literal 0.0 / 0.2 / 0.4 ... of the shader
        const float cx = __fadd_rn(blkx, m ? sipx : spx);
        const float cy = __fadd_rn(coord_y, m ? sipy : spy);
        const Bilerp bl = make_bilerp(cx, cy, LWt, LHt);
        const float4 w4 = fetch_bilerp(A.lut, LWt, bl);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (AR) {
          const float4 a4 = fetch_bilerp(A.lut_ar, LWt, bl);
          av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int k = blk * 4 + e;
          if (k < TAPS / 2) {
            const int t = m ? (TAPS - 1 - k) : k;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const float s = sample(c, t);
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
        const float hiv = hi2[c] / hi[c] - 0.1f;
        const float lov = 1.1f - lo2[c] / lo[c];
        const float cl = fminf(fmaxf(r, lov), hiv);
        r = r * (1.0f - A.ar_strength) + cl * A.ar_strength;
      } else {
        r = fminf(fmaxf(r, 0.f), 1.f);
      }
      __stcs(A.out + (int64_t)f * A.out_sn + c * A.out_sc + (int64_t)oy * A.out_sy + ox, r);
    }
  }
}

template <int R, int C, int KEYMODE, bool AR>
int launch_zoom(const ZoomArgs& a0, int device, cudaStream_t stream) {
  ZoomArgs a = a0;
  a.tiles_x = (a.ow + kTOW - 1) / kTOW;
  a.tiles_y = (a.oh + kTOH - 1) / kTOH;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * a.n;
  auto kern = ravu_zoom_kernel<R, C, KEYMODE, AR>;
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

}  // namespace
}  // namespace mpvp

using namespace mpvp;

extern "C" int mpvp_ravu_zoom_launch(const mpvp_weights* lut, const mpvp_weights* lut_ar, const mpvp_key_params* key,
                                     int radius, int key_mode, float ar_strength, const float* in, float* out, int n,
                                     int h, int w, int out_h, int out_w, int64_t in_stride_n, int64_t in_stride_c,
                                     int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_c,
                                     int64_t out_stride_y, int32_t* bucket_out, void* stream) {
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
  a.in = in; a.out = out; a.bucket = bucket_out;
  a.lut = reinterpret_cast<const float4*>(lut->lut);
  a.lut_ar = lut_ar ? reinterpret_cast<const float4*>(lut_ar->lut) : nullptr;
  a.n = n; a.h = h; a.w = w; a.oh = out_h; a.ow = out_w;
  a.in_sn = in_stride_n; a.in_sc = in_stride_c; a.in_sy = in_stride_y;
  a.out_sn = out_stride_n; a.out_sc = out_stride_c; a.out_sy = out_stride_y;
  a.ar_strength = ar_strength;
  a.key = *key;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dev = lut->device;
  const int ar = lut_ar ? 1 : 0;
  switch ((radius * 3 + key_mode) * 2 + ar) {
    case 12: return launch_zoom<2, 1, 0, false>(a, dev, st);
    case 13: return launch_zoom<2, 1, 0, true>(a, dev, st);
    case 14: return launch_zoom<2, 3, 1, false>(a, dev, st);
    case 15: return launch_zoom<2, 3, 1, true>(a, dev, st);
    case 16: return launch_zoom<2, 3, 2, false>(a, dev, st);
    case 17: return launch_zoom<2, 3, 2, true>(a, dev, st);
    case 18: return launch_zoom<3, 1, 0, false>(a, dev, st);
    case 19: return launch_zoom<3, 1, 0, true>(a, dev, st);
    case 20: return launch_zoom<3, 3, 1, false>(a, dev, st);
    case 21: return launch_zoom<3, 3, 1, true>(a, dev, st);
    case 22: return launch_zoom<3, 3, 2, false>(a, dev, st);
    case 23: return launch_zoom<3, 3, 2, true>(a, dev, st);
  }
  return MPVP_E_INVALID;
}
