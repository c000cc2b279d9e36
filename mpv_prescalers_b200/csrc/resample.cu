// The step after the path: the offset-correcting main scaler.
//
// ravu and nnedi3 declare `//!OFFSET -0.5 -0.5` (ravu-r2.hook:325; nnedi3-nns16-win8x4.hook:95,185): their output texel
// X holds the image content that belongs half a texel further right / down (out(2x, 2y) = HOOKED(x, y),
// ravu-r2.hook:327-338).  The reference leaves the correction to the host: mpv accumulates the offsets of the hooked
// plane and its main scaler (`--scale`, video/out/gpu/video.c + filter_kernels.c -- third-party, not in the reference
// snapshot) samples the plane at the shifted position while it resizes to the output size.  This file is that step:
// a separable polyphase resampler
//
//     out(ox, oy) = sum_j sum_i  wy[oy][j] * wx[ox][i] * in(bx[ox] + i, by[oy] + j)           (clamp-to-edge)
//     s(o) = (o + 0.5) * I / O - 0.5 + offset,  b = floor(s) - R + 1,  w[i] = K(b + i - s) / sum_i K(b + i - s)
//
// with mpv's filter kernels K (bilinear, catmull_rom, mitchell, spline36, lanczos = sinc windowed by sinc, radius 3) and
// mpv's default of NOT widening the kernel when downscaling (--correct-downscaling=no).  The per-coordinate tables
// (base, weights) are computed once per geometry on the host in double precision, rounded to float32 and cached on the
// device; the kernel stages a source tile in shared memory (clamp-to-edge), filters rows into a second shared tile and
// columns from there, so every source texel is read from HBM once: 4 B in + 4 B out per pixel at the same size.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace mpvp {
namespace {

constexpr int kMaxTaps = 6;          // 2 * radius, radius <= 3
constexpr int kRTW = 64;             // output tile: 64 x TH, TH = 64 (16 pixels per thread: the per-tile table loads and barriers
                                     // amortise) or 32 when the plane is reduced (the staged source tile grows with the ratio)
constexpr int kRNT = 256;

struct ResampleArgs {
  const void* __restrict__ in;
  void* __restrict__ out;
  IoFmt io;
  const int* __restrict__ bx; const float* __restrict__ wx;   // [ow], [ow][taps]
  const int* __restrict__ by; const float* __restrict__ wy;   // [oh], [oh][taps]
  int planes, h, w, oh, ow, taps;
  int sw, sh;   // shared source tile: pitch and rows
  int tma_w;    // TMA variant: box width = pitch
  int64_t in_sp, in_sy, out_sp, out_sy;
  int tiles_x, tiles_y;
  long long total_tiles;
};

// 256 threads = 64 columns x 4 row groups: thread (lx, ty) owns output column ox0 + lx (its base and T weights sit in
// registers for the whole tile) and every 4th row.  Compile-time: T taps (2, 4, 6: both filter loops unrolled), the pitch
// SWT of the staged tile (73 when the axis is not reduced, 137 for up to 2x down), F32 = float32 planes on both sides
// (plain loads / stores; other formats go through the generic per-pixel conversions).  No index is divided at run time;
// tiles that do not touch the border skip the clamps.  (The first version -- run-time tap count, one flat index per
// element -- spent 160 warp instructions per 32 pixels, 70 % issue-bound at 14 % of the HBM roofline.)
template <int T, int SWT, bool F32, int TH>
__global__ void __launch_bounds__(kRNT) resample_kernel(const __grid_constant__ ResampleArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* s_src = reinterpret_cast<float*>(smem_raw);   // [sh][SWT]
  float* s_row = s_src + A.sh * SWT;                   // [sh][kRTW]: horizontally filtered rows
  __shared__ int s_by[TH];
  __shared__ __align__(16) float s_wy[TH][8];
  const int tid = threadIdx.x;
  const int lx = tid & (kRTW - 1), ty = tid / kRTW;    // ty = 0 .. RG - 1
  constexpr int RG = kRNT / kRTW;
  TileWalk walk(blockIdx.x, gridDim.x, A.tiles_x, A.tiles_y);
  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, walk.next()) {
    const int ox0 = walk.tix * kRTW, oy0 = walk.tiy * TH, p = walk.f;
    const int nx = min(kRTW, A.ow - ox0), ny = min(TH, A.oh - oy0);
    // this thread's column: base texel and weights
    const int oxc = ox0 + min(lx, nx - 1);
    const int bxc = __ldg(A.bx + oxc);
    float wxr[T];
#pragma unroll
    for (int k = 0; k < T; ++k) wxr[k] = __ldg(A.wx + (size_t)oxc * T + k);
    __syncthreads();   // the previous tile is fully consumed
    if (tid < TH) {
      const int o = oy0 + min(tid, ny - 1);
      s_by[tid] = __ldg(A.by + o);
#pragma unroll
      for (int k = 0; k < T; ++k) s_wy[tid][k] = __ldg(A.wy + (size_t)o * T + k);
    }
    // bases are non-decreasing in the output coordinate
    const int x_lo = __ldg(A.bx + ox0), x_hi = __ldg(A.bx + ox0 + nx - 1) + T - 1;
    const int y_lo = __ldg(A.by + oy0), y_hi = __ldg(A.by + oy0 + ny - 1) + T - 1;
    const int need_w = x_hi - x_lo + 1, need_h = y_hi - y_lo + 1;   // <= SWT, sh (checked by the host)
    const int64_t src0 = (int64_t)p * A.in_sp;
    const bool inner = x_lo >= 0 && y_lo >= 0 && x_hi < A.w && y_hi < A.h && A.in_sy < (1 << 24);   // CTA-uniform
    if (F32 && inner) {
      // interior tile of float32 planes: the row filter reads its T taps straight from global memory (the tile's texels
      // are touched T times, all but the first from L1) -- no staging pass, one barrier less.  Row offsets within a tile
      // fit 32 bits.
      const float* __restrict__ g = static_cast<const float*>(A.in) + src0 + (int64_t)y_lo * A.in_sy + bxc;
      const int pitch = (int)A.in_sy;
      float* __restrict__ d = s_row + lx;
#pragma unroll 4
      for (int sy = ty; sy < need_h; sy += RG) {
        const float* __restrict__ r = g + sy * pitch;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) acc = fmaf(__ldg(r + k), wxr[k], acc);
        d[sy * kRTW] = acc;
      }
    } else {
      dispatch_in_fmt(A.io.in_fmt, [&](auto ftag) {
        constexpr int FMT = decltype(ftag)::value;
        for (int sy = ty; sy < need_h; sy += RG) {
          const int64_t row0 = src0 + (int64_t)clampi(y_lo + sy, 0, A.h - 1) * A.in_sy;
          for (int sx = lx; sx < need_w; sx += kRTW)
            s_src[sy * SWT + sx] = load_px_t<FMT>(A.in, row0 + clampi(x_lo + sx, 0, A.w - 1), A.io.in_max);
        }
      });
      __syncthreads();
      // rows: s_row[sy][lx] = sum_k wx[lx][k] * src[sy][bx[lx] - x_lo + k]
      const float* __restrict__ r = s_src + ty * SWT + (bxc - x_lo);
      float* __restrict__ d = s_row + ty * kRTW + lx;
      for (int sy = ty; sy < need_h; sy += RG, r += RG * SWT, d += RG * kRTW) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) acc = fmaf(r[k], wxr[k], acc);
        *d = acc;
      }
    }
    __syncthreads();
    {
      auto column = [&](int ly, auto store) {
        const float* __restrict__ c = s_row + (s_by[ly] - y_lo) * kRTW + lx;
        float wy[8];
        *reinterpret_cast<float4*>(wy) = *reinterpret_cast<const float4*>(&s_wy[ly][0]);
        if (T > 4) *reinterpret_cast<float2*>(wy + 4) = *reinterpret_cast<const float2*>(&s_wy[ly][4]);
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) acc = fmaf(c[k * kRTW], wy[k], acc);
        store(ly, acc);
      };
      const int64_t o0 = (int64_t)p * A.out_sp + (int64_t)oy0 * A.out_sy + ox0 + lx;
      if (F32 && nx == kRTW && ny == TH && A.out_sy < (1 << 24)) {
        // full tile, float32 out: no predicates, 32-bit row offsets.  The thread takes TH / RG CONSECUTIVE rows and keeps its
        // T filtered-row values in registers: when the base of the next row is the same or one further (every ratio >= 1),
        // the window slides by at most one shared-memory read instead of T (the test is warp-uniform: a warp shares ly)
        float* __restrict__ q = static_cast<float*>(A.out) + o0;
        const int pitch = (int)A.out_sy;
        constexpr int RPT = TH / RG;
        float win[T];
        int wb = -(1 << 20);
#pragma unroll 4
        for (int j = 0; j < RPT; ++j) {
          const int ly = ty * RPT + j;
          const int b = s_by[ly] - y_lo;
          const float* __restrict__ c = s_row + b * kRTW + lx;
          if (T >= 4 && b == wb + 1) {                  // (two taps: sliding costs what it saves)
#pragma unroll
            for (int k = 0; k + 1 < T; ++k) win[k] = win[k + 1];
            win[T - 1] = c[(T - 1) * kRTW];
          } else if (T < 4 || b != wb) {
#pragma unroll
            for (int k = 0; k < T; ++k) win[k] = c[k * kRTW];
          }
          wb = b;
          float wy[8];
          *reinterpret_cast<float4*>(wy) = *reinterpret_cast<const float4*>(&s_wy[ly][0]);
          if (T > 4) *reinterpret_cast<float2*>(wy + 4) = *reinterpret_cast<const float2*>(&s_wy[ly][4]);
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < T; ++k) acc = fmaf(win[k], wy[k], acc);
          __stcs(q + ly * pitch, acc);
        }
      } else if (lx < nx) {
        for (int ly = ty; ly < ny; ly += RG)
          column(ly, [&](int l, float v) { store_px(A.out, o0 + (int64_t)l * A.out_sy, v, A.io.out_fmt, A.io.out_max); });
      }
    }
  }
}

// float32 planes, no reduction (ratio >= 1 on both axes), TMA-legal layout: the source tile arrives by cp.async.bulk.tensor
// into one of two staging buffers -- the box starts at the 16-byte aligned texel at or left of the tile's first texel and
// is PW texels wide; out-of-image texels arrive as zeros and border tiles are patched to clamp-to-edge -- while the
// previous tile is filtered: no staging instructions, no global loads in the filters.  Two barriers per tile.
constexpr int kTmaWMax = 76;   // box width <= 73 + 3 texels of slack for the aligned origin, a multiple of 4 texels
template <int T, int TH, int PW>
__global__ void __launch_bounds__(kRNT) resample_tma_kernel(const __grid_constant__ ResampleArgs A, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // PW: box width = pitch of the staged tile, compile-time (76 for ratios near 1, 44 from 2x up)
  const int BUF = (A.sh * PW + 31) & ~31;               // floats per staging buffer (128-byte multiple)
  float* s_src = reinterpret_cast<float*>(smem_raw);   // [2][sh][PW]
  float* s_row = s_src + 2 * BUF;                      // [sh][kRTW]
  __shared__ int s_by2[2][TH];
  __shared__ __align__(16) float s_wy2[2][TH][8];
  __shared__ __align__(8) uint64_t s_mbar[2];
  const int tid = threadIdx.x;
  const int lx = tid & (kRTW - 1), ty = tid / kRTW;
  constexpr int RG = kRNT / kRTW, RPT = TH / RG;
  if (tid == 0) {
    mbar_init1(smem_addr(&s_mbar[0]));
    mbar_init1(smem_addr(&s_mbar[1]));
    mbar_init_fence();
  }
  __syncthreads();
  const uint64_t tmap_ptr = reinterpret_cast<uint64_t>(&tmap);
  auto issue = [&](const TileWalk& w, int buf) {     // one elected thread of warp 0
    const int x_lo = __ldg(A.bx + w.tix * kRTW), y_lo = __ldg(A.by + w.tiy * TH);
    const uint32_t bar = smem_addr(&s_mbar[buf]);
    tma_expect(bar, (uint32_t)(PW * A.sh * 4));
    tma_load_3d(smem_addr(s_src + buf * BUF), tmap_ptr, x_lo & ~3, y_lo, w.f, bar);
  };
  TileWalk walk(blockIdx.x, gridDim.x, A.tiles_x, A.tiles_y);
  TileWalk ahead = walk;
  if (tid < 32 && blockIdx.x < A.total_tiles) {
    if (elect_one()) issue(ahead, 0);
  }
  ahead.next();
  uint32_t it = 0;
  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, walk.next(), ahead.next(), ++it) {
    const int ox0 = walk.tix * kRTW, oy0 = walk.tiy * TH, p = walk.f;
    const int nx = min(kRTW, A.ow - ox0), ny = min(TH, A.oh - oy0);
    const int oxc = ox0 + min(lx, nx - 1);
    const int bxc = __ldg(A.bx + oxc);
    float wxr[T];
#pragma unroll
    for (int k = 0; k < T; ++k) wxr[k] = __ldg(A.wx + (size_t)oxc * T + k);
    const int x_lo = __ldg(A.bx + ox0), y_lo = __ldg(A.by + oy0);
    const int need_h = __ldg(A.by + oy0 + ny - 1) + T - y_lo;
    const int ax0 = x_lo & ~3;
    int* __restrict__ s_by = s_by2[it & 1];
    float (*__restrict__ s_wy)[8] = s_wy2[it & 1];
    if (tid < TH) {
      const int o = oy0 + min(tid, ny - 1);
      s_by[tid] = __ldg(A.by + o);
#pragma unroll
      for (int k = 0; k < T; ++k) s_wy[tid][k] = __ldg(A.wy + (size_t)o * T + k);
    }
    __syncthreads();   // the column filter of the previous tile is done with s_row; the other staging buffer is free
    if (tid < 32 && tile + gridDim.x < A.total_tiles) {
      fence_proxy_async_smem();
      if (elect_one()) issue(ahead, (it + 1) & 1);
    }
    float* __restrict__ buf = s_src + (it & 1) * BUF;
    mbar_wait_parity(smem_addr(&s_mbar[it & 1]), (it >> 1) & 1);
    const bool edge = ax0 < 0 || y_lo < 0 || ax0 + PW > A.w || y_lo + A.sh > A.h;
    if (edge) patch_clamp_to_edge(buf, PW, PW, A.sh, ax0, y_lo, A.w, A.h, tid, kRNT, [] { __syncthreads(); });
    {
      const float* __restrict__ r = buf + ty * PW + (bxc - ax0);
      float* __restrict__ d = s_row + ty * kRTW + lx;
#pragma unroll 2
      for (int sy = ty; sy < need_h; sy += RG, r += RG * PW, d += RG * kRTW) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) acc = fmaf(r[k], wxr[k], acc);
        *d = acc;
      }
    }
    __syncthreads();
    float* __restrict__ q = static_cast<float*>(A.out) + (int64_t)p * A.out_sp + (int64_t)oy0 * A.out_sy + ox0 + lx;
    if (lx < nx) {
      float win[T];
      int wb = -(1 << 20);
#pragma unroll 4
      for (int j = 0; j < RPT; ++j) {
        const int ly = ty * RPT + j;
        if (ly >= ny) break;
        const int b = s_by[ly] - y_lo;
        const float* __restrict__ c = s_row + b * kRTW + lx;
        if (T >= 4 && b == wb + 1) {
#pragma unroll
          for (int k = 0; k + 1 < T; ++k) win[k] = win[k + 1];
          win[T - 1] = c[(T - 1) * kRTW];
        } else if (T < 4 || b != wb) {
#pragma unroll
          for (int k = 0; k < T; ++k) win[k] = c[k * kRTW];
        }
        wb = b;
        const float4 w03 = *reinterpret_cast<const float4*>(&s_wy[ly][0]);
        const float2 w45 = *reinterpret_cast<const float2*>(&s_wy[ly][4]);
        const float wy[6] = {w03.x, w03.y, w03.z, w03.w, w45.x, w45.y};
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) acc = fmaf(win[k], wy[k], acc);
        __stcs(q + (int64_t)ly * A.out_sy, acc);
      }
    }
  }
}

// ---- filter kernels (mpv: video/out/filter_kernels.c), evaluated in double on the host ------------------------------------
double cubic_bc(double x, double B, double C) {
  x = fabs(x);
  if (x < 1.0) return ((12 - 9 * B - 6 * C) * x * x * x + (-18 + 12 * B + 6 * C) * x * x + (6 - 2 * B)) / 6.0;
  if (x < 2.0) return ((-B - 6 * C) * x * x * x + (6 * B + 30 * C) * x * x + (-12 * B - 48 * C) * x + (8 * B + 24 * C)) / 6.0;
  return 0.0;
}
double sinc(double x) {
  if (fabs(x) < 1e-8) return 1.0;
  const double px = M_PI * x;
  return sin(px) / px;
}
double spline36(double x) {
  x = fabs(x);
  if (x < 1.0) return ((13.0 / 11.0 * x - 453.0 / 209.0) * x - 3.0 / 209.0) * x + 1.0;
  if (x < 2.0) { x -= 1.0; return ((-6.0 / 11.0 * x + 270.0 / 209.0) * x - 156.0 / 209.0) * x; }
  if (x < 3.0) { x -= 2.0; return ((1.0 / 11.0 * x - 45.0 / 209.0) * x + 26.0 / 209.0) * x; }
  return 0.0;
}
int kernel_radius(int kernel) {
  switch (kernel) {
    case MPVP_SCALER_BILINEAR: return 1;
    case MPVP_SCALER_CATMULL_ROM: case MPVP_SCALER_MITCHELL: return 2;
    case MPVP_SCALER_SPLINE36: case MPVP_SCALER_LANCZOS: return 3;
  }
  return 0;
}
double kernel_eval(int kernel, double x) {
  switch (kernel) {
    case MPVP_SCALER_BILINEAR: return fabs(x) < 1.0 ? 1.0 - fabs(x) : 0.0;
    case MPVP_SCALER_CATMULL_ROM: return cubic_bc(x, 0.0, 0.5);
    case MPVP_SCALER_MITCHELL: return cubic_bc(x, 1.0 / 3.0, 1.0 / 3.0);
    case MPVP_SCALER_SPLINE36: return spline36(x);
    case MPVP_SCALER_LANCZOS: return fabs(x) < 3.0 ? sinc(x) * sinc(x / 3.0) : 0.0;
  }
  return 0.0;
}

struct AxisTable {
  int* base = nullptr;     // device [O]
  float* w = nullptr;      // device [O][taps]
  int need = 0;            // largest source extent of a tile of `tile` output coordinates
};
struct AxisKey {
  int device, kernel, I, O, tile;
  float off;
  bool operator<(const AxisKey& o) const {
    return memcmp(this, &o, sizeof(AxisKey)) < 0;
  }
};
std::mutex g_axis_mu;
std::map<AxisKey, AxisTable> g_axis;   // lives for the life of the library (a handful of geometries per process)

int get_axis(int device, int kernel, int I, int O, float off, int tile, AxisTable& out) {
  AxisKey key;
  memset(&key, 0, sizeof(key));
  key.device = device; key.kernel = kernel; key.I = I; key.O = O; key.tile = tile; key.off = off;
  std::lock_guard<std::mutex> lk(g_axis_mu);
  auto it = g_axis.find(key);
  if (it != g_axis.end()) {
    out = it->second;
    return MPVP_OK;
  }
  const int R = kernel_radius(kernel), T = 2 * R;
  std::vector<int> base(O);
  std::vector<float> w((size_t)O * T);
  for (int o = 0; o < O; ++o) {
    const double s = (o + 0.5) * (double)I / (double)O - 0.5 + (double)off;
    const double fl = floor(s);
    base[o] = (int)fl - R + 1;
    double k[kMaxTaps], sum = 0.0;
    for (int i = 0; i < T; ++i) {
      k[i] = kernel_eval(kernel, (double)(base[o] + i) - s);
      sum += k[i];
    }
    for (int i = 0; i < T; ++i) w[(size_t)o * T + i] = (float)(k[i] / sum);
  }
  AxisTable t;
  for (int o = 0; o < O; o += tile) {
    const int last = (o + tile < O ? o + tile : O) - 1;
    const int ext = base[last] + T - 1 - base[o] + 1;
    if (ext > t.need) t.need = ext;
  }
  cudaError_t e = cudaMalloc(&t.base, sizeof(int) * O);
  if (e == cudaSuccess) e = cudaMemcpy(t.base, base.data(), sizeof(int) * O, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&t.w, sizeof(float) * w.size());
  if (e == cudaSuccess) e = cudaMemcpy(t.w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("resample table upload failed: %s", cudaGetErrorString(e));
    if (t.base) cudaFree(t.base);
    if (t.w) cudaFree(t.w);
    return MPVP_E_CUDA;
  }
  g_axis[key] = t;
  out = t;
  return MPVP_OK;
}

}  // namespace
}  // namespace mpvp

using namespace mpvp;

extern "C" int mpvp_resample_launch_io(int device, int kernel, const void* in, void* out, int planes, int h, int w, int out_h,
                                       int out_w, float offset_x, float offset_y, int64_t in_stride_p, int64_t in_stride_y,
                                       int64_t out_stride_p, int64_t out_stride_y, const mpvp_io* io, void* stream) {
  IoFmt iof;
  if (int rc0 = parse_io(io, iof)) return rc0;
  MPVP_REQUIRE(kernel_radius(kernel) > 0, "unknown scaler kernel %d", kernel);
  MPVP_REQUIRE(in && out, "null plane pointer");
  MPVP_REQUIRE(planes >= 0 && h >= 1 && w >= 1 && out_h >= 1 && out_w >= 1, "bad geometry");
  MPVP_REQUIRE(2LL * out_w >= w && 2LL * out_h >= h, "downscaling by more than 2x is not supported (%dx%d -> %dx%d)", w, h, out_w, out_h);
  MPVP_REQUIRE(fabsf(offset_x) <= 8.0f && fabsf(offset_y) <= 8.0f, "offset out of range");
  if (planes == 0) return MPVP_OK;
  DeviceGuard guard(device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", device);
  AxisTable ax, ay;
  if (int rc = get_axis(device, kernel, w, out_w, offset_x, kRTW, ax)) return rc;
  const int th = (out_h < h || out_w < w) ? 32 : 64;
  if (int rc = get_axis(device, kernel, h, out_h, offset_y, th, ay)) return rc;
  ResampleArgs a{};
  a.io = iof;
  a.in = in; a.out = out;
  a.bx = ax.base; a.wx = ax.w; a.by = ay.base; a.wy = ay.w;
  a.planes = planes; a.h = h; a.w = w; a.oh = out_h; a.ow = out_w; a.taps = 2 * kernel_radius(kernel);
  a.sw = ax.need > 73 ? 137 : 73; a.sh = ay.need;    // the two compile-time pitches (odd: columns spread over the banks)
  MPVP_REQUIRE(ax.need <= 137, "internal: staged tile width %d", ax.need);
  a.in_sp = in_stride_p; a.in_sy = in_stride_y; a.out_sp = out_stride_p; a.out_sy = out_stride_y;
  a.tiles_x = (out_w + kRTW - 1) / kRTW;
  a.tiles_y = (out_h + th - 1) / th;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * planes;
  MPVP_REQUIRE(a.total_tiles < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", a.total_tiles);
  const size_t smem = sizeof(float) * ((size_t)a.sh * a.sw + (size_t)a.sh * kRTW);
  const bool f32 = iof.in_fmt == MPVP_FMT_F32 && iof.out_fmt == MPVP_FMT_F32;
  auto pick = [&](auto tag) {
    constexpr int T = decltype(tag)::value;
    if (th == 64) {
      if (a.sw == 73) return f32 ? resample_kernel<T, 73, true, 64> : resample_kernel<T, 73, false, 64>;
      return f32 ? resample_kernel<T, 137, true, 64> : resample_kernel<T, 137, false, 64>;
    }
    if (a.sw == 73) return f32 ? resample_kernel<T, 73, true, 32> : resample_kernel<T, 73, false, 32>;
    return f32 ? resample_kernel<T, 137, true, 32> : resample_kernel<T, 137, false, 32>;
  };
  auto kern = a.taps == 2 ? pick(std::integral_constant<int, 2>{})
              : (a.taps == 4 ? pick(std::integral_constant<int, 4>{}) : pick(std::integral_constant<int, 6>{}));
  // TMA-staged variant: float32 planes, no reduction (th == 64, staged width <= 73), layout within TMA's 16-byte rules
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  a.tma_w = (ax.need + 3 + 3) & ~3;     // staged extent + up to 3 texels left of it (aligned origin), a multiple of 4
  a.tma_w = a.tma_w <= 44 ? 44 : kTmaWMax;
  if (f32 && th == 64 && ax.need + 3 <= kTmaWMax && a.sh <= 256 &&
      make_plane_tmap(&tmap, in, 4, w, h, planes, in_stride_y, in_stride_p, a.tma_w, a.sh)) {
    auto tpick = [&](auto tag) {
      constexpr int T = decltype(tag)::value;
      return a.tma_w == 44 ? resample_tma_kernel<T, 64, 44> : resample_tma_kernel<T, 64, kTmaWMax>;
    };
    auto tk = a.taps == 2 ? tpick(std::integral_constant<int, 2>{})
              : (a.taps == 4 ? tpick(std::integral_constant<int, 4>{}) : tpick(std::integral_constant<int, 6>{}));
    const size_t tsm = sizeof(float) * (2 * (size_t)((a.sh * a.tma_w + 31) & ~31) + (size_t)a.sh * kRTW) + 128;
    MPVP_CUDA_OK(cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
    int per = 0;
    MPVP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, tk, kRNT, tsm));
    if (per >= 1) {
      long long grid = (long long)sm_count(device) * per;
      if (grid > a.total_tiles) grid = a.total_tiles;
      grid = cap_grid(grid);
      tk<<<(unsigned)grid, kRNT, tsm, static_cast<cudaStream_t>(stream)>>>(a, tmap);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      MPVP_CUDA_OK(cudaGetLastError());
      return MPVP_OK;
    }
  }
  MPVP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MPVP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRNT, smem));
  MPVP_REQUIRE(per_sm >= 1, "resample kernel does not fit on an SM (smem %zu B)", smem);
  long long grid = (long long)sm_count(device) * per_sm;
  if (grid > a.total_tiles) grid = a.total_tiles;
  grid = cap_grid(grid);
  kern<<<(unsigned)grid, kRNT, smem, static_cast<cudaStream_t>(stream)>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}
