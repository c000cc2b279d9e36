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

namespace mpvp {
namespace {

constexpr int kMaxTaps = 6;          // 2 * radius, radius <= 3
constexpr int kRTW = 64, kRTH = 32;  // output tile
constexpr int kRNT = 256;

struct ResampleArgs {
  const void* __restrict__ in;
  void* __restrict__ out;
  IoFmt io;
  const int* __restrict__ bx; const float* __restrict__ wx;   // [ow], [ow][taps]
  const int* __restrict__ by; const float* __restrict__ wy;   // [oh], [oh][taps]
  int planes, h, w, oh, ow, taps;
  int sw, sh;   // shared source tile: pitch and rows
  int64_t in_sp, in_sy, out_sp, out_sy;
  int tiles_x, tiles_y;
  long long total_tiles;
};

__global__ void __launch_bounds__(kRNT) resample_kernel(const __grid_constant__ ResampleArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_src = reinterpret_cast<float*>(smem_raw);   // [sh][sw]
  float* s_row = s_src + A.sh * A.sw;                  // [sh][kRTW]: horizontally filtered rows
  __shared__ int s_bx[kRTW], s_by[kRTH];
  __shared__ float s_wx[kRTW][kMaxTaps], s_wy[kRTH][kMaxTaps];
  const int tid = threadIdx.x;
  const int T = A.taps;
  TileWalk walk(blockIdx.x, gridDim.x, A.tiles_x, A.tiles_y);
  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, walk.next()) {
    const int ox0 = walk.tix * kRTW, oy0 = walk.tiy * kRTH, p = walk.f;
    const int nx = min(kRTW, A.ow - ox0), ny = min(kRTH, A.oh - oy0);
    __syncthreads();
    if (tid < kRTW) {
      const int o = ox0 + min(tid, nx - 1);
      s_bx[tid] = A.bx[o];
      for (int i = 0; i < T; ++i) s_wx[tid][i] = A.wx[(size_t)o * T + i];
    } else if (tid < kRTW + kRTH) {
      const int j = tid - kRTW;
      const int o = oy0 + min(j, ny - 1);
      s_by[j] = A.by[o];
      for (int i = 0; i < T; ++i) s_wy[j][i] = A.wy[(size_t)o * T + i];
    }
    // bases are non-decreasing in the output coordinate
    const int x_lo = A.bx[ox0], x_hi = A.bx[ox0 + nx - 1] + T - 1;
    const int y_lo = A.by[oy0], y_hi = A.by[oy0 + ny - 1] + T - 1;
    const int need_w = x_hi - x_lo + 1, need_h = y_hi - y_lo + 1;   // <= sw, sh (checked by the host)
    const int64_t src0 = (int64_t)p * A.in_sp;
    dispatch_in_fmt(A.io.in_fmt, [&](auto ftag) {
      constexpr int FMT = decltype(ftag)::value;
      for (int i = tid; i < need_w * need_h; i += kRNT) {
        const int sy = i / need_w, sx = i - sy * need_w;
        const int gx = clampi(x_lo + sx, 0, A.w - 1), gy = clampi(y_lo + sy, 0, A.h - 1);
        s_src[sy * A.sw + sx] = load_px_t<FMT>(A.in, src0 + (int64_t)gy * A.in_sy + gx, A.io.in_max);
      }
    });
    __syncthreads();
    // rows: s_row[sy][lx] = sum_i wx[lx][i] * src[sy][bx[lx] - x_lo + i]
    for (int i = tid; i < need_h * kRTW; i += kRNT) {
      const int sy = i / kRTW, lx = i - sy * kRTW;
      const float* __restrict__ r = s_src + sy * A.sw + (s_bx[lx] - x_lo);
      float acc = 0.f;
      for (int k = 0; k < T; ++k) acc = fmaf(r[k], s_wx[lx][k], acc);
      s_row[i] = acc;
    }
    __syncthreads();
    for (int i = tid; i < kRTW * kRTH; i += kRNT) {
      const int ly = i / kRTW, lx = i - ly * kRTW;
      if (lx >= nx || ly >= ny) continue;
      const float* __restrict__ c = s_row + (s_by[ly] - y_lo) * kRTW + lx;
      float acc = 0.f;
      for (int k = 0; k < T; ++k) acc = fmaf(c[k * kRTW], s_wy[ly][k], acc);
      store_px(A.out, (int64_t)p * A.out_sp + (int64_t)(oy0 + ly) * A.out_sy + ox0 + lx, acc, A.io.out_fmt, A.io.out_max);
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
  if (int rc = get_axis(device, kernel, h, out_h, offset_y, kRTH, ay)) return rc;
  ResampleArgs a{};
  a.io = iof;
  a.in = in; a.out = out;
  a.bx = ax.base; a.wx = ax.w; a.by = ay.base; a.wy = ay.w;
  a.planes = planes; a.h = h; a.w = w; a.oh = out_h; a.ow = out_w; a.taps = 2 * kernel_radius(kernel);
  a.sw = ax.need | 1; a.sh = ay.need;
  a.in_sp = in_stride_p; a.in_sy = in_stride_y; a.out_sp = out_stride_p; a.out_sy = out_stride_y;
  a.tiles_x = (out_w + kRTW - 1) / kRTW;
  a.tiles_y = (out_h + kRTH - 1) / kRTH;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * planes;
  MPVP_REQUIRE(a.total_tiles < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", a.total_tiles);
  const size_t smem = sizeof(float) * ((size_t)a.sh * a.sw + (size_t)a.sh * kRTW);
  MPVP_CUDA_OK(cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MPVP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, resample_kernel, kRNT, smem));
  MPVP_REQUIRE(per_sm >= 1, "resample kernel does not fit on an SM (smem %zu B)", smem);
  long long grid = (long long)sm_count(device) * per_sm;
  if (grid > a.total_tiles) grid = a.total_tiles;
  grid = cap_grid(grid);
  resample_kernel<<<(unsigned)grid, kRNT, smem, static_cast<cudaStream_t>(stream)>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}
