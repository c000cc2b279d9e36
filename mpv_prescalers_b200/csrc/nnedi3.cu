// NNEDI3: one doubling pass (predictor + interleave) per launch.
//
//   double_y + combine_y   nnedi3-nns16-win8x4.hook:15-104   (compute form: compute/...:15-102)
//   double_x + combine_x   nnedi3-nns16-win8x4.hook:105-194
//
// This file holds the CUDA-core (fp32 FFMA) predictor.  It is the numerically exact on-device
// statement of the shader (fp32 weights, fp32 accumulation, neuron-serial softmax/elliott sums) and
// serves (a) small neuron counts where a 128-row MMA tile cannot be filled economically and (b) as
// the on-GPU cross-check of the tcgen05 path in nnedi3_tc.cu.
#include <cstdlib>

#include "common.cuh"

namespace mpvp {
namespace {

struct NnArgs {
  const float* __restrict__ in;
  float* __restrict__ out;
  const float* __restrict__ w;     // [2*nns][K], rows interleaved (w1_n, w2_n)
  const float* __restrict__ bias;  // [2*nns] interleaved (b1_n, b2_n)
  int n, h, w_;
  int64_t in_sn, in_sy, out_sn, out_sy;
  int nns;
  int tiles_x, tiles_y;
  long long total_tiles;
};

constexpr int kTW = 32, kTH = 8, kNT = 256;

// S = short window side (4 or 6); DIR 0 = double_y (long axis x), 1 = double_x (long axis y)
template <int S, int DIR>
__global__ void __launch_bounds__(kNT) nnedi3_simt_kernel(const __grid_constant__ NnArgs A) {
  constexpr int K = 8 * S;
  constexpr int HX = DIR == 0 ? 8 : S, HY = DIR == 0 ? S : 8;      // window extent in x / y
  constexpr int OX = DIR == 0 ? 3 : (S / 2 - 1), OY = DIR == 0 ? (S / 2 - 1) : 3;  // offset of the window origin
  constexpr int SW = kTW + HX - 1, SH = kTH + HY - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_w = reinterpret_cast<float*>(smem_raw);   // [2*nns][K]
  float* s_b = s_w + 2 * A.nns * K;                  // [2*nns]
  float* s_t = s_b + 2 * A.nns;                      // [SH][SW]
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * A.nns * K; i += kNT) s_w[i] = A.w[i];
  for (int i = tid; i < 2 * A.nns; i += kNT) s_b[i] = A.bias[i];
  const int tx = tid % kTW, ty = tid / kTW;

  for (long long tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x) {
    const unsigned tq = (unsigned)tile / (unsigned)A.tiles_x;   // total_tiles < 2^31 (checked by the host)
    const int tix = (int)((unsigned)tile - tq * (unsigned)A.tiles_x);
    const int f = (int)(tq / (unsigned)A.tiles_y);
    const int tiy = (int)(tq - (unsigned)f * (unsigned)A.tiles_y);
    const int x0 = tix * kTW, y0 = tiy * kTH;
    const float* __restrict__ src = A.in + (int64_t)f * A.in_sn;
    __syncthreads();
    for (int i = tid; i < SW * SH; i += kNT) {
      const int sy = i / SW, sx = i - sy * SW;
      const int gx = clampi(x0 + sx - OX, 0, A.w_ - 1), gy = clampi(y0 + sy - OY, 0, A.h - 1);
      s_t[i] = __ldg(src + (int64_t)gy * A.in_sy + gx);
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    if (x >= A.w_ || y >= A.h) continue;
    // canonical sample order k = a*S + b, a along the long axis (8), b along the short axis (S)
    float xs[K];
    float sum = 0.f, sumsq = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int a = k / S, b = k % S;
      const int dx = DIR == 0 ? a : b, dy = DIR == 0 ? b : a;
      xs[k] = s_t[(ty + dy) * SW + tx + dx];
      sum += xs[k];
      sumsq = fmaf(xs[k], xs[k], sumsq);
    }
    const float mstd0 = sum / (float)K;
    float mstd1 = sumsq / (float)K - mstd0 * mstd0;
    const float mstd2 = mstd1 >= kEps ? rsqrtf(mstd1) : 0.0f;
    mstd1 *= mstd2;
    float vsum = 0.f, wsum = 0.f;
    for (int nn = 0; nn < A.nns; ++nn) {
      const float4* __restrict__ w1 = reinterpret_cast<const float4*>(s_w + (2 * nn) * K);
      const float4* __restrict__ w2 = reinterpret_cast<const float4*>(s_w + (2 * nn + 1) * K);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int q = 0; q < K / 4; ++q) {
        const float4 a = w1[q], b = w2[q];
        s1 = fmaf(xs[4 * q], a.x, s1); s1 = fmaf(xs[4 * q + 1], a.y, s1);
        s1 = fmaf(xs[4 * q + 2], a.z, s1); s1 = fmaf(xs[4 * q + 3], a.w, s1);
        s2 = fmaf(xs[4 * q], b.x, s2); s2 = fmaf(xs[4 * q + 1], b.y, s2);
        s2 = fmaf(xs[4 * q + 2], b.z, s2); s2 = fmaf(xs[4 * q + 3], b.w, s2);
      }
      s1 = expf(fmaf(s1, mstd2, s_b[2 * nn]));
      s2 = fmaf(s2, mstd2, s_b[2 * nn + 1]);
      wsum += s1;
      vsum += s1 * (s2 / (1.0f + fabsf(s2)));
    }
    const float pred = fminf(fmaxf(mstd0 + 5.0f * vsum / wsum * mstd1, 0.f), 1.f);
    const float orig = s_t[(ty + OY) * SW + tx + OX];
    float* __restrict__ o = A.out + (int64_t)f * A.out_sn;
    if (DIR == 0) {  // out(x, 2y) = in, out(x, 2y+1) = interp   (nnedi3-nns16-win8x4.hook:97-104)
      __stcs(o + (int64_t)(2 * y) * A.out_sy + x, orig);
      __stcs(o + (int64_t)(2 * y + 1) * A.out_sy + x, pred);
    } else {         // out(2x, y) = in, out(2x+1, y) = interp
      __stcs(reinterpret_cast<float2*>(o + (int64_t)y * A.out_sy + 2 * x), make_float2(orig, pred));
    }
  }
}

template <int S, int DIR>
int launch_simt(const NnArgs& a0, int device, cudaStream_t stream) {
  constexpr int K = 8 * S;
  constexpr int HX = DIR == 0 ? 8 : S, HY = DIR == 0 ? S : 8;
  NnArgs a = a0;
  const size_t smem = sizeof(float) * ((size_t)2 * a.nns * K + 2 * a.nns + (kTW + HX - 1) * (kTH + HY - 1));
  a.tiles_x = (a.w_ + kTW - 1) / kTW;
  a.tiles_y = (a.h + kTH - 1) / kTH;
  a.total_tiles = (long long)a.tiles_x * a.tiles_y * a.n;
  MPVP_REQUIRE(a.total_tiles < (1LL << 31), "batch too large: %lld tiles (limit 2^31)", a.total_tiles);
  auto kern = nnedi3_simt_kernel<S, DIR>;
  MPVP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  MPVP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kNT, smem));
  if (per_sm < 1) {
    set_error("nnedi3 kernel does not fit on an SM (smem %zu B)", smem);
    return MPVP_E_UNSUPPORTED;
  }
  long long grid = (long long)sm_count(device) * per_sm;
  if (grid > a.total_tiles) grid = a.total_tiles;
  grid = cap_grid(grid);
  if (grid < 1) return MPVP_OK;
  kern<<<(unsigned)grid, kNT, smem, stream>>>(a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MPVP_CUDA_OK(cudaGetLastError());
  return MPVP_OK;
}

}  // namespace

int nnedi3_simt(const mpvp_weights* nn, int direction, const float* in, float* out, int n, int h, int w,
                int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_y,
                cudaStream_t st) {
  NnArgs a{};
  a.in = in; a.out = out; a.w = nn->nn_w; a.bias = nn->nn_bias + 2 * nn->nns;  // unscaled biases
  a.n = n; a.h = h; a.w_ = w;
  a.in_sn = in_stride_n; a.in_sy = in_stride_y; a.out_sn = out_stride_n; a.out_sy = out_stride_y;
  a.nns = nn->nns;
  if (nn->win_short == 4) return direction == 0 ? launch_simt<4, 0>(a, nn->device, st) : launch_simt<4, 1>(a, nn->device, st);
  return direction == 0 ? launch_simt<6, 0>(a, nn->device, st) : launch_simt<6, 1>(a, nn->device, st);
}

int nnedi3_tc(const mpvp_weights* nn, int direction, const void* in, void* out, int n, int h, int w,
              int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_y, const IoFmt& io,
              cudaStream_t st);

// MPVP_NNEDI3_IMPL=simt forces the CUDA-core predictor (the on-device cross-check of the tensor path, used by
// tests/test_gpu_parity.py); default is tcgen05.  Read at every launch so that a test can switch it.
static bool use_simt() {
  const char* e = getenv("MPVP_NNEDI3_IMPL");
  return e && e[0] == 's';
}

}  // namespace mpvp

using namespace mpvp;

extern "C" int mpvp_nnedi3_launch(const mpvp_weights* nn, int direction, const float* in, float* out, int n, int h,
                                  int w, int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                                  int64_t out_stride_y, void* stream) {
  return mpvp_nnedi3_launch_io(nn, direction, in, out, n, h, w, in_stride_n, in_stride_y, out_stride_n, out_stride_y,
                               nullptr, stream);
}

extern "C" int mpvp_nnedi3_launch_io(const mpvp_weights* nn, int direction, const void* in, void* out, int n, int h,
                                     int w, int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                                     int64_t out_stride_y, const mpvp_io* io, void* stream) {
  IoFmt iof;
  if (int rc0 = parse_io(io, iof)) return rc0;
  MPVP_REQUIRE(nn && nn->kind == 1 && nn->nn_w && nn->nn_bias, "nn handle is null or not an NNEDI3 weight set");
  MPVP_REQUIRE(direction == 0 || direction == 1, "direction %d", direction);
  MPVP_REQUIRE(in && out, "null frame pointer");
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad frame geometry n=%d h=%d w=%d", n, h, w);
  if (direction == 1)
    MPVP_REQUIRE((out_stride_y % 2) == 0 && (out_stride_n % 2) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) % (2 * fmt_bytes(iof.out_fmt))) == 0,
                 "double_x output rows must be aligned to a pixel pair (even strides, base aligned to two elements)");
  if (n == 0) return MPVP_OK;
  DeviceGuard guard(nn->device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", nn->device);
  if (use_simt()) {
    MPVP_REQUIRE(iof.in_fmt == MPVP_FMT_F32 && iof.out_fmt == MPVP_FMT_F32, "the CUDA-core cross-check path is float32 only");
    return nnedi3_simt(nn, direction, static_cast<const float*>(in), static_cast<float*>(out), n, h, w, in_stride_n,
                       in_stride_y, out_stride_n, out_stride_y, static_cast<cudaStream_t>(stream));
  }
  MPVP_REQUIRE(nn->nn_b, "NNEDI3 weight set has no packed tensor-core operand");
  return nnedi3_tc(nn, direction, in, out, n, h, w, in_stride_n, in_stride_y, out_stride_n, out_stride_y, iof,
                   static_cast<cudaStream_t>(stream));
}
