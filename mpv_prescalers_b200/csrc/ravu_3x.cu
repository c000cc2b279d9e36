// RAVU-3x entry points of the C ABI (kernel: ravu_lite_kernel.cuh, SCALE == 3).
#include "ravu_lite_kernel.cuh"

using namespace mpvp;

extern "C" int mpvp_ravu3x_launch(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                                  const float* in, float* out, int n, int h, int w, int64_t in_stride_n,
                                  int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n, int64_t out_stride_c,
                                  int64_t out_stride_y, int32_t* bucket_out, void* stream) {
  return mpvp_ravu3x_launch_io(lut, key, radius, key_mode, in, out, n, h, w, in_stride_n, in_stride_c, in_stride_y,
                               out_stride_n, out_stride_c, out_stride_y, bucket_out, nullptr, stream);
}

extern "C" int mpvp_ravu3x_launch_io(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                                     const void* in, void* out, int n, int h, int w, int64_t in_stride_n,
                                     int64_t in_stride_c, int64_t in_stride_y, int64_t out_stride_n,
                                     int64_t out_stride_c, int64_t out_stride_y, int32_t* bucket_out,
                                     const mpvp_io* io, void* stream) {
  IoFmt iof;
  if (int rc0 = parse_io(io, iof)) return rc0;
  const int taps = (2 * radius - 1) * (2 * radius - 1);
  const int g = radius == 4 ? 5 : 3;
  int rc = check_common(lut, key, radius, in, out, n, h, w, taps + 1, 216, g * g);
  if (rc) return rc;
  MPVP_REQUIRE(key->n_strength == 3 && key->n_strength_thr == 2, "ravu-3x expects 2 strength thresholds");
  MPVP_REQUIRE(key_mode >= 0 && key_mode <= 2, "key_mode %d", key_mode);
  if (!exact_key()) {
    if (int rck = check_fast_key(key)) return rck;
  }
  if (n == 0) return MPVP_OK;
  DeviceGuard guard(lut->device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", lut->device);
  LiteArgs a{};
  a.io = iof;
  a.in = in; a.out = out; a.lut = reinterpret_cast<const float4*>(lut->lut); a.lut_half = reinterpret_cast<const uint2*>(lut->lut_half); a.bucket = bucket_out;
  a.n = n; a.h = h; a.w = w;
  a.in_sn = in_stride_n; a.in_sc = in_stride_c; a.in_sy = in_stride_y;
  a.out_sn = out_stride_n; a.out_sc = out_stride_c; a.out_sy = out_stride_y;
  a.key = *key;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dev = lut->device;
  switch (radius * 3 + key_mode) {
    case 6: return launch_lite<2, false, 3, 4, 2, 1, 0>(a, dev, st);
    case 7: return launch_lite<2, false, 3, 4, 2, 3, 1>(a, dev, st);
    case 8: return launch_lite<2, false, 3, 4, 2, 3, 2>(a, dev, st);
    case 9: return launch_lite<3, false, 3, 4, 2, 1, 0>(a, dev, st);
    case 10: return launch_lite<3, false, 3, 4, 2, 3, 1>(a, dev, st);
    case 11: return launch_lite<3, false, 3, 4, 2, 3, 2>(a, dev, st);
    case 12: return launch_lite<4, false, 3, 2, 4, 1, 0>(a, dev, st);
    case 13: return launch_lite<4, false, 3, 2, 4, 3, 1>(a, dev, st);
    case 14: return launch_lite<4, false, 3, 2, 4, 3, 2>(a, dev, st);
  }
  return MPVP_E_INVALID;
}

// Host-buffer convenience: H2D, kernel, D2H inside the call (frames are staged in chunks through two
// streams so that copies of chunk k+1 overlap the kernel of chunk k when the host memory is pinned).
