// RAVU-Lite(-AR) entry points of the C ABI (kernel: ravu_lite_kernel.cuh).
#include "ravu_lite_kernel.cuh"

using namespace mpvp;

extern "C" int mpvp_ravu_lite_launch(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int ar,
                                     float ar_strength, const float* in, float* out, int n, int h, int w,
                                     int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                                     int64_t out_stride_y, int32_t* bucket_out, void* stream) {
  return mpvp_ravu_lite_launch_io(lut, key, radius, ar, ar_strength, in, out, n, h, w, in_stride_n, in_stride_y,
                                  out_stride_n, out_stride_y, bucket_out, nullptr, stream);
}

extern "C" int mpvp_ravu_lite_launch_io(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int ar,
                                        float ar_strength, const void* in, void* out, int n, int h, int w,
                                        int64_t in_stride_n, int64_t in_stride_y, int64_t out_stride_n,
                                        int64_t out_stride_y, int32_t* bucket_out, const mpvp_io* io, void* stream) {
  IoFmt iof;
  if (int rc0 = parse_io(io, iof)) return rc0;
  const int taps = (2 * radius - 1) * (2 * radius - 1);
  const int g = radius == 4 ? 5 : 3;
  int rc = check_common(lut, key, radius, in, out, n, h, w, (taps + 1) / 2, 288, g * g);
  if (rc) return rc;
  MPVP_REQUIRE(key->n_strength == 4 && key->n_strength_thr == 3, "ravu-lite expects 3 strength thresholds");
  if (!exact_key()) {
    if (int rck = check_fast_key(key)) return rck;
  }
  MPVP_REQUIRE((out_stride_y % 2) == 0 && (out_stride_n % 2) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) % (2 * fmt_bytes(iof.out_fmt))) == 0,
               "output rows must be aligned to a pixel pair (even strides, base aligned to two elements)");
  if (n == 0) return MPVP_OK;
  DeviceGuard guard(lut->device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", lut->device);
  LiteArgs a{};
  a.io = iof;
  a.in = in; a.out = out; a.lut = reinterpret_cast<const float4*>(lut->lut); a.lut_half = reinterpret_cast<const uint2*>(lut->lut_half); a.bucket = bucket_out;
  a.n = n; a.h = h; a.w = w;
  a.in_sn = in_stride_n; a.in_sy = in_stride_y; a.out_sn = out_stride_n; a.out_sy = out_stride_y;
  a.in_sc = 0; a.out_sc = 0;
  a.ar_strength = ar_strength;
  a.key = *key;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (radius * 2 + (ar ? 1 : 0)) {
    case 4: return launch_lite<2, false, 2, MPVP_X_LITE_P, MPVP_X_LITE_STRIPS>(a, lut->device, st);
    case 5: return ravu_lite_ar_dispatch(a, radius, lut->device, st);
    case 6: return launch_lite<3, false, 2, MPVP_X_LITE_P, MPVP_X_LITE_STRIPS>(a, lut->device, st);
    case 7: return ravu_lite_ar_dispatch(a, radius, lut->device, st);
    case 8: return launch_lite<4, false, 2, 2, 4>(a, lut->device, st);
    case 9: return ravu_lite_ar_dispatch(a, radius, lut->device, st);
  }
  return MPVP_E_INVALID;
}
