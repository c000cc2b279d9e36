// Host-buffer entry points of the C ABI: HOST pointers in (pageable or pinned), HOST pointers out.  Frames are staged
// through device scratch in chunks on two streams, so that the host->device copy of chunk k+1 and the device->host copy
// of chunk k-1 overlap the kernel of chunk k when the host memory is pinned; the kernels are the ones the device entry
// points launch.  These calls allocate their scratch, synchronise before returning, and are meant for callers that do not
// manage device memory themselves (the reference's host player owns all textures: this is the closest C analogue of
// "hand the plane to the shader and get the scaled plane back").
#include <functional>

#include "common.cuh"

using namespace mpvp;

namespace {

// launch(device_in, device_out, frames_in_chunk, stream) -> MPVP_* code
int run_host(int device, const void* host_in, void* host_out, int n, size_t in_frame_bytes, size_t out_frame_bytes,
             const std::function<int(const void*, void*, int, cudaStream_t)>& launch, const char* who) {
  MPVP_REQUIRE(host_in && host_out, "%s: null host pointer", who);
  if (n == 0) return MPVP_OK;
  DeviceGuard guard(device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", device);
  const size_t per_frame = in_frame_bytes + out_frame_bytes;
  int chunk = (int)((size_t)(64u << 20) / (per_frame ? per_frame : 1));
  if (chunk < 1) chunk = 1;
  if (chunk > n) chunk = n;
  cudaStream_t st[2] = {nullptr, nullptr};
  void* din[2] = {nullptr, nullptr};
  void* dout[2] = {nullptr, nullptr};
  int rc = MPVP_OK;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&din[i], chunk * in_frame_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&dout[i], chunk * out_frame_bytes);
  }
  for (int f0 = 0, k = 0; f0 < n && e == cudaSuccess && rc == MPVP_OK; f0 += chunk, ++k) {
    const int b = k & 1, m = (n - f0) < chunk ? (n - f0) : chunk;
    e = cudaMemcpyAsync(din[b], static_cast<const char*>(host_in) + (size_t)f0 * in_frame_bytes, m * in_frame_bytes,
                        cudaMemcpyHostToDevice, st[b]);
    if (e != cudaSuccess) break;
    rc = launch(din[b], dout[b], m, st[b]);
    if (rc != MPVP_OK) break;
    e = cudaMemcpyAsync(static_cast<char*>(host_out) + (size_t)f0 * out_frame_bytes, dout[b], m * out_frame_bytes,
                        cudaMemcpyDeviceToHost, st[b]);
  }
  for (int i = 0; i < 2; ++i) {
    if (st[i]) {
      cudaError_t e2 = cudaStreamSynchronize(st[i]);
      if (e == cudaSuccess) e = e2;
      cudaStreamDestroy(st[i]);
    }
    if (din[i]) cudaFree(din[i]);
    if (dout[i]) cudaFree(dout[i]);
  }
  if (rc != MPVP_OK) return rc;
  if (e != cudaSuccess) {
    set_error("%s: %s", who, cudaGetErrorString(e));
    return MPVP_E_CUDA;
  }
  return MPVP_OK;
}

int io_bytes(const mpvp_io* io, bool out) {
  if (!io) return 4;
  return fmt_bytes(out ? io->out_format : io->in_format);
}

}  // namespace

extern "C" int mpvp_ravu_lite_host(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int ar,
                                   float ar_strength, const float* host_in, float* host_out, int n, int h, int w) {
  MPVP_REQUIRE(lut && lut->kind == 0, "lut handle is null or not a LUT");
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad frame geometry n=%d h=%d w=%d", n, h, w);
  const size_t in_frame = (size_t)h * w, out_frame = in_frame * 4;
  return run_host(lut->device, host_in, host_out, n, in_frame * 4, out_frame * 4,
                  [&](const void* di, void* dout_, int m, cudaStream_t st) {
                    return mpvp_ravu_lite_launch(lut, key, radius, ar, ar_strength, static_cast<const float*>(di),
                                                 static_cast<float*>(dout_), m, h, w, (int64_t)in_frame, w, (int64_t)out_frame, 2 * w,
                                                 nullptr, st);
                  }, "mpvp_ravu_lite_host");
}

extern "C" int mpvp_ravu_host(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                              const void* host_in, void* host_out, int n, int h, int w, const mpvp_io* io) {
  MPVP_REQUIRE(lut && lut->kind == 0, "lut handle is null or not a LUT");
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad frame geometry n=%d h=%d w=%d", n, h, w);
  const int c = key_mode == MPVP_KEY_LUMA ? 1 : 3;
  const int64_t ip = (int64_t)h * w, op = ip * 4;
  return run_host(lut->device, host_in, host_out, n, (size_t)c * ip * io_bytes(io, false), (size_t)c * op * io_bytes(io, true),
                  [&](const void* di, void* dout_, int m, cudaStream_t st) {
                    return mpvp_ravu_launch_io(lut, key, radius, key_mode, di, dout_, m, h, w, c * ip, ip, w, c * op, op, 2 * w, nullptr, io, st);
                  }, "mpvp_ravu_host");
}

extern "C" int mpvp_ravu3x_host(const mpvp_weights* lut, const mpvp_key_params* key, int radius, int key_mode,
                                const void* host_in, void* host_out, int n, int h, int w, const mpvp_io* io) {
  MPVP_REQUIRE(lut && lut->kind == 0, "lut handle is null or not a LUT");
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad frame geometry n=%d h=%d w=%d", n, h, w);
  const int c = key_mode == MPVP_KEY_LUMA ? 1 : 3;
  const int64_t ip = (int64_t)h * w, op = ip * 9;
  return run_host(lut->device, host_in, host_out, n, (size_t)c * ip * io_bytes(io, false), (size_t)c * op * io_bytes(io, true),
                  [&](const void* di, void* dout_, int m, cudaStream_t st) {
                    return mpvp_ravu3x_launch_io(lut, key, radius, key_mode, di, dout_, m, h, w, c * ip, ip, w, c * op, op, 3 * w, nullptr, io, st);
                  }, "mpvp_ravu3x_host");
}

extern "C" int mpvp_ravu_zoom_host(const mpvp_weights* lut, const mpvp_weights* lut_ar, const mpvp_key_params* key, int radius,
                                   int key_mode, float ar_strength, const void* host_in, void* host_out, int n, int h, int w,
                                   int out_h, int out_w, const mpvp_io* io) {
  MPVP_REQUIRE(lut && lut->kind == 0, "lut handle is null or not a LUT");
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1 && out_h >= 1 && out_w >= 1, "bad geometry");
  const int c = key_mode == MPVP_KEY_LUMA ? 1 : 3;
  const int64_t ip = (int64_t)h * w, op = (int64_t)out_h * out_w;
  return run_host(lut->device, host_in, host_out, n, (size_t)c * ip * io_bytes(io, false), (size_t)c * op * io_bytes(io, true),
                  [&](const void* di, void* dout_, int m, cudaStream_t st) {
                    return mpvp_ravu_zoom_launch_io(lut, lut_ar, key, radius, key_mode, ar_strength, di, dout_, m, h, w, out_h, out_w,
                                                    c * ip, ip, w, c * op, op, out_w, nullptr, io, st);
                  }, "mpvp_ravu_zoom_host");
}

// nn_y / nn_x: the weights of double_y / double_x; either may be null (per-axis //!WHEN, nnedi3-nns16-win8x4.hook:19,109)
extern "C" int mpvp_nnedi3_host(const mpvp_weights* nn_y, const mpvp_weights* nn_x, const void* host_in, void* host_out, int n,
                                int h, int w, const mpvp_io* io) {
  MPVP_REQUIRE(nn_y || nn_x, "both weight sets are null");
  MPVP_REQUIRE(!nn_y || nn_y->kind == 1, "nn_y is not an NNEDI3 weight set");
  MPVP_REQUIRE(!nn_x || nn_x->kind == 1, "nn_x is not an NNEDI3 weight set");
  MPVP_REQUIRE(!(nn_y && nn_x) || nn_y->device == nn_x->device, "the two weight sets live on different devices");
  MPVP_REQUIRE(n >= 0 && h >= 1 && w >= 1, "bad frame geometry n=%d h=%d w=%d", n, h, w);
  const int device = (nn_y ? nn_y : nn_x)->device;
  const int64_t ip = (int64_t)h * w;
  const int h2 = nn_y ? 2 * h : h, w2 = nn_x ? 2 * w : w;
  const int64_t op = (int64_t)h2 * w2;
  // the W x 2H image between the two passes stays float32 on the device (one scratch buffer per stream slot)
  mpvp_io io_first{MPVP_FMT_F32, MPVP_FMT_F32, 1.f, 1.f}, io_last = io_first;
  if (io) {
    io_first.in_format = io->in_format; io_first.in_max = io->in_max;
    io_last.out_format = io->out_format; io_last.out_max = io->out_max;
  }
  float* mid[2] = {nullptr, nullptr};
  int slot = 0, chunk_cap = 0;
  int rc = run_host(device, host_in, host_out, n, (size_t)ip * io_bytes(io, false), (size_t)op * io_bytes(io, true),
                    [&](const void* di, void* dout_, int m, cudaStream_t st) -> int {
                      if (!(nn_y && nn_x)) {
                        mpvp_io one = io_first;
                        one.out_format = io_last.out_format; one.out_max = io_last.out_max;
                        return nn_y ? mpvp_nnedi3_launch_io(nn_y, 0, di, dout_, m, h, w, ip, w, op, w, &one, st)
                                    : mpvp_nnedi3_launch_io(nn_x, 1, di, dout_, m, h, w, ip, w, op, 2 * w, &one, st);
                      }
                      if (m > chunk_cap) chunk_cap = m;   // the first chunk is the largest
                      const int b = slot++ & 1;
                      if (!mid[b]) {
                        DeviceGuard g2(device);
                        if (cudaMalloc(&mid[b], sizeof(float) * (size_t)chunk_cap * 2 * ip) != cudaSuccess) {
                          set_error("mpvp_nnedi3_host: out of device memory");
                          return MPVP_E_NOMEM;
                        }
                      }
                      int r = mpvp_nnedi3_launch_io(nn_y, 0, di, mid[b], m, h, w, ip, w, 2 * ip, w, &io_first, st);
                      if (r != MPVP_OK) return r;
                      return mpvp_nnedi3_launch_io(nn_x, 1, mid[b], dout_, m, 2 * h, w, 2 * ip, w, op, 2 * w, &io_last, st);
                    }, "mpvp_nnedi3_host");
  for (int b = 0; b < 2; ++b)
    if (mid[b]) {
      DeviceGuard g2(device);
      cudaFree(mid[b]);
    }
  return rc;
}
