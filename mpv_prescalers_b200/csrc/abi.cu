// libmpvp: error channel, weight handles (LUT upload, NNEDI3 repacking), small shared host helpers.
#include <cuda_fp16.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "nnedi3_pack.cuh"

namespace mpvp {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int sm_count(int device) {
  static std::mutex mu;
  static int cache[64];
  static bool have[64];
  std::lock_guard<std::mutex> lk(mu);
  if (device >= 0 && device < 64 && have[device]) return cache[device];
  int v = 148;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) v = 148;
  if (device >= 0 && device < 64) {
    cache[device] = v;
    have[device] = true;
  }
  return v;
}

}  // namespace mpvp

using namespace mpvp;

extern "C" const char* mpvp_last_error(void) { return t_err; }
extern "C" int mpvp_abi_version(void) { return MPVP_ABI_VERSION; }
extern "C" uint64_t mpvp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int mpvp_weights_create_lut(int device, const float* host, int w, int h, int round_to_fp16,
                                       mpvp_weights** out) {
  MPVP_REQUIRE(out, "out is null");
  *out = nullptr;
  MPVP_REQUIRE(host && w > 0 && h > 0, "bad LUT arguments");
  DeviceGuard guard(device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", device);
  const size_t count = (size_t)w * h * 4;
  std::vector<float> tmp(host, host + count);
  if (round_to_fp16) {
    // rgba16f storage: round-to-nearest-even to binary16
    for (size_t i = 0; i < count; ++i) tmp[i] = __half2float(__float2half_rn(tmp[i]));
  }
  mpvp_weights* W = new (std::nothrow) mpvp_weights();
  if (!W) {
    set_error("out of host memory");
    return MPVP_E_NOMEM;
  }
  W->device = device;
  W->kind = 0;
  W->lut_w = w;
  W->lut_h = h;
  cudaError_t e = cudaMalloc(&W->lut, count * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(W->lut, tmp.data(), count * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && round_to_fp16) {
    std::vector<__half> hv(count);
    for (size_t i = 0; i < count; ++i) hv[i] = __float2half_rn(tmp[i]);
    e = cudaMalloc(&W->lut_half, count * sizeof(__half));
    if (e == cudaSuccess) e = cudaMemcpy(W->lut_half, hv.data(), count * sizeof(__half), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && round_to_fp16 && h == 288 * 9) {
    // ravu-zoom LUT (//!FILTER LINEAR, ravu-zoom-r2.hook:135-139): a texture object lets the texture unit do the
    // bilinear blend the GL sampler does in the reference
    const cudaChannelFormatDesc cd = cudaCreateChannelDescHalf4();
    e = cudaMallocArray(&W->lut_arr, &cd, (size_t)w, (size_t)h);
    if (e == cudaSuccess)
      e = cudaMemcpy2DToArray(W->lut_arr, 0, 0, W->lut_half, (size_t)w * 8, (size_t)w * 8, (size_t)h, cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) {
      cudaResourceDesc rd{};
      rd.resType = cudaResourceTypeArray;
      rd.res.array.array = W->lut_arr;
      cudaTextureDesc td{};
      td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
      td.filterMode = cudaFilterModeLinear;
      td.readMode = cudaReadModeElementType;
      td.normalizedCoords = 0;
      e = cudaCreateTextureObject(&W->lut_tex, &rd, &td, nullptr);
    }
  }
  if (e != cudaSuccess) {
    set_error("LUT upload failed: %s", cudaGetErrorString(e));
    mpvp_weights_destroy(W);
    return MPVP_E_CUDA;
  }
  *out = W;
  return MPVP_OK;
}

namespace mpvp {
int nnedi3_group_size(int nns);  // nnedi3_tc.cu: which kernel (hence which B row order) serves this nns
}

extern "C" int mpvp_weights_create_nnedi3(int device, const float* w1, const float* w2, const float* b1,
                                          const float* b2, int nns, int win_short, mpvp_weights** out) {
  MPVP_REQUIRE(out, "out is null");
  *out = nullptr;
  MPVP_REQUIRE(w1 && w2 && b1 && b2, "null weight pointer");
  MPVP_REQUIRE(nns == 16 || nns == 32 || nns == 64 || nns == 128 || nns == 256, "nns %d unsupported", nns);
  MPVP_REQUIRE(win_short == 4 || win_short == 6, "window 8x%d unsupported", win_short);
  DeviceGuard guard(device);
  MPVP_REQUIRE(guard.ok, "cannot switch to device %d", device);
  mpvp_weights* W = new (std::nothrow) mpvp_weights();
  if (!W) {
    set_error("out of host memory");
    return MPVP_E_NOMEM;
  }
  W->device = device;
  W->kind = 1;
  W->nns = nns;
  W->win_short = win_short;
  const int K = 8 * win_short;
  std::vector<unsigned char> packed;
  std::vector<float> bias, wf;
  W->nn_group = nnedi3_group_size(nns);
  nnedi3_pack_host(w1, w2, b1, b2, nns, K, W->nn_group, packed, bias, wf);
  W->nn_b_bytes = packed.size();
  cudaError_t e = cudaMalloc(&W->nn_b, packed.size());
  if (e == cudaSuccess) e = cudaMemcpy(W->nn_b, packed.data(), packed.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&W->nn_bias, bias.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(W->nn_bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&W->nn_w, wf.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(W->nn_w, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("NNEDI3 weight upload failed: %s", cudaGetErrorString(e));
    mpvp_weights_destroy(W);
    return MPVP_E_CUDA;
  }
  *out = W;
  return MPVP_OK;
}

extern "C" int mpvp_weights_destroy(mpvp_weights* W) {
  if (!W) return MPVP_OK;
  DeviceGuard guard(W->device);
  if (W->lut) cudaFree(W->lut);
  if (W->lut_half) cudaFree(W->lut_half);
  if (W->lut_tex) cudaDestroyTextureObject(W->lut_tex);
  if (W->lut_arr) cudaFreeArray(W->lut_arr);
  if (W->nn_b) cudaFree(W->nn_b);
  if (W->nn_bias) cudaFree(W->nn_bias);
  if (W->nn_w) cudaFree(W->nn_w);
  delete W;
  return MPVP_OK;
}
