// libmpvp: error channel, weight handles (LUT upload, NNEDI3 repacking), small shared host helpers.
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "nnedi3_pack.cuh"

namespace mpvp {

static thread_local char t_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_grid_limit{0};

int check_fast_key(const mpvp_key_params* key) {
  MPVP_REQUIRE(key, "key params are null");
  MPVP_REQUIRE(key->n_strength >= 2 && key->n_strength <= 9, "key params: n_strength %d out of range", key->n_strength);
  MPVP_REQUIRE(key->n_l1_thr == key->n_strength - 1,
               "key params: l1_thr[] holds %d entries, expected n_strength - 1 = %d (derived tables missing: call "
               "mpvp_key_params_finalize() after filling the shader constants)", key->n_l1_thr, key->n_strength - 1);
  float prev = 0.0f;
  for (int i = 0; i < key->n_l1_thr; ++i) {
    MPVP_REQUIRE(key->l1_thr[i] > prev, "key params: l1_thr[%d] = %g is not positive and increasing (call mpvp_key_params_finalize())",
                 i, (double)key->l1_thr[i]);
    prev = key->l1_thr[i];
  }
  MPVP_REQUIRE(key->coh_ratio[0] > 1.0f && key->coh_ratio[1] > key->coh_ratio[0],
               "key params: coh_ratio = (%g, %g) must be increasing and > 1 (call mpvp_key_params_finalize())",
               (double)key->coh_ratio[0], (double)key->coh_ratio[1]);
  return MPVP_OK;
}

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
extern "C" int mpvp_debug_set_grid_limit(int max_ctas) {
  const int prev = g_grid_limit.exchange(max_ctas > 0 ? max_ctas : 0, std::memory_order_relaxed);
  return prev;
}

// The shader's strength expression as a function of the eigenvalue L1, float32 op for op (ravu-lite-ar-r3.hook:86,
// ravu-r2.hook:97): lambda = sqrt(L1); thresholds form: #{t : lambda >= t}; log2 form:
// clamp(floor(log2(lambda * scale + eps)), 0, n_strength - 1).  sqrt via double is correctly rounded for float32; the
// log2 is taken in double and rounded to float32 once (a correctly rounded float32 log2).
static int strength_of_l1(const mpvp_key_params* k, float x) {
  const float lam = (float)sqrt((double)x);
  if (k->n_strength_thr > 0) {
    int s = 0;
    for (int i = 0; i < k->n_strength_thr; ++i) s += lam >= k->strength_thr[i] ? 1 : 0;
    return s;
  }
  volatile float prod = lam * k->strength_log2_scale;   // volatile: keep the two float32 roundings of the shader
  volatile float arg = prod + 1.192092896e-7f;
  if (!(arg > 0.0f)) return 0;
  const float v = floorf((float)log2((double)arg));
  const float hi = (float)(k->n_strength - 1);
  return (int)(v < 0.0f ? 0.0f : (v > hi ? hi : v));
}

extern "C" int mpvp_key_params_finalize(mpvp_key_params* key) {
  MPVP_REQUIRE(key, "key params are null");
  MPVP_REQUIRE(key->n_strength >= 2 && key->n_strength <= 9, "n_strength %d out of range [2, 9]", key->n_strength);
  MPVP_REQUIRE(key->n_strength_thr == 0 || key->n_strength_thr == key->n_strength - 1,
               "n_strength_thr %d does not match n_strength %d", key->n_strength_thr, key->n_strength);
  MPVP_REQUIRE(key->n_strength_thr > 0 || key->strength_log2_scale > 0.0f, "log2-form strength needs a positive scale");
  for (int c = 0; c < 2; ++c)
    MPVP_REQUIRE(key->coherence_thr[c] > 0.0f && key->coherence_thr[c] < 1.0f, "coherence threshold %g not in (0, 1)",
                 (double)key->coherence_thr[c]);
  // for each level k >= 1: the smallest positive float32 L1 whose strength is >= k (the quantiser is monotone in L1),
  // by bisection over the float32 bit patterns
  for (int lvl = 1; lvl < key->n_strength; ++lvl) {
    uint32_t lo = 0, hi = 0x7F7FFFFFu;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      float x;
      memcpy(&x, &mid, 4);
      if (strength_of_l1(key, x) >= lvl) hi = mid; else lo = mid + 1;
    }
    memcpy(&key->l1_thr[lvl - 1], &lo, 4);
  }
  key->n_l1_thr = key->n_strength - 1;
  for (int c = 0; c < 2; ++c) {
    const double t = key->coherence_thr[c], r = (1.0 + t) / (1.0 - t);
    key->coh_ratio[c] = (float)(r * r);
  }
  return check_fast_key(key);
}

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
  if (W->zoom_plans && W->zoom_plans_free) W->zoom_plans_free(W->zoom_plans);
  if (W->lut_tex) cudaDestroyTextureObject(W->lut_tex);
  if (W->lut_arr) cudaFreeArray(W->lut_arr);
  if (W->nn_b) cudaFree(W->nn_b);
  if (W->nn_bias) cudaFree(W->nn_bias);
  if (W->nn_w) cudaFree(W->nn_w);
  delete W;
  return MPVP_OK;
}
