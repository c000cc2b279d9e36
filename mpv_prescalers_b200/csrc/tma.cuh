#pragma once
// TMA (cp.async.bulk.tensor) staging of halo tiles, shared by the RAVU / RAVU-Zoom / NNEDI3 kernels.
//
// A tile + halo is fetched by ONE elected thread as a box of a tensor map laid over the input planes
// ({w, h, planes}); completion is signalled on an mbarrier.  Two hardware rules shape the callers:
//   * the innermost box coordinate must be a multiple of 16 bytes and the box row a multiple of 16 bytes (an x origin of
//     x0 - halo is rejected as an illegal instruction, tools/tma_min.cu): boxes start at x0 - XO with XO a multiple of
//     16 bytes >= halo, and are as wide as the next multiple of 16 bytes;
//   * out-of-image texels are ZERO-filled, the reference's samplers CLAMP to the edge: tiles that touch the image border
//     are patched in shared memory after arrival (patch_clamp_to_edge), tile-uniform and rare.
// Planes that break the 16-byte rules (odd pitches, unaligned views) make make_plane_tmap() return false and the kernels
// fall back to plain clamped loads.  MPVP_TMA=0 forces that path (A/B switch).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace mpvp {
namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init1(uint32_t addr) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_wait_parity(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(addr),
      "r"(parity)
      : "memory");
}

// true for exactly one lane of a fully converged warp (elect.sync)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tma_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// box of a 3-D tensor map {x, y, plane} -> shared memory, completion on `bar` (coordinates may be negative: OOB -> 0)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, uint64_t tmap_ptr, int x, int y, int p, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(tmap_ptr), "r"(x), "r"(y), "r"(p), "r"(bar)
      : "memory");
}
// the buffer a finished tile lived in may be refilled by the async proxy: order the generic-proxy reads before it
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Replicate the image border over the zero-filled out-of-image texels of a staged tile [rows][pitch] whose texel (0, 0) is
// image texel (gx0, gy0).  Call with all `nthreads` threads of the group that owns the tile; `sync` is its barrier.
template <class SyncF>
__device__ __forceinline__ void patch_clamp_to_edge(float* __restrict__ tile, int pitch, int cols, int rows, int gx0, int gy0, int w, int h,
                                                    int tid, int nthreads, SyncF sync) {
  for (int i = tid; i < cols * rows; i += nthreads) {
    const int sy = i / cols, sx = i - sy * cols;
    const int gx = gx0 + sx, gy = gy0 + sy;
    if (gy >= 0 && gy < h && (gx < 0 || gx >= w)) tile[sy * pitch + sx] = tile[sy * pitch + clampi(gx, 0, w - 1) - gx0];
  }
  sync();
  for (int i = tid; i < cols * rows; i += nthreads) {
    const int sy = i / cols, sx = i - sy * cols;
    const int gy = gy0 + sy;
    if (gy < 0 || gy >= h) tile[sy * pitch + sx] = tile[(clampi(gy, 0, h - 1) - gy0) * pitch + sx];
  }
  sync();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// MPVP_TMA=0 forces the plain-load staging path (A/B switch)
inline bool tma_enabled() {
  static const bool v = [] {
    const char* e = getenv("MPVP_TMA");
    return !(e && e[0] == '0');
  }();
  return v;
}

// 3-D tensor map {w, h, n} over the input planes with a (box_w x box_h x 1) box; false if the layout does not
// meet TMA's 16-byte rules (then the kernel stages with plain loads).  eb = bytes per element (4, 2 or 1); sy / sn =
// row / plane pitch in elements.
inline bool make_plane_tmap(CUtensorMap* tm, const void* base, int eb, int w, int h, int n, int64_t sy, int64_t sn, int box_w, int box_h) {
  if (!tma_enabled() || !encode_tiled()) return false;
  if (n == 1) sn = (int64_t)h * sy;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (sy * eb) % 16 || (sn * eb) % 16 || sy < w || box_w > 256 || box_h > 256 ||
      (box_w * eb) % 16 || sn <= 0)
    return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  const cuuint64_t strides[2] = {(cuuint64_t)sy * eb, (cuuint64_t)sn * eb};
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = eb == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  return encode_tiled()(tm, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace mpvp
