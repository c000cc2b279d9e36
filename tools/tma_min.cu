// Minimal TMA (cp.async.bulk.tensor.3d) probe: which box / coordinate combinations does sm_100a accept?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int bw, int bh, int cx, int cy, int cz) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  float* tile = reinterpret_cast<float*>(smem);
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(tile);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bw * bh * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(cx), "r"(cy), "r"(cz), "r"(bar_a)
                 : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(bar_a),
      "r"(0)
      : "memory");
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  const int W = 240, H = 135, N = 2;
  std::vector<float> h((size_t)W * H * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100000);
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 256 * 256 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  struct Case { int bw, bh, cx, cy, cz; } cases[] = {{64, 32, 0, 0, 0}, {64, 32, 64, 32, 1}, {68, 36, 0, 0, 0}, {68, 36, 62, 30, 0}, {68, 36, -2, -2, 0}, {68, 36, 190, 126, 1}, {72, 38, -3, -3, 1}};
  for (auto c : cases) {
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    cuuint64_t dims[3] = {W, H, N};
    cuuint64_t strides[2] = {W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {(cuuint32_t)c.bw, (cuuint32_t)c.bh, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaMemset(o, 0xff, 256 * 256 * 4);
    probe<<<1, 128, c.bw * c.bh * 4 + 128>>>(tm, o, c.bw, c.bh, c.cx, c.cy, c.cz);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> r0(c.bw * c.bh);
    cudaMemcpy(r0.data(), o, r0.size() * 4, cudaMemcpyDeviceToHost);
    // expected value at tile (sx, sy): in-bounds -> h[(cz*H + gy)*W + gx], else 0
    int bad = 0;
    for (int sy = 0; sy < c.bh; ++sy)
      for (int sx = 0; sx < c.bw; ++sx) {
        int gx = c.cx + sx, gy = c.cy + sy;
        float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[((size_t)c.cz * H + gy) * W + gx] : 0.f;
        if (r0[sy * c.bw + sx] != want) ++bad;
      }
    printf("box %dx%d at (%d,%d,%d): encode=%d run=%s mismatches=%d\n", c.bw, c.bh, c.cx, c.cy, c.cz, (int)r, cudaGetErrorString(e), bad);
    if (e != cudaSuccess) break;
  }
  return 0;
}
