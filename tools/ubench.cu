// Pipe-throughput microbenchmarks for B200 (sm_100a): what does one SM sustain per clock for the
// instruction kinds the RAVU kernels are made of?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
// Output: one line per test, "name threads/block ops/clk/SM".
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITERS 2048

template <int KIND>
__global__ void k(float* out, long long* cyc, float a0, float b0) {
  float a = a0 + threadIdx.x * 1e-6f, b = b0;
  float r[16];
  float2 q[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = a + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = make_float2(a + i, b + i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (KIND == 0) {  // scalar FFMA, 3 distinct register sources
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
    } else if (KIND == 1) {  // packed FFMA2
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = __ffma2_rn(q[i], a2, b2);
    } else if (KIND == 2) {  // scalar FMUL
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __fmul_rn(r[i], a);
    } else if (KIND == 3) {  // scalar FADD
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __fadd_rn(r[i], a);
    } else if (KIND == 4) {  // packed FMUL2
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = __fmul2_rn(q[i], a2);
    } else if (KIND == 5) {  // packed FADD2
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = __fadd2_rn(q[i], a2);
    } else if (KIND == 6) {  // FMNMX
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = fmaxf(r[i], a + i);
    } else if (KIND == 7) {  // MUFU.RCP
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __frcp_rn(r[i]) , r[i] = r[i];
    } else if (KIND == 8) {  // MUFU.EX2
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = exp2f(r[i]) ;
    } else if (KIND == 9) {  // FFMA with two accumulators sharing b (reuse-friendly): r = r*a + r' pattern
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = fmaf(a, b, r[i]);
    } else if (KIND == 10) {  // IEEE division
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __fdiv_rn(r[i], a);
    } else if (KIND == 11) {  // IEEE sqrt
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __fsqrt_rn(r[i] + 1.0f);
    } else if (KIND == 12) {  // atan2f
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = atan2f(r[i], a);
    } else if (KIND == 13) {  // approximate division
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = __fdividef(r[i], a);
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += q[i].x + q[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// LDS.128 with per-lane pseudo-random rows (the LUT access pattern): bytes/clk/SM
__global__ void lds_k(float* out, long long* cyc, int stride4, int rows, int mode) {
  extern __shared__ float4 s[];
  for (int i = threadIdx.x; i < rows * stride4; i += blockDim.x) s[i] = make_float4(i, i + 1, i + 2, i + 3);
  __syncthreads();
  unsigned h = threadIdx.x * 2654435761u + 12345u;
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    int row;
    if (mode == 0) row = (h >> 8) % rows;                    // random row per lane
    else if (mode == 1) row = ((threadIdx.x / 8) * 7 + it) % rows;  // same row for 8 neighbouring lanes
    else row = it % rows;                                      // broadcast
    h = h * 1664525u + 1013904223u;
#pragma unroll
    for (int t = 0; t < 13; ++t) {
      float4 v = s[row * stride4 + t];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, int ops_per_iter, int threads, int blocks_per_sm) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int blocks = sms * blocks_per_sm;
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  cudaMalloc(&cyc, sizeof(long long) * blocks);
  k<KIND><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
  k<KIND><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
  cudaDeviceSynchronize();
  long long* h = (long long*)malloc(sizeof(long long) * blocks);
  cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; ++i) avg += h[i];
  avg /= blocks;
  double ops = (double)ops_per_iter * ITERS * threads * blocks_per_sm;
  printf("%-28s threads=%4d x%d  %8.1f lane-ops/clk/SM  (err=%s)\n", name, threads, blocks_per_sm, ops / avg,
         cudaGetErrorString(cudaGetLastError()));
  free(h); cudaFree(out); cudaFree(cyc);
}

void run_lds(const char* name, int mode, int threads) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int rows = 288, stride4 = 13;
  size_t smem = sizeof(float4) * rows * stride4;
  cudaFuncSetAttribute(lds_k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * threads);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  lds_k<<<sms, threads, smem>>>(out, cyc, stride4, rows, mode);
  lds_k<<<sms, threads, smem>>>(out, cyc, stride4, rows, mode);
  cudaDeviceSynchronize();
  long long* h = (long long*)malloc(sizeof(long long) * sms);
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += h[i];
  avg /= sms;
  double bytes = 16.0 * 13 * ITERS * threads;
  printf("%-28s threads=%4d      %8.1f B/clk/SM  (err=%s)\n", name, threads, bytes / avg, cudaGetErrorString(cudaGetLastError()));
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {256, 1024}) {
    run<0>("FFMA r=r*a+b", 16, threads, 1);
    run<9>("FFMA r=a*b+r", 16, threads, 1);
    run<1>("FFMA2 (x2 lanes counted)", 16, threads, 1);
    run<2>("FMUL", 16, threads, 1);
    run<3>("FADD", 16, threads, 1);
    run<4>("FMUL2", 16, threads, 1);
    run<5>("FADD2", 16, threads, 1);
    run<6>("FMNMX", 16, threads, 1);
    run<7>("RCP (IEEE)", 16, threads, 1);
    run<8>("EX2", 16, threads, 1);
    run<10>("FDIV IEEE", 16, threads, 1);
    run<13>("FDIV approx", 16, threads, 1);
    run<11>("FSQRT IEEE", 16, threads, 1);
    run<12>("atan2f", 4, threads, 1);
  }
  run<0>("FFMA 2 blocks/SM", 16, 512, 2);
  run<1>("FFMA2 2 blocks/SM", 16, 512, 2);
  for (int threads : {256, 512}) {
    run_lds("LDS.128 random rows", 0, threads);
    run_lds("LDS.128 8-lane groups", 1, threads);
    run_lds("LDS.128 broadcast", 2, threads);
  }
  return 0;
}
