// Host-side repacking of NNEDI3 weights for the tcgen05 kernel (nnedi3_tc.cu).
//
// B operand (K-major, SWIZZLE_NONE canonical layout).  Accumulator column (= B row) order: groups of 32
// columns hold 16 neurons, first their 16 softmax logits (W1 * log2(e)), then their 16 elliott inputs
// (W2), i.e. row n' = 32*(n/16) + 16*which + n%16 -- one tcgen05.ld.x32 then yields both halves with
// neighbouring neurons in neighbouring registers, which is what the packed f32x2 epilogue wants.  Rows are
// stored as [(K+16)/8][2*nns][8] binary16, i.e. element (n', k) at ((k/8) * 2*nns + n') * 8 + k%8.
// In UMMA descriptor terms: core matrix = 8 rows x 16 B contiguous (SBO = 128 B between 8-row groups),
// LBO = 2*nns*16 B between K chunks of 8 elements.  log2(e) is folded into W1 and b1 so that the
// epilogue's exp() is a bare ex2.  The biases ride in the GEMM: K is extended by one 16-wide step whose
// A columns are (1, 1, 1, 0, ...) and whose B rows carry bias = hi + mid + lo as three binary16 terms
// (exact to ~2^-33 relative), so the accumulator already holds W.x*inv_std + b.
#pragma once
#include <cuda_fp16.h>

#include <cstring>
#include <vector>

namespace mpvp {

// w1, w2: [nns][K]; outputs: packed B operand bytes, bias [2*nns] = (b1*log2e, b2) interleaved,
// fp32 weights [2*nns][K] (unscaled, for the CUDA-core path)
// gn = neurons per accumulator block: 16 (32-column blocks, the layout described above) or 8 (16-column blocks:
// row n' = 16*(n/8) + 8*which + n%8, for the pipelined kernel whose register staging is tcgen05.ld.x16)
inline void nnedi3_pack_host(const float* w1, const float* w2, const float* b1, const float* b2, int nns, int K, int gn,
                             std::vector<unsigned char>& packed, std::vector<float>& bias, std::vector<float>& wf) {
  const float kLog2e = 1.4426950408889634f;
  const int N = 2 * nns;
  const int KX = K + 16;
  std::vector<__half> hb((size_t)N * KX, __float2half_rn(0.f));
  bias.resize(2 * (size_t)N);  // [0, N): tensor path (b1*log2e, b2); [N, 2N): CUDA-core path (b1, b2)
  wf.resize((size_t)N * K);
  for (int n = 0; n < nns; ++n) {
    const int r1 = 2 * gn * (n / gn) + (n % gn), r2 = r1 + gn;  // tensor-path rows of (W1_n, W2_n)
    bias[r1] = b1[n] * kLog2e;
    bias[r2] = b2[n];
    bias[N + 2 * n] = b1[n];
    bias[N + 2 * n + 1] = b2[n];
    for (int k = 0; k < K; ++k) {
      const float a = w1[(size_t)n * K + k], b = w2[(size_t)n * K + k];
      wf[(size_t)(2 * n) * K + k] = a;
      wf[(size_t)(2 * n + 1) * K + k] = b;
      hb[((size_t)(k / 8) * N + r1) * 8 + (k % 8)] = __float2half_rn(a * kLog2e);
      hb[((size_t)(k / 8) * N + r2) * 8 + (k % 8)] = __float2half_rn(b);
    }
  }
  for (int r = 0; r < N; ++r) {
    float rem = bias[r];
    for (int t = 0; t < 3; ++t) {
      const __half h = __float2half_rn(rem);
      hb[((size_t)(K / 8) * N + r) * 8 + t] = h;
      rem -= __half2float(h);
    }
  }
  packed.resize(hb.size() * sizeof(__half));
  memcpy(packed.data(), hb.data(), packed.size());
}

}  // namespace mpvp
