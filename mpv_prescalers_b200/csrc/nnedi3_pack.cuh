// Host-side repacking of NNEDI3 weights (placeholder layout until the tcgen05 kernel lands).
#pragma once
#include <vector>

namespace mpvp {

// w1, w2: [nns][K]; outputs: packed B operand bytes, bias [2*nns], fp32 weights [2*nns][K]
inline void nnedi3_pack_host(const float* w1, const float* w2, const float* b1, const float* b2, int nns, int K,
                             std::vector<unsigned char>& packed, std::vector<float>& bias, std::vector<float>& wf) {
  packed.assign(16, 0);
  bias.resize(2 * (size_t)nns);
  wf.resize(2 * (size_t)nns * K);
  for (int n = 0; n < nns; ++n) {
    bias[2 * n] = b1[n];
    bias[2 * n + 1] = b2[n];
    for (int k = 0; k < K; ++k) {
      wf[(size_t)(2 * n) * K + k] = w1[(size_t)n * K + k];
      wf[(size_t)(2 * n + 1) * K + k] = w2[(size_t)n * K + k];
    }
  }
}

}  // namespace mpvp
