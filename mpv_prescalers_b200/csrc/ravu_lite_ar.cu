// RAVU-Lite-AR kernels (anti-ringing variants of the 2x family), instantiated apart from ravu_lite.cu so that the
// translation units compile in parallel (kernel: ravu_lite_kernel.cuh).
#include "ravu_lite_kernel.cuh"

namespace mpvp {

int ravu_lite_ar_dispatch(const LiteArgs& a, int radius, int device, cudaStream_t st) {
  switch (radius) {
    case 2: return launch_lite<2, true, 2, 4, 2>(a, device, st);
    case 3: return launch_lite<3, true, 2, MPVP_X_AR3_P, MPVP_X_AR3_STRIPS>(a, device, st);  // 64x40 tiles: LUT + tiles + power tile = 103 KB, 2 CTAs/SM
    case 4: return launch_lite<4, true, 2, 2, MPVP_X_AR4_STRIPS>(a, device, st);
  }
  return MPVP_E_INVALID;
}

}  // namespace mpvp
