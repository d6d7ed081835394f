// Instantiations of the fast feature kernel (features_fast.cuh): n_fft 1024, float32 input.
#include "features_fast.cuh"

namespace seld {
SELD_FAST_UNIT_DEFINE(r32_f32, 32, false, true)
}  // namespace seld
