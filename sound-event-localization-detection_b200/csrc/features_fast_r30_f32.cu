// Instantiations of the fast feature kernel (features_fast.cuh): n_fft 960, float32 input.
#include "features_fast.cuh"

namespace seld {
SELD_FAST_UNIT_DEFINE(r30_f32, 30, false, true)
}  // namespace seld
