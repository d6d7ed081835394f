// Instantiations of the fast feature kernel (features_fast.cuh): n_fft 1024, int16 PCM input.
#include "features_fast.cuh"

namespace seld {
SELD_FAST_UNIT_DEFINE(r32_i16, 32, true, false)
}  // namespace seld
