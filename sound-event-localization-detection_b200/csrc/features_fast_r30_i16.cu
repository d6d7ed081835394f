// Instantiations of the fast feature kernel (features_fast.cuh): n_fft 960, int16 PCM input.
#include "features_fast.cuh"

namespace seld {
SELD_FAST_UNIT_DEFINE(r30_i16, 30, true, false)
}  // namespace seld
