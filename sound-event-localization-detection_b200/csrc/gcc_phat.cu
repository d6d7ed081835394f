// K3 placeholder until the GCC-PHAT kernel lands (next commit).
#include "seld_common.h"
namespace seld {
int launch_gcc(const seld_plan*, const FeatArgs&, cudaStream_t) {
    set_error("GCC-PHAT mode not built yet");
    return SELD_ERR_UNSUPPORTED;
}
}  // namespace seld
