// tcgen05 (5th-gen tensor core) mean-shift iteration -- placeholder until the kernel lands.
#include "internal.h"
namespace sed {
int ms_shift_tc(const float*, const float*, int, int, int, int, int, int, float*, float*, cudaStream_t) {
    return SED_ERR_UNSUPPORTED;
}
}  // namespace sed
