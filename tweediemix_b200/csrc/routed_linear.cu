// placeholder until the grouped projection kernel lands
#include "tmx_common.cuh"
extern "C" int tmx_routed_linear_fwd(const void*, const void* const*, const void* const*, const void* const*,
                                     void*, int, int, int, int, int, int, void*) {
    tmx::set_error("tmx_routed_linear_fwd: not built yet");
    return TMX_ESHAPE;
}
