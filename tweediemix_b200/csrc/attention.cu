// placeholder until the tcgen05 kernel lands (next commit)
#include "tmx_common.cuh"
namespace tmx { int attn_init() { return TMX_OK; } }
extern "C" int tmx_attn_fwd(const void*, const void*, const void*, void*, int, int, int, int, int,
                            int64_t, int64_t, int64_t, int64_t, float, int, void*) {
    tmx::set_error("tmx_attn_fwd: not built yet");
    return TMX_ESHAPE;
}
