// Library-level entry points of liblctgan_sm100.so (see include/lctgan.h).
#include "common.cuh"

int g_lct_kernel_launches = 0;

LCT_API int lct_version(void) { return 1; }
LCT_API int lct_kernel_launches(void) { return g_lct_kernel_launches; }
LCT_API int lct_reset_kernel_launches(void) {
    g_lct_kernel_launches = 0;
    return 0;
}
