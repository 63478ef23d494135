// Library-wide state of libdprnn_b200: the last-error slot and the build descriptor.
#include "common.cuh"
#include "../../include/dprnn_b200.h"
#include <cstdarg>

namespace dprnn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace dprnn

extern "C" const char* dprnn_last_error(void) { return dprnn::g_err; }

extern "C" const char* dprnn_build_info(void) {
    return "libdprnn_b200;arch=sm_100a;modes=fp32"
           ",bf16-tcgen05"
           ";cuda=" DPRNN_STR(CUDART_VERSION);
}
