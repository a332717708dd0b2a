// api.cu — library-level entry points and the error plumbing shared by every op.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace udape {

static thread_local char g_last_error[512] = {0};

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        return fail(static_cast<int>(e), "%s: %s (%s)", what, cudaGetErrorName(e),
                    cudaGetErrorString(e));
    }
    return UDAPE_OK;
}

}  // namespace udape

extern "C" {

int udape_version(void) { return UDAPE_VERSION; }

const char* udape_build_info(void) {
#define UDAPE_STR2(x) #x
#define UDAPE_STR(x) UDAPE_STR2(x)
    return "udape-b200 " UDAPE_STR(UDAPE_VERSION) " sm_100a cuda " UDAPE_STR(CUDART_VERSION);
}

int udape_last_error(char* buf, size_t buf_bytes) {
    size_t n = strlen(udape::g_last_error);
    if (buf && buf_bytes) {
        size_t m = n < buf_bytes - 1 ? n : buf_bytes - 1;
        memcpy(buf, udape::g_last_error, m);
        buf[m] = 0;
    }
    return static_cast<int>(n);
}

}  // extern "C"
