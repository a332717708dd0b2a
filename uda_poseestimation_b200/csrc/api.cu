// api.cu — library-level entry points and the error plumbing shared by every op.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "pipeline.cuh"

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

// ---- persistent-grid sizing for the TMA-staged kernels (pipeline.cuh) ------------------------
int sm_count_of_current_device() {
    static int cached[64] = {0};  // benign race: every thread writes the same value
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int pipe_grid(int64_t planes) {
    const int64_t sms = sm_count_of_current_device();
    // one persistent CTA per SM; with fewer planes than SMs, one plane per CTA
    const int64_t g = planes < sms ? planes : sms;
    return static_cast<int>(g > 0 ? g : 1);
}

int pipe_stage_cap() {
    const char* e = std::getenv("UDAPE_PIPE_STAGES");
    const int v = e ? std::atoi(e) : 0;
    return (v > 0 && v <= kMaxStages) ? v : kMaxStages;
}

bool pipe_enabled() {
    // read on every call so that tests can compare both code paths within one process
    const char* e = std::getenv("UDAPE_NO_TMA");
    return !(e && e[0] == '1');
}

// out[0..cols) = table[(*counter % rows)][0..cols);  *counter += 1   (one warp)
__global__ void table_feed_kernel(const float* __restrict__ table, int rows, int cols, uint32_t* __restrict__ counter,
                                  float* __restrict__ out) {
    const uint32_t r = *counter % static_cast<uint32_t>(rows);
    for (int i = threadIdx.x; i < cols; i += 32) out[i] = table[static_cast<size_t>(r) * cols + i];
    __syncwarp();
    if (threadIdx.x == 0) *counter = *counter + 1u;
}

}  // namespace udape

extern "C" {

int udape_table_feed(const float* table, int rows, int cols, uint32_t* counter, float* out, void* stream) {
    UDAPE_REQUIRE(table && counter && out, UDAPE_ERR_NULL, "udape_table_feed: NULL argument");
    UDAPE_REQUIRE(rows > 0 && cols > 0 && cols <= 1024, UDAPE_ERR_SHAPE, "udape_table_feed: rows=%d cols=%d", rows, cols);
    udape::table_feed_kernel<<<1, 32, 0, udape::as_stream(stream)>>>(table, rows, cols, counter, out);
    return udape::check_launch("udape_table_feed");
}

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
