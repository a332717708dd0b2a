// ema.cu — multi-tensor EMA teacher update (and buffer copy).
//
// Replaces utils.py:9-25 (OldWeightEMA.step: per-tensor mul_ + temp + add_, 969 launches
// and 7 passes over 212 MB for PoseResNet-101) and lib/models/ema.py:18-44 (ModelEMA).
// One launch walks a device-resident chunk table {dst, src, numel}; each CTA owns one
// chunk (a contiguous run of <= chunk_elems elements of one parameter tensor) and
// streams it with 128-bit loads: algorithmic traffic 2 reads + 1 write per element.
//
// Arithmetic is  fl(fl(dst*a) + fl(src*b))  — the same three fp32 roundings as the
// reference's  p.mul_(alpha); p.add_(src * (1-alpha))  — so fp32 results are bit-identical
// to the eager reference (no FMA contraction).
#include <cstdlib>

#include "common.cuh"

namespace udape {

constexpr int kEmaThreads = 256;
constexpr int kEmaUnroll = 4;

__device__ __forceinline__ float ema_op(float p, float s, float a, float b) {
    return __fadd_rn(__fmul_rn(p, a), __fmul_rn(s, b));
}

template <typename T, bool COPY>
__global__ void __launch_bounds__(kEmaThreads)
ema_multi_kernel(const udape_ema_chunk* __restrict__ chunks, float a, float b, int evict_first) {
    constexpr int EPV = Vec16<T>::EPV;
    const uint64_t pol = l2_evict_first_policy();
    const udape_ema_chunk c = chunks[blockIdx.x];
    T* __restrict__ dst = static_cast<T*>(c.dst);
    const T* __restrict__ src = static_cast<const T*>(c.src);
    const int n = static_cast<int>(c.numel);
    if (aligned16(dst) && aligned16(src)) {
        const int nvec = n / EPV;
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        for (int base = 0; base < nvec; base += kEmaThreads * kEmaUnroll) {
            uint4 pv[kEmaUnroll], sv[kEmaUnroll];
#pragma unroll
            for (int u = 0; u < kEmaUnroll; ++u) {
                const int i = base + u * kEmaThreads + threadIdx.x;
                if (i < nvec) {
                    if (evict_first) {
                        sv[u] = ldg_stream_ef(s4 + i, pol);
                        if (!COPY) pv[u] = ldg_ef(d4 + i, pol);
                    } else {
                        sv[u] = ldg_stream(s4 + i);
                        if (!COPY) pv[u] = ldg_cached(d4 + i);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kEmaUnroll; ++u) {
                const int i = base + u * kEmaThreads + threadIdx.x;
                if (i < nvec) {
                    if (COPY) {
                        if (evict_first) stg_ef(d4 + i, sv[u], pol); else stg_plain(d4 + i, sv[u]);
                    } else {
                        float fp[EPV], fs[EPV];
                        unpack16<T>(pv[u], fp);
                        unpack16<T>(sv[u], fs);
#pragma unroll
                        for (int e = 0; e < EPV; ++e) fp[e] = ema_op(fp[e], fs[e], a, b);
                        if (evict_first) stg_ef(d4 + i, pack16<T>(fp), pol); else stg_plain(d4 + i, pack16<T>(fp));
                    }
                }
            }
        }
        for (int i = nvec * EPV + threadIdx.x; i < n; i += kEmaThreads)
            dst[i] = COPY ? src[i] : from_f32<T>(ema_op(to_f32<T>(dst[i]), to_f32<T>(src[i]), a, b));
    } else {
        for (int i = threadIdx.x; i < n; i += kEmaThreads)
            dst[i] = COPY ? src[i] : from_f32<T>(ema_op(to_f32<T>(dst[i]), to_f32<T>(src[i]), a, b));
    }
}

// byte copy (buffers of arbitrary dtype)
__global__ void __launch_bounds__(kEmaThreads)
copy_multi_kernel(const udape_ema_chunk* __restrict__ chunks) {
    const udape_ema_chunk c = chunks[blockIdx.x];
    uint8_t* dst = static_cast<uint8_t*>(c.dst);
    const uint8_t* src = static_cast<const uint8_t*>(c.src);
    const int n = static_cast<int>(c.numel);
    int done = 0;
    if (aligned16(dst) && aligned16(src)) {
        const int nvec = n >> 4;
        for (int i = threadIdx.x; i < nvec; i += kEmaThreads)
            reinterpret_cast<uint4*>(dst)[i] = ldg_stream(reinterpret_cast<const uint4*>(src) + i);
        done = nvec << 4;
    }
    for (int i = done + threadIdx.x; i < n; i += kEmaThreads) dst[i] = src[i];
}

}  // namespace udape

using namespace udape;

extern "C" int64_t udape_ema_plan(void* const* dst, const void* const* src, const int64_t* numel,
                                  int64_t n_tensors, int64_t elem_bytes, int64_t chunk_elems,
                                  udape_ema_chunk* out, int64_t capacity) {
    if (!dst || !src || !numel) return fail(UDAPE_ERR_NULL, "udape_ema_plan: NULL table");
    if (elem_bytes != 1 && elem_bytes != 2 && elem_bytes != 4)
        return fail(UDAPE_ERR_DTYPE, "udape_ema_plan: elem_bytes must be 1, 2 or 4");
    // chunk boundaries stay 16-byte aligned relative to the tensor base
    if (n_tensors < 0 || chunk_elems <= 0 || chunk_elems >= (1ll << 31) || (chunk_elems % 16) != 0)
        return fail(UDAPE_ERR_ARG, "udape_ema_plan: chunk_elems must be a positive multiple of 16 below 2^31");
    int64_t n = 0;
    for (int64_t t = 0; t < n_tensors; ++t) {
        if (numel[t] < 0) return fail(UDAPE_ERR_SHAPE, "udape_ema_plan: tensor %lld has negative numel", (long long)t);
        if (numel[t] > 0 && (!dst[t] || !src[t]))
            return fail(UDAPE_ERR_NULL, "udape_ema_plan: tensor %lld has a NULL pointer", (long long)t);
        for (int64_t off = 0; off < numel[t]; off += chunk_elems, ++n) {
            if (out && n < capacity) {
                const int64_t rem = numel[t] - off;
                out[n].dst = static_cast<char*>(dst[t]) + off * elem_bytes;
                out[n].src = static_cast<const char*>(src[t]) + off * elem_bytes;
                out[n].numel = rem < chunk_elems ? rem : chunk_elems;
            }
        }
    }
    return n;
}

extern "C" int udape_ema_multi(const udape_ema_chunk* chunks_dev, int64_t n_chunks, int64_t chunk_elems,
                               float a, float b, int dtype, int mode, void* stream) {
    if (n_chunks == 0) return UDAPE_OK;
    UDAPE_REQUIRE(chunks_dev, UDAPE_ERR_NULL, "udape_ema_multi: chunk table is NULL");
    UDAPE_REQUIRE(n_chunks > 0 && n_chunks < (1ll << 31), UDAPE_ERR_SHAPE, "udape_ema_multi: bad n_chunks=%lld", (long long)n_chunks);
    UDAPE_REQUIRE(chunk_elems > 0 && chunk_elems < (1ll << 31), UDAPE_ERR_ARG, "udape_ema_multi: bad chunk_elems");
    UDAPE_REQUIRE(mode == 0 || mode == 1, UDAPE_ERR_ARG, "udape_ema_multi: mode must be 0 (ema) or 1 (copy)");
    cudaStream_t st = as_stream(stream);
    const unsigned grid = static_cast<unsigned>(n_chunks);
    if (mode == 1 && dtype == UDAPE_U8) {
        copy_multi_kernel<<<grid, kEmaThreads, 0, st>>>(chunks_dev);
        return check_launch("udape_ema_multi");
    }
    // L2 evict-first for the EMA's once-touched bytes (common.cuh); UDAPE_EMA_EVICT_FIRST=0 turns the hint off
    const char* e_ef = std::getenv("UDAPE_EMA_EVICT_FIRST");
    const int ef = e_ef ? std::atoi(e_ef) : 1;
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        if (mode == 0) ema_multi_kernel<T, false><<<grid, kEmaThreads, 0, st>>>(chunks_dev, a, b, ef);
        else ema_multi_kernel<T, true><<<grid, kEmaThreads, 0, st>>>(chunks_dev, a, b, ef);
    });
    return check_launch("udape_ema_multi");
}
