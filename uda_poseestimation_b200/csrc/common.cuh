// common.cuh — shared device helpers for the udape sm_100a kernels.
//
// All hot-path kernels are HBM-bound streaming reductions / elementwise passes
// (SURVEY.md §8d).  The building blocks here are:
//   * 128-bit streaming loads/stores (ld.global.nc.L1::no_allocate / st.global.cs),
//   * an order-preserving float -> uint32 key so that argmax with numpy/torch
//     semantics (first index wins, NaN is the maximum) becomes an exact u64 max,
//   * warp-shuffle + shared-memory block reductions with a fixed (deterministic) tree.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/udape.h"

namespace udape {

// ---- host-side error plumbing (api.cu) -------------------------------------------
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);
int sm_count_of_current_device();  // api.cu (cached per device)
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned_to(const void* p, size_t a) {
    return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0;
}
inline int dtype_size(int dtype) {
    switch (dtype) {
        case UDAPE_F32: return 4;
        case UDAPE_F16: return 2;
        case UDAPE_BF16: return 2;
        case UDAPE_U8: return 1;
        default: return 0;
    }
}

#define UDAPE_REQUIRE(cond, code, ...)                 \
    do {                                               \
        if (!(cond)) return ::udape::fail((code), __VA_ARGS__); \
    } while (0)

// Dispatch a floating dtype code to a template type.
#define UDAPE_DISPATCH_FLOAT(code, T, ...)                                   \
    do {                                                                     \
        switch (code) {                                                      \
            case UDAPE_F32: { using T = float; __VA_ARGS__; } break;         \
            case UDAPE_F16: { using T = __half; __VA_ARGS__; } break;        \
            case UDAPE_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
            default: return ::udape::fail(UDAPE_ERR_DTYPE, "unsupported dtype code %d", (int)(code)); \
        }                                                                    \
    } while (0)

// ---- element conversion -------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- 128-bit streaming access ---------------------------------------------------------
// Read-once inputs go through the non-coherent path without allocating in L1;
// write-once outputs use the streaming (evict-first) store policy.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
// L2 evict-first variants for the EMA, the biggest streaming pass of the step: its bytes are touched once,
// and marking them as the first victims keeps the small tensors the heatmap chains hand from kernel to
// kernel (re-warped maps, gradients, the inverse plan: a few MB each) resident in the 126 MB L2.
// Measured inside the step (tools/step_probe.py): 181.0 -> 178.8 us.  The same hint on the AdaIN loads made
// the step slower (183.5 us), plain instead of streaming stores for the chain's outputs changed nothing.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_stream_ef(const void* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ uint4 ldg_ef(const void* p, uint64_t pol) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol)
                 : "memory");
    return r;
}
__device__ __forceinline__ void stg_ef(void* p, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}
// Cached variants (data that is re-read by the same CTA, or read-modify-write).
__device__ __forceinline__ uint4 ldg_cached(const void* p) {
    return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void stg_plain(void* p, const uint4& v) {
    *reinterpret_cast<uint4*>(p) = v;
}

// A 16-byte vector holds EPV elements of T.
template <typename T> struct Vec16 { static constexpr int EPV = 16 / sizeof(T); };

// unpack a 16-byte vector of T into EPV floats
template <typename T> __device__ __forceinline__ void unpack16(const uint4& v, float* f);
template <> __device__ __forceinline__ void unpack16<float>(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
}
template <> __device__ __forceinline__ void unpack16<__half>(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        float2 t = __half22float2(h);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
template <> __device__ __forceinline__ void unpack16<__nv_bfloat16>(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        // bf16 -> f32 is a 16-bit shift
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// pack EPV floats into a 16-byte vector of T (round-to-nearest-even)
template <typename T> __device__ __forceinline__ uint4 pack16(const float* f);
template <> __device__ __forceinline__ uint4 pack16<float>(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack16<__half>(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <> __device__ __forceinline__ uint4 pack16<__nv_bfloat16>(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// 32-bit streaming load (lane-consecutive words: still 128 bytes per warp)
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
// 64-bit streaming access (the narrow operand of a mixed-dtype pair)
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream8(void* p, const uint2& v) {
    asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// Pack<T, G>: G consecutive elements of T owned by one thread, G*sizeof(T) in {8, 16}
// bytes.  Mixed-dtype kernels (fp16 student heatmap vs fp32 label) give every thread the
// same G elements of both operands, G = 16 / sizeof(wider type): the wide operand moves
// as 128-bit vectors, the narrow one as 64-bit vectors, and consecutive lanes always
// touch consecutive addresses (no strided half-sector requests).
template <typename T, int G> struct Pack {
    static constexpr int BYTES = G * static_cast<int>(sizeof(T));
    static_assert(BYTES == 8 || BYTES == 16, "Pack must be 8 or 16 bytes");
    uint4 raw;
    __device__ __forceinline__ void load(const T* p) {
        if constexpr (BYTES == 16) raw = ldg_stream(p);
        else { const uint2 t = ldg_stream8(p); raw = make_uint4(t.x, t.y, 0u, 0u); }
    }
    // same slice out of a shared-memory staged plane
    __device__ __forceinline__ void load_shared(const void* p) {
        if constexpr (BYTES == 16) raw = *reinterpret_cast<const uint4*>(p);
        else { const uint2 t = *reinterpret_cast<const uint2*>(p); raw = make_uint4(t.x, t.y, 0u, 0u); }
    }
    __device__ __forceinline__ void get(float* f) const {
        if constexpr (BYTES == 16) unpack16<T>(raw, f);
        else {  // 4 x 16-bit
            float t[8];
            unpack16<T>(raw, t);
#pragma unroll
            for (int i = 0; i < G; ++i) f[i] = t[i];
        }
    }
    static __device__ __forceinline__ void store(T* p, const float* f) {
        if constexpr (BYTES == 16) stg_stream(p, pack16<T>(f));
        else {
            float t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = i < G ? f[i] : 0.0f;
            const uint4 v = pack16<T>(t);
            stg_stream8(p, make_uint2(v.x, v.y));
        }
    }
};
template <typename TA, typename TB> struct PairGroup {
    static constexpr int G = 16 / static_cast<int>(sizeof(TA) > sizeof(TB) ? sizeof(TA) : sizeof(TB));
};

// ---- ordered key for exact arg-reduction ------------------------------------------------
// key(x) is monotone in x for non-NaN x, maps -0.0 and +0.0 to the same key (they compare
// equal, so the first index must win) and maps every NaN to the top key (numpy/torch
// argmax treat NaN as the maximum and return the first one).
__device__ __forceinline__ uint32_t order_key(float x) {
    if (x != x) return 0xffffffffu;
    uint32_t u = __float_as_uint(x + 0.0f);  // -0.0 + 0.0 = +0.0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t key) {
    if (key == 0xffffffffu) return __int_as_float(0x7fc00000);
    uint32_t u = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    return __uint_as_float(u);
}
// pack (key, index): larger key wins; for equal keys the smaller index wins.
__device__ __forceinline__ unsigned long long pack_arg(uint32_t key, uint32_t idx) {
    return (static_cast<unsigned long long>(key) << 32) | static_cast<unsigned long long>(~idx);
}
__device__ __forceinline__ uint32_t arg_key(unsigned long long v) { return static_cast<uint32_t>(v >> 32); }
__device__ __forceinline__ uint32_t arg_idx(unsigned long long v) { return ~static_cast<uint32_t>(v); }

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

// Block-wide sum with a fixed tree (deterministic for a given block size).  `red` is
// >= 32 floats of shared memory.  Result is valid in every thread.
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
    constexpr int WARPS = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` against the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < WARPS) ? red[lane] : 0.0f;
    return warp_sum(t);
}
template <int THREADS>
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v,
                                                            unsigned long long* red) {
    constexpr int WARPS = THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max_u64(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    unsigned long long t = (lane < WARPS) ? red[lane] : 0ull;
    return warp_max_u64(t);
}


// Runtime-size variants (any blockDim.x that is a multiple of 32, <= 1024).
__device__ __forceinline__ float block_sum_rt(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` against the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (lane < warps) ? red[lane] : 0.0f;
    return warp_sum(t);
}

// ---- last-CTA-done ticket ------------------------------------------------------------------------
// Deterministic grid-wide reductions: every CTA publishes its partials, takes a ticket, and the CTA
// that draws the last one reduces all partials in a fixed order.  `ticket` is a caller-owned uint32
// that must be ZERO on entry; atomicInc wraps it back to zero with the last ticket, so the same word
// can be reused by the next launch without a memset (no extra graph node, no extra launch).
// Call from all threads of the CTA, after the CTA's partials have been written by thread 0 / lane 0s
// and a __syncthreads().  Returns true in every thread of the last CTA.
__device__ __forceinline__ bool last_block_done(uint32_t* ticket, unsigned total) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        __threadfence();  // publish this CTA's partials before taking a ticket
        const unsigned t = atomicInc(ticket, total - 1u);
        is_last = (t == total - 1u);
        if (is_last) __threadfence();  // acquire the other CTAs' partials
    }
    __syncthreads();
    return is_last;
}
// Fixed-order sum of n floats by one CTA (any block size): thread t adds v[t], v[t + T], ... in that order, then the
// block tree.  The values were written by OTHER CTAs of the grid (published by the ticket's fences), so they are read
// from L2 (ld.global.cg) — sixteen loads of a thread in flight at a time: as a volatile loop the tail of the LAST CTA
// was a chain of dependent L2 round trips (42 per thread at C5, ~8 us of the fused loss step's 59) while every other
// SM had already drained.  Same order of additions as before, same bits.
// (not inlined: its sixteen load registers must not raise the register count of the streaming path around it)
static __device__ __noinline__ float cta_sum_array(const float* v, int64_t n, float* red) {
    constexpr int UN = 16;
    const int64_t stride = blockDim.x;
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += UN * stride) {
        float x[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) x[u] = (i + u * stride < n) ? __ldcg(v + i + u * stride) : 0.0f;
#pragma unroll
        for (int u = 0; u < UN; ++u)
            if (i + u * stride < n) s += x[u];
    }
    return block_sum_rt(s, red);
}


}  // namespace udape
