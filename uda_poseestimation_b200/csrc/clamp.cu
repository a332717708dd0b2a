// clamp.cu — per-channel clamp of the stylised images back into the normalised pixel range.
//
// Replaces the inline expression the trainers apply to every style-transfer output
// (train_human.py:276,351,356; train_animal.py:301,376,381 of the reference):
//     x = torch.maximum(torch.minimum(x.permute(0,2,3,1), recover_max), recover_min).permute(0,3,1,2)
// which costs two permuted (strided) elementwise passes.  Here: one contiguous NCHW pass,
// 128-bit loads/stores, the bound pair of the plane's channel held in registers.  torch's
// minimum/maximum propagate NaN from either operand; so does this kernel.
#include "common.cuh"

namespace udape {

constexpr int kClampThreads = 256;
constexpr int kClampUnroll = 4;

__device__ __forceinline__ float clamp_nan(float x, float lo, float hi) {
    // torch.minimum(x, hi): NaN if either is NaN, else the smaller; then torch.maximum(., lo)
    float m = (x != x || hi != hi) ? __int_as_float(0x7fc00000) : fminf(x, hi);
    return (m != m || lo != lo) ? __int_as_float(0x7fc00000) : fmaxf(m, lo);
}

// grid: (chunks per plane, planes); every CTA owns a contiguous chunk of one (n,c) plane
template <typename T, bool VEC>
__global__ void __launch_bounds__(kClampThreads)
channel_clamp_kernel(const T* __restrict__ x, T* __restrict__ out, const float* __restrict__ lo,
                     const float* __restrict__ hi, int channels, int64_t hw) {
    const int64_t plane = blockIdx.y;
    const int c = static_cast<int>(plane % channels);
    const float l = __ldg(lo + c), h = __ldg(hi + c);
    const T* p = x + plane * hw;
    T* o = out + plane * hw;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const int64_t nvec = hw / EPV;
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
        uint4* o4 = reinterpret_cast<uint4*>(o);
        const int64_t base = static_cast<int64_t>(blockIdx.x) * (kClampThreads * kClampUnroll);
        uint4 v[kClampUnroll];
#pragma unroll
        for (int u = 0; u < kClampUnroll; ++u) {
            const int64_t i = base + u * kClampThreads + threadIdx.x;
            if (i < nvec) v[u] = ldg_stream(p4 + i);
        }
#pragma unroll
        for (int u = 0; u < kClampUnroll; ++u) {
            const int64_t i = base + u * kClampThreads + threadIdx.x;
            if (i < nvec) {
                float f[EPV];
                unpack16<T>(v[u], f);
#pragma unroll
                for (int e = 0; e < EPV; ++e) f[e] = clamp_nan(f[e], l, h);
                stg_stream(o4 + i, pack16<T>(f));
            }
        }
    } else {
        const int64_t base = static_cast<int64_t>(blockIdx.x) * (kClampThreads * kClampUnroll);
        for (int u = 0; u < kClampUnroll; ++u) {
            const int64_t i = base + u * kClampThreads + threadIdx.x;
            if (i < hw) o[i] = from_f32<T>(clamp_nan(to_f32<T>(p[i]), l, h));
        }
    }
}

}  // namespace udape

using namespace udape;

extern "C" int udape_channel_clamp(const void* x, int dtype, int64_t planes, int64_t channels, int64_t hw,
                                   const float* lo, const float* hi, void* out, void* stream) {
    UDAPE_REQUIRE(x && out && lo && hi, UDAPE_ERR_NULL, "udape_channel_clamp: NULL pointer");
    UDAPE_REQUIRE(planes > 0 && channels > 0 && hw > 0 && planes < 65536 * channels && hw < (1ll << 40) &&
                      planes % channels == 0,
                  UDAPE_ERR_SHAPE, "udape_channel_clamp: bad extents planes=%lld channels=%lld hw=%lld",
                  (long long)planes, (long long)channels, (long long)hw);
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(es == 2 || es == 4, UDAPE_ERR_DTYPE, "udape_channel_clamp: unsupported dtype code %d", dtype);
    UDAPE_REQUIRE(aligned_to(x, es) && aligned_to(out, es) && aligned_to(lo, 4) && aligned_to(hi, 4), UDAPE_ERR_ALIGN,
                  "udape_channel_clamp: misaligned pointer");
    UDAPE_REQUIRE(planes <= 65535, UDAPE_ERR_SHAPE, "udape_channel_clamp: more than 65535 planes (N*C)");
    cudaStream_t st = as_stream(stream);
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        constexpr int EPV = Vec16<T>::EPV;
        const bool vec = aligned16(x) && aligned16(out) && (hw % EPV) == 0;
        const int64_t items = vec ? hw / EPV : hw;
        const int64_t per_cta = kClampThreads * kClampUnroll;
        const dim3 grid(static_cast<unsigned>((items + per_cta - 1) / per_cta), static_cast<unsigned>(planes));
        if (vec) channel_clamp_kernel<T, true><<<grid, kClampThreads, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(out), lo, hi, static_cast<int>(channels), hw);
        else channel_clamp_kernel<T, false><<<grid, kClampThreads, 0, st>>>(static_cast<const T*>(x), static_cast<T*>(out), lo, hi, static_cast<int>(channels), hw);
    });
    return check_launch("udape_channel_clamp");
}
