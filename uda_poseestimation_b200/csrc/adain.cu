// adain.cu — per-(n,c) channel statistics and fused AdaIN (+ s2t/t2s alpha mix).
//
// Replaces adain/function.py:3-22 and lib/models/Style_net.py:4-29,167-168 of the
// reference.  The eager reference makes ~19 passes over the feature tensors
// (SURVEY.md §2.1); here one launch reads content and style once and writes the
// result once: algorithmic bytes = planes*(hw_c+hw_s)*E read + planes*hw_c*E write.
//
// Two code paths:
//   * warp-per-plane (planes of <= 256 16-byte vectors, i.e. the 32x32 relu4_1
//     planes of the trainers): each lane issues all of its 128-bit loads for BOTH
//     planes up front (up to 16 outstanding per lane), keeps the plane in registers,
//     and does an exact two-pass mean / sum((x-mean)^2) with warp shuffles.  No shared
//     memory, no block barrier.
//   * block-per-plane (any plane size / alignment): three sweeps over the plane, the
//     2nd and 3rd served by L1/L2 (a plane is at most a few hundred KB), so DRAM still
//     sees one read.
#include <cstdlib>

#include "common.cuh"

namespace udape {

constexpr int kWarpsPerBlock = 8;
constexpr int kCtaThreads = 256;

// ------------------------------------------------------------------------------------
// warp-per-plane register path
// ------------------------------------------------------------------------------------
template <typename T, int J>
__device__ __forceinline__ void warp_load_plane(const T* plane, int nvec, int lane, uint4 (&v)[J]) {
    const uint4* p = reinterpret_cast<const uint4*>(plane);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int i = lane + 32 * j;
        v[j] = (i < nvec) ? ldg_stream(p + i) : make_uint4(0, 0, 0, 0);
    }
}

// exact two-pass statistics of a register-resident plane
template <typename T, int J>
__device__ __forceinline__ void warp_plane_stats(const uint4 (&v)[J], int nvec, int hw, int lane,
                                                 float eps, float& mean, float& stdv) {
    constexpr int EPV = Vec16<T>::EPV;
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        if (lane + 32 * j < nvec) {
            float f[EPV];
            unpack16<T>(v[j], f);
            float t = 0.0f;
#pragma unroll
            for (int e = 0; e < EPV; ++e) t += f[e];
            s += t;
        }
    }
    s = warp_sum(s);
    mean = s / static_cast<float>(hw);
    float m2 = 0.0f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
        if (lane + 32 * j < nvec) {
            float f[EPV];
            unpack16<T>(v[j], f);
            float t = 0.0f;
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const float d = f[e] - mean;
                t = fmaf(d, d, t);
            }
            m2 += t;
        }
    }
    m2 = warp_sum(m2);
    // unbiased variance (torch .var default); hw == 1 -> 0/0 = NaN like the reference
    const float var = m2 / static_cast<float>(hw - 1);
    stdv = sqrtf(var + eps);
}

template <typename T, int J>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
mean_std_warp_kernel(const T* __restrict__ feat, T* __restrict__ mean_out, T* __restrict__ std_out,
                     int64_t planes, int nvec, int hw, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t plane = static_cast<int64_t>(blockIdx.x) * kWarpsPerBlock + (threadIdx.x >> 5);
    if (plane >= planes) return;
    uint4 v[J];
    warp_load_plane<T, J>(feat + plane * hw, nvec, lane, v);
    float mean, stdv;
    warp_plane_stats<T, J>(v, nvec, hw, lane, eps, mean, stdv);
    if (lane == 0) {
        mean_out[plane] = from_f32<T>(mean);
        std_out[plane] = from_f32<T>(stdv);
    }
}

// up to UDAPE_MAX_ADAIN_JOBS independent (content, style, alpha) -> out jobs of one shape in ONE launch
// (blockIdx.y = job): the s2t and t2s directions of a train step (train_human.py:348-356) are independent, and a
// launch of this size spends ~10 % of its time ramping up and draining — paid once instead of twice.
struct AdainJobs {
    const void* content[UDAPE_MAX_ADAIN_JOBS];
    const void* style[UDAPE_MAX_ADAIN_JOBS];
    void* out[UDAPE_MAX_ADAIN_JOBS];
    const float* alpha_dev[UDAPE_MAX_ADAIN_JOBS];
    float alpha[UDAPE_MAX_ADAIN_JOBS];
    int mix[UDAPE_MAX_ADAIN_JOBS];
};

template <typename T, int J, int WPB>
__global__ void __launch_bounds__(WPB * 32, 24 / WPB)
adain_warp_kernel(const AdainJobs jobs, int64_t planes, int nvec_c, int nvec_s, int hw_c, int hw_s, float eps) {
    constexpr int EPV = Vec16<T>::EPV;
    const int lane = threadIdx.x & 31;
    const int64_t plane = static_cast<int64_t>(blockIdx.x) * WPB + (threadIdx.x >> 5);
    if (plane >= planes) return;
    const int job = blockIdx.y;
    const T* content = static_cast<const T*>(jobs.content[job]);
    const T* style = static_cast<const T*>(jobs.style[job]);
    T* out = static_cast<T*>(jobs.out[job]);
    const float* alpha_dev = jobs.alpha_dev[job];
    const int mix = jobs.mix[job];

    // issue every load of both planes before the first use
    uint4 cv[J], sv[J];
    warp_load_plane<T, J>(content + plane * hw_c, nvec_c, lane, cv);
    warp_load_plane<T, J>(style + plane * hw_s, nvec_s, lane, sv);
    const float a = alpha_dev ? __ldg(alpha_dev) : jobs.alpha[job];

    float mean_s, std_s, mean_c, std_c;
    warp_plane_stats<T, J>(sv, nvec_s, hw_s, lane, eps, mean_s, std_s);
    warp_plane_stats<T, J>(cv, nvec_c, hw_c, lane, eps, mean_c, std_c);

    const float inv_c = 1.0f / std_c;
    const float one_minus_a = 1.0f - a;
    uint4* o4 = reinterpret_cast<uint4*>(out + plane * hw_c);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int i = lane + 32 * j;
        if (i < nvec_c) {
            float f[EPV];
            unpack16<T>(cv[j], f);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const float n = (f[e] - mean_c) * inv_c;
                const float t = fmaf(n, std_s, mean_s);
                f[e] = mix ? fmaf(a, t, one_minus_a * f[e]) : t;
            }
            stg_stream(o4 + i, pack16<T>(f));
        }
    }
}

// ------------------------------------------------------------------------------------
// block-per-plane generic path
// ------------------------------------------------------------------------------------
template <typename T, bool VEC>
__device__ __forceinline__ float cta_plane_sum(const T* p, int64_t hw) {
    float s = 0.0f;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const int64_t nvec = hw / EPV;
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
        for (int64_t i = threadIdx.x; i < nvec; i += kCtaThreads) {
            float f[EPV];
            unpack16<T>(ldg_cached(p4 + i), f);
#pragma unroll
            for (int e = 0; e < EPV; ++e) s += f[e];
        }
    } else {
        for (int64_t i = threadIdx.x; i < hw; i += kCtaThreads) s += to_f32<T>(p[i]);
    }
    return s;
}
template <typename T, bool VEC>
__device__ __forceinline__ float cta_plane_m2(const T* p, int64_t hw, float mean) {
    float s = 0.0f;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const int64_t nvec = hw / EPV;
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
        for (int64_t i = threadIdx.x; i < nvec; i += kCtaThreads) {
            float f[EPV];
            unpack16<T>(ldg_cached(p4 + i), f);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const float d = f[e] - mean;
                s = fmaf(d, d, s);
            }
        }
    } else {
        for (int64_t i = threadIdx.x; i < hw; i += kCtaThreads) {
            const float d = to_f32<T>(p[i]) - mean;
            s = fmaf(d, d, s);
        }
    }
    return s;
}

template <typename T, bool VEC>
__device__ __forceinline__ void cta_plane_stats(const T* p, int64_t hw, float eps, float* red,
                                                float& mean, float& stdv) {
    const float s = block_sum<kCtaThreads>(cta_plane_sum<T, VEC>(p, hw), red);
    mean = s / static_cast<float>(hw);
    const float m2 = block_sum<kCtaThreads>(cta_plane_m2<T, VEC>(p, hw, mean), red);
    stdv = sqrtf(m2 / static_cast<float>(hw - 1) + eps);
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kCtaThreads)
mean_std_cta_kernel(const T* __restrict__ feat, T* __restrict__ mean_out, T* __restrict__ std_out,
                    int64_t hw, float eps) {
    __shared__ float red[32];
    const int64_t plane = blockIdx.x;
    float mean, stdv;
    cta_plane_stats<T, VEC>(feat + plane * hw, hw, eps, red, mean, stdv);
    if (threadIdx.x == 0) {
        mean_out[plane] = from_f32<T>(mean);
        std_out[plane] = from_f32<T>(stdv);
    }
}

template <typename T, bool VEC_C, bool VEC_S>
__global__ void __launch_bounds__(kCtaThreads)
adain_cta_kernel(const T* __restrict__ content, const T* __restrict__ style, T* __restrict__ out,
                 int64_t hw_c, int64_t hw_s, float eps, float alpha,
                 const float* __restrict__ alpha_dev, int mix) {
    __shared__ float red[32];
    const int64_t plane = blockIdx.x;
    const T* c = content + plane * hw_c;
    const T* s = style + plane * hw_s;
    T* o = out + plane * hw_c;
    float mean_s, std_s, mean_c, std_c;
    cta_plane_stats<T, VEC_S>(s, hw_s, eps, red, mean_s, std_s);
    cta_plane_stats<T, VEC_C>(c, hw_c, eps, red, mean_c, std_c);
    const float a = alpha_dev ? __ldg(alpha_dev) : alpha;
    const float inv_c = 1.0f / std_c;
    const float one_minus_a = 1.0f - a;
    if (VEC_C) {
        constexpr int EPV = Vec16<T>::EPV;
        const int64_t nvec = hw_c / EPV;
        const uint4* c4 = reinterpret_cast<const uint4*>(c);
        uint4* o4 = reinterpret_cast<uint4*>(o);
        for (int64_t i = threadIdx.x; i < nvec; i += kCtaThreads) {
            float f[EPV];
            unpack16<T>(ldg_cached(c4 + i), f);
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                const float n = (f[e] - mean_c) * inv_c;
                const float t = fmaf(n, std_s, mean_s);
                f[e] = mix ? fmaf(a, t, one_minus_a * f[e]) : t;
            }
            stg_stream(o4 + i, pack16<T>(f));
        }
    } else {
        for (int64_t i = threadIdx.x; i < hw_c; i += kCtaThreads) {
            const float x = to_f32<T>(c[i]);
            const float n = (x - mean_c) * inv_c;
            const float t = fmaf(n, std_s, mean_s);
            o[i] = from_f32<T>(mix ? fmaf(a, t, one_minus_a * x) : t);
        }
    }
}

// ------------------------------------------------------------------------------------
// single-pass streaming statistics for large planes (relu1_1 .. relu3_1 of the AdaIN decoder
// pre-training job: 256x256 / 128x128 / 64x64 planes, adain/net.py:137-143)
// ------------------------------------------------------------------------------------
// A 256 KB plane fits neither a warp's registers nor (with 1000+ CTAs resident) the L2, so the
// three-sweep kernel above would read it from DRAM up to three times.  Here every element is read
// ONCE: a thread takes its vectors in batches of kStreamBatch (all loads issued before the first
// use), computes the batch's exact two-pass (count, mean, M2) in registers and merges it into its
// running triple with Chan's pairwise update
//     n = na + nb,  d = mb - ma,  mean = ma + d*nb/n,  M2 = M2a + M2b + d*d*na*nb/n
// — the same update then merges lanes (shuffles), warps (shared memory) and gives the plane's
// mean and unbiased variance with the accuracy of a pairwise two-pass sum.
constexpr int kStreamBatch = 8;

struct Moments { float n, mean, m2; };

__device__ __forceinline__ Moments merge_moments(const Moments& a, const Moments& b) {
    const float n = a.n + b.n;
    if (n == 0.0f) return a;
    const float d = b.mean - a.mean;
    const float rb = b.n / n;
    Moments r;
    r.n = n;
    r.mean = fmaf(d, rb, a.mean);
    r.m2 = a.m2 + b.m2 + d * d * a.n * rb;
    return r;
}

// G threads (a multiple of 32, G | kCtaThreads) share one plane; a CTA holds kCtaThreads / G planes, so
// that mid-sized planes (64x64: 1024 vectors) still give every thread a full batch of loads to issue.
template <typename T, int G>
__device__ __forceinline__ Moments group_plane_moments(const T* __restrict__ p, int64_t hw, Moments* red) {
    constexpr int EPV = Vec16<T>::EPV;
    const int64_t nvec = hw / EPV;
    const uint4* p4 = reinterpret_cast<const uint4*>(p);
    const int gt = threadIdx.x % G;   // thread within the group
    Moments run = {0.0f, 0.0f, 0.0f};
    for (int64_t base = gt; base < nvec; base += static_cast<int64_t>(G) * kStreamBatch) {
        uint4 v[kStreamBatch];
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < kStreamBatch; ++u) {
            const int64_t i = base + static_cast<int64_t>(u) * G;
            if (i < nvec) { v[u] = ldg_stream(p4 + i); ++cnt; }
        }
        float s = 0.0f;
#pragma unroll
        for (int u = 0; u < kStreamBatch; ++u) {
            if (u < cnt) {
                float f[EPV];
                unpack16<T>(v[u], f);
                float t = 0.0f;
#pragma unroll
                for (int e = 0; e < EPV; ++e) t += f[e];
                s += t;
            }
        }
        Moments b;
        b.n = static_cast<float>(cnt * EPV);
        b.mean = s / b.n;
        float m2 = 0.0f;
#pragma unroll
        for (int u = 0; u < kStreamBatch; ++u) {
            if (u < cnt) {
                float f[EPV];
                unpack16<T>(v[u], f);
                float t = 0.0f;
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    const float d = f[e] - b.mean;
                    t = fmaf(d, d, t);
                }
                m2 += t;
            }
        }
        b.m2 = m2;
        run = merge_moments(run, b);
    }
    // lanes, then the group's warps (fixed tree: deterministic)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Moments other;
        other.n = __shfl_xor_sync(0xffffffffu, run.n, o);
        other.mean = __shfl_xor_sync(0xffffffffu, run.mean, o);
        other.m2 = __shfl_xor_sync(0xffffffffu, run.m2, o);
        // both partners must compute the same result: merge in lane order
        run = (threadIdx.x & o) ? merge_moments(other, run) : merge_moments(run, other);
    }
    if (G == 32) return run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = run;
    __syncthreads();
    const int w0 = (threadIdx.x / G) * (G / 32);
    Moments tot = red[w0];
#pragma unroll
    for (int w = 1; w < G / 32; ++w) tot = merge_moments(tot, red[w0 + w]);
    return tot;
}

template <typename T, int G>
__global__ void __launch_bounds__(kCtaThreads)
mean_std_stream_kernel(const T* __restrict__ feat, T* __restrict__ mean_out, T* __restrict__ std_out,
                       int64_t planes, int64_t hw, float eps) {
    __shared__ Moments red[kCtaThreads / 32];
    int64_t plane = static_cast<int64_t>(blockIdx.x) * (kCtaThreads / G) + threadIdx.x / G;
    const bool live = plane < planes;
    if (!live) plane = planes - 1;   // keep every thread on the barrier; the duplicate result is not stored
    const Moments m = group_plane_moments<T, G>(feat + plane * hw, hw, red);
    if (live && threadIdx.x % G == 0) {
        mean_out[plane] = from_f32<T>(m.mean);
        std_out[plane] = from_f32<T>(sqrtf(m.m2 / static_cast<float>(hw - 1) + eps));
    }
}

// Backward of calc_mean_std (the style loss of adain/net.py:137-143 differentiates through it):
//   dfeat[p, i] = dmean[p] / hw + dstd[p] * (feat[p, i] - mean[p]) / ((hw - 1) * std[p])
// (std = sqrt(var_unbiased + eps)  =>  dstd/dx_i = (x_i - mean) / ((hw-1) std)).  One read, one write.
template <typename T, bool VEC>
__global__ void __launch_bounds__(kCtaThreads)
mean_std_bwd_kernel(const T* __restrict__ feat, const T* __restrict__ mean, const T* __restrict__ stdv,
                    const T* __restrict__ dmean, const T* __restrict__ dstd, T* __restrict__ dfeat,
                    int64_t hw, int splits, int64_t span) {
    const int64_t plane = blockIdx.x / splits;
    const int64_t lo = static_cast<int64_t>(blockIdx.x % splits) * span;   // span is a multiple of the vector width
    const int64_t hi = lo + span < hw ? lo + span : hw;
    const float mu = to_f32<T>(mean[plane]);
    const float a = dmean ? to_f32<T>(dmean[plane]) / static_cast<float>(hw) : 0.0f;
    const float b = dstd ? to_f32<T>(dstd[plane]) / (static_cast<float>(hw - 1) * to_f32<T>(stdv[plane])) : 0.0f;
    const T* x = feat + plane * hw;
    T* o = dfeat + plane * hw;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const uint4* x4 = reinterpret_cast<const uint4*>(x);
        uint4* o4 = reinterpret_cast<uint4*>(o);
        const int64_t v0 = lo / EPV, v1 = hi / EPV;
        for (int64_t base = v0 + threadIdx.x; base < v1; base += static_cast<int64_t>(kCtaThreads) * 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = base + static_cast<int64_t>(u) * kCtaThreads;
                if (i < v1) v[u] = ldg_stream(x4 + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = base + static_cast<int64_t>(u) * kCtaThreads;
                if (i < v1) {
                    float f[EPV];
                    unpack16<T>(v[u], f);
#pragma unroll
                    for (int e = 0; e < EPV; ++e) f[e] = fmaf(b, f[e] - mu, a);
                    stg_stream(o4 + i, pack16<T>(f));
                }
            }
        }
    } else {
        for (int64_t i = lo + threadIdx.x; i < hi; i += kCtaThreads) o[i] = from_f32<T>(fmaf(b, to_f32<T>(x[i]) - mu, a));
    }
}

// ------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------
template <typename T>
static bool plane_vectorizable(const void* base, int64_t hw) {
    return aligned16(base) && (hw % Vec16<T>::EPV) == 0;
}
static int pick_j(int64_t nvec) {  // vectors per lane needed to hold a plane in one warp
    if (nvec <= 32) return 1;
    if (nvec <= 64) return 2;
    if (nvec <= 128) return 4;
    if (nvec <= 256) return 8;
    return 0;
}

template <typename T>
static int launch_mean_std(const void* feat, int64_t planes, int64_t hw, float eps, void* mean,
                           void* stdv, cudaStream_t st) {
    const T* f = static_cast<const T*>(feat);
    T* m = static_cast<T*>(mean);
    T* s = static_cast<T*>(stdv);
    const bool vec = plane_vectorizable<T>(feat, hw);
    const int j = vec ? pick_j(hw / Vec16<T>::EPV) : 0;
    if (j) {
        const unsigned grid = static_cast<unsigned>((planes + kWarpsPerBlock - 1) / kWarpsPerBlock);
        const int nvec = static_cast<int>(hw / Vec16<T>::EPV);
        const int ihw = static_cast<int>(hw);
        switch (j) {
            case 1: mean_std_warp_kernel<T, 1><<<grid, kWarpsPerBlock * 32, 0, st>>>(f, m, s, planes, nvec, ihw, eps); break;
            case 2: mean_std_warp_kernel<T, 2><<<grid, kWarpsPerBlock * 32, 0, st>>>(f, m, s, planes, nvec, ihw, eps); break;
            case 4: mean_std_warp_kernel<T, 4><<<grid, kWarpsPerBlock * 32, 0, st>>>(f, m, s, planes, nvec, ihw, eps); break;
            default: mean_std_warp_kernel<T, 8><<<grid, kWarpsPerBlock * 32, 0, st>>>(f, m, s, planes, nvec, ihw, eps); break;
        }
    } else if (vec) {
        // threads per plane: enough vectors per thread for two full batches of loads
        const int64_t nvec = hw / Vec16<T>::EPV;
        const int64_t vpt = 2 * kStreamBatch;   // (8 .. 64 measured within 5 % of each other on B200)
        const int g = nvec <= 32 * vpt ? 32 : nvec <= 64 * vpt ? 64 : nvec <= 128 * vpt ? 128 : 256;
        const unsigned grid = static_cast<unsigned>((planes + kCtaThreads / g - 1) / (kCtaThreads / g));
        switch (g) {
            case 32: mean_std_stream_kernel<T, 32><<<grid, kCtaThreads, 0, st>>>(f, m, s, planes, hw, eps); break;
            case 64: mean_std_stream_kernel<T, 64><<<grid, kCtaThreads, 0, st>>>(f, m, s, planes, hw, eps); break;
            case 128: mean_std_stream_kernel<T, 128><<<grid, kCtaThreads, 0, st>>>(f, m, s, planes, hw, eps); break;
            default: mean_std_stream_kernel<T, 256><<<grid, kCtaThreads, 0, st>>>(f, m, s, planes, hw, eps); break;
        }
    } else {
        mean_std_cta_kernel<T, false><<<static_cast<unsigned>(planes), kCtaThreads, 0, st>>>(f, m, s, hw, eps);
    }
    return check_launch("udape_mean_std");
}

static int adain_warps_per_block() {
    // planes per CTA.  Measured on B200 (N=32, fp32, r02e): 4 warps 32.6 us = 94.3 % of the HBM peak, 8 warps 33.1 us;
    // both directions in one launch 61.6 us = 99.8 % vs 63.1 us — smaller CTAs drain faster at the end of a launch
    const char* e = std::getenv("UDAPE_ADAIN_WARPS");   // tuning: 4 (default) | 8
    const int v = e ? std::atoi(e) : 0;
    return v == 8 ? 8 : 4;
}

template <typename T, int J>
static void launch_adain_warp(const AdainJobs& jobs, int n_jobs, int64_t planes, int nvc, int nvs, int ihc, int ihs, float eps,
                              cudaStream_t st) {
    const int wpb = adain_warps_per_block();
    const dim3 grid(static_cast<unsigned>((planes + wpb - 1) / wpb), static_cast<unsigned>(n_jobs));
    if (wpb == 4) adain_warp_kernel<T, J, 4><<<grid, 4 * 32, 0, st>>>(jobs, planes, nvc, nvs, ihc, ihs, eps);
    else adain_warp_kernel<T, J, 8><<<grid, 8 * 32, 0, st>>>(jobs, planes, nvc, nvs, ihc, ihs, eps);
}

// n_jobs jobs of one shape; the register path takes them in one launch, the generic path one launch per job
template <typename T>
static int launch_adain(const AdainJobs& jobs, int n_jobs, int64_t planes, int64_t hw_c, int64_t hw_s, float eps, cudaStream_t st) {
    bool vec_c = true, vec_s = true;
    for (int i = 0; i < n_jobs; ++i) {
        vec_c = vec_c && plane_vectorizable<T>(jobs.content[i], hw_c) && aligned16(jobs.out[i]);
        vec_s = vec_s && plane_vectorizable<T>(jobs.style[i], hw_s);
    }
    int j = 0;
    if (vec_c && vec_s) {
        const int64_t nv = (hw_c > hw_s ? hw_c : hw_s) / Vec16<T>::EPV;
        j = pick_j(nv);
    }
    if (j) {
        const int nvc = static_cast<int>(hw_c / Vec16<T>::EPV), nvs = static_cast<int>(hw_s / Vec16<T>::EPV);
        const int ihc = static_cast<int>(hw_c), ihs = static_cast<int>(hw_s);
        switch (j) {
            case 1: launch_adain_warp<T, 1>(jobs, n_jobs, planes, nvc, nvs, ihc, ihs, eps, st); break;
            case 2: launch_adain_warp<T, 2>(jobs, n_jobs, planes, nvc, nvs, ihc, ihs, eps, st); break;
            case 4: launch_adain_warp<T, 4>(jobs, n_jobs, planes, nvc, nvs, ihc, ihs, eps, st); break;
            default: launch_adain_warp<T, 8>(jobs, n_jobs, planes, nvc, nvs, ihc, ihs, eps, st); break;
        }
    } else {
        const unsigned grid = static_cast<unsigned>(planes);
        for (int i = 0; i < n_jobs; ++i) {
            const T* c = static_cast<const T*>(jobs.content[i]);
            const T* sy = static_cast<const T*>(jobs.style[i]);
            T* o = static_cast<T*>(jobs.out[i]);
            const float alpha = jobs.alpha[i];
            const float* alpha_dev = jobs.alpha_dev[i];
            const int mix = jobs.mix[i];
            if (vec_c && vec_s) adain_cta_kernel<T, true, true><<<grid, kCtaThreads, 0, st>>>(c, sy, o, hw_c, hw_s, eps, alpha, alpha_dev, mix);
            else if (vec_c) adain_cta_kernel<T, true, false><<<grid, kCtaThreads, 0, st>>>(c, sy, o, hw_c, hw_s, eps, alpha, alpha_dev, mix);
            else if (vec_s) adain_cta_kernel<T, false, true><<<grid, kCtaThreads, 0, st>>>(c, sy, o, hw_c, hw_s, eps, alpha, alpha_dev, mix);
            else adain_cta_kernel<T, false, false><<<grid, kCtaThreads, 0, st>>>(c, sy, o, hw_c, hw_s, eps, alpha, alpha_dev, mix);
        }
    }
    return check_launch("udape_adain_mix");
}

}  // namespace udape

using namespace udape;

extern "C" int udape_mean_std(const void* feat, int dtype, int64_t planes, int64_t hw, float eps,
                              void* mean, void* std, void* stream) {
    UDAPE_REQUIRE(feat && mean && std, UDAPE_ERR_NULL, "udape_mean_std: NULL pointer");
    UDAPE_REQUIRE(planes > 0 && hw > 0 && planes < (1ll << 31) && hw < (1ll << 31), UDAPE_ERR_SHAPE,
                  "udape_mean_std: bad extents planes=%lld hw=%lld", (long long)planes, (long long)hw);
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(es == 2 || es == 4, UDAPE_ERR_DTYPE, "udape_mean_std: unsupported dtype code %d", dtype);
    UDAPE_REQUIRE(aligned_to(feat, es) && aligned_to(mean, es) && aligned_to(std, es), UDAPE_ERR_ALIGN,
                  "udape_mean_std: pointer not aligned to element size");
    UDAPE_DISPATCH_FLOAT(dtype, T, return launch_mean_std<T>(feat, planes, hw, eps, mean, std, as_stream(stream)));
    return UDAPE_OK;
}

extern "C" int udape_mean_std_bwd(const void* feat, const void* mean, const void* std, const void* dmean,
                                  const void* dstd, int dtype, int64_t planes, int64_t hw, void* dfeat, void* stream) {
    UDAPE_REQUIRE(feat && mean && std && dfeat, UDAPE_ERR_NULL, "udape_mean_std_bwd: NULL pointer");
    UDAPE_REQUIRE(planes > 0 && hw > 0 && planes < (1ll << 31) && hw < (1ll << 31), UDAPE_ERR_SHAPE,
                  "udape_mean_std_bwd: bad extents planes=%lld hw=%lld", (long long)planes, (long long)hw);
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(es == 2 || es == 4, UDAPE_ERR_DTYPE, "udape_mean_std_bwd: unsupported dtype code %d", dtype);
    UDAPE_REQUIRE(aligned_to(feat, es) && aligned_to(mean, es) && aligned_to(std, es) && aligned_to(dfeat, es) &&
                      (!dmean || aligned_to(dmean, es)) && (!dstd || aligned_to(dstd, es)),
                  UDAPE_ERR_ALIGN, "udape_mean_std_bwd: pointer not aligned to element size");
    cudaStream_t st = as_stream(stream);
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        constexpr int EPV = Vec16<T>::EPV;
        const bool vec = plane_vectorizable<T>(feat, hw) && aligned16(dfeat);
        // a CTA streams up to 16 vectors per thread; larger planes are split so that small batches still fill the GPU
        const int64_t per_cta = static_cast<int64_t>(kCtaThreads) * 16 * EPV;
        int64_t splits = (hw + per_cta - 1) / per_cta;
        if (splits < 1) splits = 1;
        int64_t span = (hw + splits - 1) / splits;
        span = (span + EPV - 1) / EPV * EPV;
        splits = (hw + span - 1) / span;
        UDAPE_REQUIRE(planes * splits < (1ll << 31), UDAPE_ERR_SHAPE, "udape_mean_std_bwd: grid too large");
        const unsigned grid = static_cast<unsigned>(planes * splits);
        const T* f = static_cast<const T*>(feat);
        const T* m = static_cast<const T*>(mean);
        const T* s = static_cast<const T*>(std);
        const T* dm = static_cast<const T*>(dmean);
        const T* ds = static_cast<const T*>(dstd);
        T* o = static_cast<T*>(dfeat);
        if (vec) mean_std_bwd_kernel<T, true><<<grid, kCtaThreads, 0, st>>>(f, m, s, dm, ds, o, hw, static_cast<int>(splits), span);
        else mean_std_bwd_kernel<T, false><<<grid, kCtaThreads, 0, st>>>(f, m, s, dm, ds, o, hw, static_cast<int>(splits), span);
    });
    return check_launch("udape_mean_std_bwd");
}

static int adain_entry(const udape_adain_job* in_jobs, int n_jobs, int dtype, int64_t planes, int64_t hw_c, int64_t hw_s,
                       float eps, void* stream, const char* name) {
    UDAPE_REQUIRE(in_jobs, UDAPE_ERR_NULL, "%s: jobs is NULL", name);
    UDAPE_REQUIRE(n_jobs >= 1 && n_jobs <= UDAPE_MAX_ADAIN_JOBS, UDAPE_ERR_ARG, "%s: n_jobs=%d (1..%d)", name, n_jobs, UDAPE_MAX_ADAIN_JOBS);
    UDAPE_REQUIRE(planes > 0 && hw_c > 0 && hw_s > 0 && planes < (1ll << 31) && hw_c < (1ll << 31) &&
                      hw_s < (1ll << 31),
                  UDAPE_ERR_SHAPE, "%s: bad extents planes=%lld hw_c=%lld hw_s=%lld", name,
                  (long long)planes, (long long)hw_c, (long long)hw_s);
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(es == 2 || es == 4, UDAPE_ERR_DTYPE, "%s: unsupported dtype code %d", name, dtype);
    AdainJobs jobs = {};
    for (int i = 0; i < n_jobs; ++i) {
        const udape_adain_job& j = in_jobs[i];
        UDAPE_REQUIRE(j.content && j.style && j.out, UDAPE_ERR_NULL, "%s: NULL pointer (job %d)", name, i);
        UDAPE_REQUIRE(aligned_to(j.content, es) && aligned_to(j.style, es) && aligned_to(j.out, es), UDAPE_ERR_ALIGN,
                      "%s: pointer not aligned to element size (job %d)", name, i);
        if (!j.alpha_dev) {
            // Style_net.py:164 asserts 0 <= alpha <= 1
            UDAPE_REQUIRE(j.alpha >= 0.0f && j.alpha <= 1.0f, UDAPE_ERR_ARG, "%s: alpha %g outside [0,1]", name, (double)j.alpha);
        }
        jobs.content[i] = j.content; jobs.style[i] = j.style; jobs.out[i] = j.out;
        jobs.alpha_dev[i] = j.alpha_dev; jobs.alpha[i] = j.alpha;
        jobs.mix[i] = (j.alpha_dev != nullptr) || (j.alpha != 1.0f);
    }
    UDAPE_DISPATCH_FLOAT(dtype, T, return launch_adain<T>(jobs, n_jobs, planes, hw_c, hw_s, eps, as_stream(stream)));
    return UDAPE_OK;
}

extern "C" int udape_adain_mix(const void* content, const void* style, int dtype, int64_t planes,
                               int64_t hw_c, int64_t hw_s, float eps, float alpha,
                               const float* alpha_dev, void* out, void* stream) {
    UDAPE_REQUIRE(content && style && out, UDAPE_ERR_NULL, "udape_adain_mix: NULL pointer");
    const udape_adain_job job = {content, style, out, alpha_dev, alpha};
    return adain_entry(&job, 1, dtype, planes, hw_c, hw_s, eps, stream, "udape_adain_mix");
}

extern "C" int udape_adain_mix_multi(const udape_adain_job* jobs, int n_jobs, int dtype, int64_t planes, int64_t hw_c,
                                     int64_t hw_s, float eps, void* stream) {
    return adain_entry(jobs, n_jobs, dtype, planes, hw_c, hw_s, eps, stream, "udape_adain_mix_multi");
}
