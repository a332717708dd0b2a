// loss.cu — masked JointsMSELoss and teacher-student ConsLoss, forward and backward.
//
// Replaces lib/models/loss.py:11-49 (JointsMSELoss) and :119-132 (ConsLoss) of the
// reference, whose eager form makes ~8 passes forward plus autograd's backward
// (SURVEY.md §2.1).  Here: forward = one read of each operand, backward = one read of
// each operand + one write of the gradient, mixed dtypes (fp16/bf16 student output
// under autocast against an fp32 label / rectified teacher) loaded as 128-bit vectors
// and accumulated in fp32.
//
// Reductions are deterministic: every CTA reduces its plane with a fixed shuffle/smem
// tree and writes one partial; the last CTA to finish (ticket counter) sums the
// partials in a fixed order.  No floating-point atomics anywhere.
#include <cstdlib>

#include "common.cuh"
#include "gauss.cuh"

namespace udape {

constexpr int kLossThreads = 256;
constexpr int kLossUnroll = 4;  // independent vector groups in flight per thread
// the fused step runs 128-thread CTAs (two batches of loads per plane): measured 68 vs 72 us at 2x5376 planes
constexpr int kStepThreads = 128;

template <typename T> __device__ __forceinline__ float load_scalar(const void* p, int64_t i) {
    return to_f32<T>(static_cast<const T*>(p)[i]);
}
// per-plane weight / mask in one of {f32,f16,bf16,u8}; NULL = 1
__device__ __forceinline__ float load_plane_scalar(const void* p, int dtype, int64_t i) {
    if (p == nullptr) return 1.0f;
    switch (dtype) {
        case UDAPE_F32: return static_cast<const float*>(p)[i];
        case UDAPE_F16: return __half2float(static_cast<const __half*>(p)[i]);
        case UDAPE_BF16: return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]);
        default: return static_cast<const uint8_t*>(p)[i] ? 1.0f : 0.0f;
    }
}

// G-byte slice of the u8 validity mask owned by one thread (G in {4, 8})
template <int G> struct MaskBytes {
    uint32_t w[G / 4];
    __device__ __forceinline__ void load(const uint8_t* p) {
        if constexpr (G == 4) w[0] = *reinterpret_cast<const uint32_t*>(p);
        else { const uint2 t = *reinterpret_cast<const uint2*>(p); w[0] = t.x; w[1] = t.y; }
    }
    __device__ __forceinline__ bool on(int e) const { return (w[e >> 2] >> (8 * (e & 3))) & 0xffu; }
};

// sum over the plane of  vm[i] * (a[i]-b[i])^2   (vm = optional u8 validity mask)
template <typename TA, typename TB, bool VEC>
__device__ __forceinline__ float thread_sq_diff(const TA* __restrict__ a, const TB* __restrict__ b,
                                                const uint8_t* __restrict__ vm, int hw) {
    float acc = 0.0f;
    if (VEC) {
        constexpr int G = PairGroup<TA, TB>::G;
        const int ngrp = hw / G;
        for (int base = 0; base < ngrp; base += kLossUnroll * kLossThreads) {
            Pack<TA, G> ga[kLossUnroll];
            Pack<TB, G> gb[kLossUnroll];
            MaskBytes<G> m[kLossUnroll];
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int g = base + u * kLossThreads + threadIdx.x;
                if (g < ngrp) {
                    ga[u].load(a + G * g);
                    gb[u].load(b + G * g);
                    if (vm) m[u].load(vm + G * g);
                }
            }
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int g = base + u * kLossThreads + threadIdx.x;
                if (g < ngrp) {
                    float fa[G], fb[G];
                    ga[u].get(fa);
                    gb[u].get(fb);
                    float t = 0.0f;
#pragma unroll
                    for (int e = 0; e < G; ++e) {
                        const float d = fa[e] - fb[e];
                        const float sq = d * d;
                        t += (vm == nullptr || m[u].on(e)) ? sq : 0.0f;
                    }
                    acc += t;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kLossThreads) {
            const float d = to_f32<TA>(a[i]) - to_f32<TB>(b[i]);
            acc += (vm == nullptr || vm[i]) ? d * d : 0.0f;
        }
    }
    return acc;
}

// ---- JointsMSELoss ------------------------------------------------------------------------
template <typename TO, typename TT, bool VEC>
__global__ void __launch_bounds__(kLossThreads)
joints_mse_fwd_kernel(const TO* __restrict__ output, const TT* __restrict__ target,
                      const void* __restrict__ weight, int w_dtype, int hw,
                      float* __restrict__ plane_loss, float* __restrict__ loss_mean,
                      uint32_t* __restrict__ ticket) {
    __shared__ float red[32];
    const int64_t plane = blockIdx.x;
    const float acc = block_sum<kLossThreads>(
        thread_sq_diff<TO, TT, VEC>(output + plane * hw, target + plane * hw, nullptr, hw), red);
    if (threadIdx.x == 0) {
        // loss.py:43-45: mse(none) * 0.5 * weight, then mean over the plane (:49)
        const float wgt = load_plane_scalar(weight, w_dtype, plane);
        plane_loss[plane] = 0.5f * wgt * (acc / static_cast<float>(hw));
    }
    if (loss_mean == nullptr) return;
    if (last_block_done(ticket, gridDim.x)) {
        const float s = cta_sum_array(plane_loss, gridDim.x, red);
        if (threadIdx.x == 0) *loss_mean = s / static_cast<float>(gridDim.x);
    }
}

template <typename TO, typename TT, bool VEC>
__global__ void __launch_bounds__(kLossThreads)
joints_mse_bwd_kernel(const TO* __restrict__ output, const TT* __restrict__ target,
                      const void* __restrict__ weight, int w_dtype, int hw,
                      const float* __restrict__ grad_out, int grad_per_plane, float inv_count,
                      TO* __restrict__ grad_in) {
    const int64_t plane = blockIdx.x;
    const float g = grad_per_plane ? grad_out[plane] : grad_out[0];
    // d/do [0.5*w*(o-t)^2] * g / count
    const float coef = g * load_plane_scalar(weight, w_dtype, plane) * inv_count;
    const TO* o = output + plane * hw;
    const TT* t = target + plane * hw;
    TO* gi = grad_in + plane * hw;
    if (VEC) {
        constexpr int G = PairGroup<TO, TT>::G;
        const int ngrp = hw / G;
        for (int base = 0; base < ngrp; base += kLossUnroll * kLossThreads) {
            Pack<TO, G> go[kLossUnroll];
            Pack<TT, G> gt[kLossUnroll];
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int gidx = base + u * kLossThreads + threadIdx.x;
                if (gidx < ngrp) { go[u].load(o + G * gidx); gt[u].load(t + G * gidx); }
            }
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int gidx = base + u * kLossThreads + threadIdx.x;
                if (gidx < ngrp) {
                    float fo[G], ft[G];
                    go[u].get(fo);
                    gt[u].get(ft);
#pragma unroll
                    for (int e = 0; e < G; ++e) fo[e] = coef * (fo[e] - ft[e]);
                    Pack<TO, G>::store(gi + G * gidx, fo);
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kLossThreads)
            gi[i] = from_f32<TO>(coef * (to_f32<TO>(o[i]) - to_f32<TT>(t[i])));
    }
}

// ---- ConsLoss -----------------------------------------------------------------------------
// count of valid (b,i) positions; only launched when a valid_mask is supplied
__global__ void __launch_bounds__(kLossThreads)
count_valid_kernel(const uint8_t* __restrict__ vm, int64_t n, int32_t* __restrict__ count) {
    int c = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kLossThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kLossThreads)
        c += vm[i] ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);  // integer: order-independent
}

template <typename TS, typename TT, bool VEC>
__global__ void __launch_bounds__(kLossThreads)
cons_fwd_kernel(const TS* __restrict__ stu, const TT* __restrict__ tea,
                const void* __restrict__ tea_mask, int mask_dtype,
                const uint8_t* __restrict__ valid_mask, int joints, int hw,
                float* __restrict__ plane_partial, const int32_t* __restrict__ valid_count,
                float* __restrict__ loss, uint32_t* __restrict__ ticket) {
    __shared__ float red[32];
    const int64_t plane = blockIdx.x;
    const int64_t b = plane / joints;
    const float m = load_plane_scalar(tea_mask, mask_dtype, plane);
    const uint8_t* vm = valid_mask ? valid_mask + b * hw : nullptr;
    const float acc = block_sum<kLossThreads>(
        thread_sq_diff<TS, TT, VEC>(stu + plane * hw, tea + plane * hw, vm, hw), red);
    // (diff*m)^2 = m^2 * diff^2  (loss.py:125-128)
    if (threadIdx.x == 0) plane_partial[plane] = m * m * acc;
    if (last_block_done(ticket, gridDim.x)) {
        const float s = cta_sum_array(plane_partial, gridDim.x, red);
        if (threadIdx.x == 0) {
            // mean over k (loss.py:128) then mean over the kept (b,i) positions (:132)
            const float npos = valid_mask ? static_cast<float>(*valid_count)
                                          : static_cast<float>(gridDim.x / joints) * static_cast<float>(hw);
            *loss = s / static_cast<float>(joints) / npos;
        }
    }
}

template <typename TS, typename TT, bool VEC>
__global__ void __launch_bounds__(kLossThreads)
cons_bwd_kernel(const TS* __restrict__ stu, const TT* __restrict__ tea,
                const void* __restrict__ tea_mask, int mask_dtype,
                const uint8_t* __restrict__ valid_mask, int joints, int hw, int64_t batch,
                const float* __restrict__ grad_out, const int32_t* __restrict__ valid_count,
                TS* __restrict__ grad_stu) {
    const int64_t plane = blockIdx.x;
    const int64_t b = plane / joints;
    const float m = load_plane_scalar(tea_mask, mask_dtype, plane);
    const float npos = valid_mask ? static_cast<float>(*valid_count)
                                  : static_cast<float>(batch) * static_cast<float>(hw);
    const float coef = 2.0f * grad_out[0] * m * m / (static_cast<float>(joints) * npos);
    const uint8_t* vm = valid_mask ? valid_mask + b * hw : nullptr;
    const TS* s = stu + plane * hw;
    const TT* t = tea + plane * hw;
    TS* gs = grad_stu + plane * hw;
    if (VEC) {
        constexpr int G = PairGroup<TS, TT>::G;
        const int ngrp = hw / G;
        for (int base = 0; base < ngrp; base += kLossUnroll * kLossThreads) {
            Pack<TS, G> g1[kLossUnroll];
            Pack<TT, G> g2[kLossUnroll];
            MaskBytes<G> mk[kLossUnroll];
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int gidx = base + u * kLossThreads + threadIdx.x;
                if (gidx < ngrp) {
                    g1[u].load(s + G * gidx);
                    g2[u].load(t + G * gidx);
                    if (vm) mk[u].load(vm + G * gidx);
                }
            }
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int gidx = base + u * kLossThreads + threadIdx.x;
                if (gidx < ngrp) {
                    float fs[G], ft[G];
                    g1[u].get(fs);
                    g2[u].get(ft);
#pragma unroll
                    for (int e = 0; e < G; ++e) {
                        const float gval = coef * (fs[e] - ft[e]);
                        fs[e] = (vm == nullptr || mk[u].on(e)) ? gval : 0.0f;
                    }
                    Pack<TS, G>::store(gs + G * gidx, fs);
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kLossThreads) {
            const float gval = coef * (to_f32<TS>(s[i]) - to_f32<TT>(t[i]));
            gs[i] = from_f32<TS>((vm == nullptr || vm[i]) ? gval : 0.0f);
        }
    }
}


// Consistency plane of the fused step when the rectified teacher map is NOT materialised: only the
// student heatmap is read (128-bit vectors of the student dtype) and its gradient written; the
// teacher value is zero for almost every vector and a shared-memory table lookup inside the
// (6*sigma+1)^2 window around the decoded arg-max.  Returns the thread's sum of (s - t)^2.
// The loads of a batch (kLossUnroll vectors of a thread), apart from their use: the pair kernel below issues the first
// batch of BOTH of its planes before anything else.
template <typename TS, int THREADS>
__device__ __forceinline__ void cons_batch_load(Pack<TS, 16 / static_cast<int>(sizeof(TS))> (&gs_)[kLossUnroll],
                                                const TS* __restrict__ s, int base, int ngrp) {
    constexpr int G = 16 / static_cast<int>(sizeof(TS));
#pragma unroll
    for (int u = 0; u < kLossUnroll; ++u) {
        const int gi = base + u * THREADS + threadIdx.x;
        if (gi < ngrp) gs_[u].load(s + G * gi);
    }
}

// PRE: `gs_` already holds the first batch (cons_batch_load(gs_, s, 0, ngrp) was issued by the caller).
template <typename TS, typename TT, int THREADS, bool PRE = false>
__device__ __forceinline__ float cons_plane_analytic(const TS* __restrict__ s, TS* __restrict__ gout, int hw, int w,
                                                     const RectGeom& geom, const GaussWindow& gw,
                                                     const float* __restrict__ tab, float coef,
                                                     Pack<TS, 16 / static_cast<int>(sizeof(TS))> (&gs_)[kLossUnroll]) {
    constexpr int G = 16 / static_cast<int>(sizeof(TS));
    const int ngrp = hw / G;
    int y = (threadIdx.x * G) / w, x = threadIdx.x * G - y * w;  // one division, then incremental
    const int step_y = (THREADS * G) / w, step_x = THREADS * G - step_y * w;
    float acc = 0.0f;
    for (int base = 0; base < ngrp; base += kLossUnroll * THREADS) {
        if (!PRE || base > 0) cons_batch_load<TS, THREADS>(gs_, s, base, ngrp);
#pragma unroll
        for (int u = 0; u < kLossUnroll; ++u) {
            const int gi = base + u * THREADS + threadIdx.x;
            if (gi < ngrp) {
                float fs[G];
                gs_[u].get(fs);
                float tsum = 0.0f;
                if (x + G <= w && (y < geom.y0i || y >= geom.y1i || x + G <= geom.x0i || x >= geom.x1i)) {
                    // the group misses the window: the teacher map is zero here
#pragma unroll
                    for (int e = 0; e < G; ++e) {
                        tsum = fmaf(fs[e], fs[e], tsum);
                        fs[e] = coef * fs[e];
                    }
                } else {
                    int xx = x, yy = y;
#pragma unroll
                    for (int e = 0; e < G; ++e) {
                        // rounded through the teacher map's dtype, like the materialised map
                        const float tv = to_f32<TT>(from_f32<TT>(rectified_value(xx, yy, geom, gw, tab)));
                        if (++xx == w) { xx = 0; ++yy; }
                        const float d = fs[e] - tv;
                        tsum = fmaf(d, d, tsum);
                        fs[e] = coef * d;
                    }
                }
                acc += tsum;
                if (gout) Pack<TS, G>::store(gout + G * gi, fs);
                x += step_x; y += step_y;
                if (x >= w) { x -= w; ++y; }
            }
        }
    }
    return acc;
}


// ---- fused loss step ------------------------------------------------------------------------
// One launch for  loss_s = JointsMSELoss(y_s, label, weight),  loss_c = ConsLoss(y_t_stu, tea,
// tea_mask),  loss_all = loss_s + lambda_c*loss_c  AND the gradients of  grad_scale*loss_all
// w.r.t. both student heatmaps (train_human.py:425-436): every operand is read once and each
// gradient written once.  CTAs [0, planes_s) own a supervised plane, CTAs [planes_s,
// planes_s+planes_t) a consistency plane.  When `tea` is NULL the rectified teacher map is not
// read at all: it is evaluated from the decoded arg-max (tea_preds) with the very functions
// the decode kernel uses to materialise it (gauss.cuh), so both routes give identical values.
struct LossStepArgs {
    const void* weight; int w_dtype;
    const void* tea_mask; int mask_dtype;
    const float* tea_preds;      // [planes_t, 2] float (x, y), zeroed when max <= 0 (decode's `preds`)
    GaussWindow gw;
    int planes_s, planes_t, joints, hw, w, h;
    float lambda_c, grad_scale;
    const float* grad_scale_dev;
    float* partial;              // [planes_s + planes_t]
    float* losses;               // [3] = loss_all, loss_s, loss_c
    uint32_t* ticket;
};

template <typename TS, typename TT, bool VEC>
__global__ void __launch_bounds__(kStepThreads)
loss_step_kernel(const TS* __restrict__ y_s, const TT* __restrict__ label, const TS* __restrict__ y_t,
                 const TT* __restrict__ tea, TS* __restrict__ grad_s, TS* __restrict__ grad_t,
                 const LossStepArgs a) {
    __shared__ float red[32];
    __shared__ float s_tab[kWinTabN * kWinTabN];
    const int hw = a.hw;
    // Supervised and consistency planes ALTERNATE over the grid (pairs first, the longer set's rest behind them): a
    // consistency plane of the analytic route brings 8 KB into flight, a supervised one 24 KB, and a wave made of
    // consistency planes only (the second half of a grid ordered by kind) left HBM idle — Little's law:
    // 8 resident CTAs x 8 KB per SM at 2-4 us of loaded latency is 2.7 TB/s, which is what that half ran at.
    const int npair = min(a.planes_s, a.planes_t);
    const int bid = static_cast<int>(blockIdx.x);
    const bool sup = bid < 2 * npair ? (bid & 1) == 0 : a.planes_s > a.planes_t;
    const int64_t plane = bid < 2 * npair ? (bid >> 1) : bid - npair;
    const int slot = sup ? static_cast<int>(plane) : a.planes_s + static_cast<int>(plane);   // partial[]: supervised planes first
    const float* tab = nullptr;
    if (!sup && tea == nullptr) {  // CTA-uniform: the window table of the analytic teacher map
        tab = build_window_table(s_tab, a.gw);
        __syncthreads();
    }
    const float g = a.grad_scale_dev ? __ldg(a.grad_scale_dev) : a.grad_scale;
    const TS* s;
    const TT* t;
    TS* gout;
    float coef, plane_scale;
    RectGeom geom = {};
    bool analytic = false;
    if (sup) {
        const float wgt = load_plane_scalar(a.weight, a.w_dtype, plane);
        s = y_s + plane * hw;
        t = label + plane * hw;
        gout = grad_s ? grad_s + plane * hw : nullptr;
        // d/do [0.5*w*(o-t)^2] / (planes*hw)
        coef = g * wgt / (static_cast<float>(a.planes_s) * static_cast<float>(hw));
        plane_scale = 0.5f * wgt / static_cast<float>(hw);
    } else {
        const float m = load_plane_scalar(a.tea_mask, a.mask_dtype, plane);
        s = y_t + plane * hw;
        analytic = (tea == nullptr);
        t = analytic ? nullptr : tea + plane * hw;
        gout = grad_t ? grad_t + plane * hw : nullptr;
        // d/ds [m^2 (s-t)^2] / (K * B*hw)   with B = planes_t / K
        coef = 2.0f * (g * a.lambda_c) * m * m / (static_cast<float>(a.joints) *
               (static_cast<float>(a.planes_t / a.joints) * static_cast<float>(hw)));
        plane_scale = m * m;
        if (analytic) geom = rect_geometry(a.tea_preds[2 * plane], a.tea_preds[2 * plane + 1], a.h, a.w, a.gw);
    }
    float acc = 0.0f;
    if (VEC && analytic) {
        Pack<TS, 16 / static_cast<int>(sizeof(TS))> gs_[kLossUnroll];
        acc = cons_plane_analytic<TS, TT, kStepThreads>(s, gout, hw, a.w, geom, a.gw, tab, coef, gs_);
    } else if (VEC) {
        constexpr int G = PairGroup<TS, TT>::G;
        const int ngrp = hw / G;
        for (int base = 0; base < ngrp; base += kLossUnroll * kStepThreads) {
            Pack<TS, G> gs_[kLossUnroll];
            Pack<TT, G> gt_[kLossUnroll];
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int gi = base + u * kStepThreads + threadIdx.x;
                if (gi < ngrp) {
                    gs_[u].load(s + G * gi);
                    gt_[u].load(t + G * gi);
                }
            }
#pragma unroll
            for (int u = 0; u < kLossUnroll; ++u) {
                const int gi = base + u * kStepThreads + threadIdx.x;
                if (gi < ngrp) {
                    float fs[G], ft[G];
                    gs_[u].get(fs);
                    gt_[u].get(ft);
                    float tsum = 0.0f;
#pragma unroll
                    for (int e = 0; e < G; ++e) {
                        const float d = fs[e] - ft[e];
                        tsum = fmaf(d, d, tsum);
                        fs[e] = coef * d;
                    }
                    acc += tsum;
                    if (gout) Pack<TS, G>::store(gout + G * gi, fs);
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kStepThreads) {
            float tv;
            if (!analytic) tv = to_f32<TT>(t[i]);
            else {
                const int y = i / a.w, x = i - y * a.w;
                tv = to_f32<TT>(from_f32<TT>(rectified_value(x, y, geom, a.gw, tab)));
            }
            const float d = to_f32<TS>(s[i]) - tv;
            acc = fmaf(d, d, acc);
            if (gout) gout[i] = from_f32<TS>(coef * d);
        }
    }
    acc = block_sum<kStepThreads>(acc, red);
    if (threadIdx.x == 0) a.partial[slot] = plane_scale * acc;
    if (last_block_done(a.ticket, gridDim.x)) {
        const float ss = a.planes_s ? cta_sum_array(a.partial, a.planes_s, red) : 0.0f;
        const float sc = a.planes_t ? cta_sum_array(a.partial + a.planes_s, a.planes_t, red) : 0.0f;
        if (threadIdx.x == 0) {
            const float loss_s = a.planes_s ? ss / static_cast<float>(a.planes_s) : 0.0f;
            const float loss_c = a.planes_t ? sc / static_cast<float>(a.joints) /
                                     (static_cast<float>(a.planes_t / a.joints) * static_cast<float>(hw)) : 0.0f;
            a.losses[0] = loss_s + a.lambda_c * loss_c;   // train_human.py:434
            a.losses[1] = loss_s;
            a.losses[2] = loss_c;
        }
    }
}

// ---- fused loss step, one supervised + one consistency plane per CTA ---------------------------------------------
// The analytic route of the step (no teacher map) at equal plane counts: CTA i owns supervised plane i AND consistency
// plane i.  The first batch of loads of both planes is issued before the window table is built and before any
// per-plane scalar is looked at (160 bytes per thread in flight instead of 64 / 96), a CTA pays one table, one
// launch slot and one ticket for 40 KB of traffic instead of one each for 16 KB and 32 KB, and the grid is half as
// long.  Per-thread order of additions, block trees, partial slots and the last CTA's sums are those of
// loss_step_kernel: the results are bit-identical.
template <typename TS, typename TT>
__global__ void __launch_bounds__(kStepThreads)
loss_step_pair_kernel(const TS* __restrict__ y_s, const TT* __restrict__ label, const TS* __restrict__ y_t,
                      TS* __restrict__ grad_s, TS* __restrict__ grad_t, const LossStepArgs a) {
    __shared__ float red[32];
    __shared__ float s_tab[kWinTabN * kWinTabN];
    constexpr int T = kStepThreads, U = kLossUnroll;
    constexpr int GC = 16 / static_cast<int>(sizeof(TS));
    constexpr int GS = PairGroup<TS, TT>::G;
    const int hw = a.hw, ngc = hw / GC, ngs = hw / GS;
    const int64_t plane = blockIdx.x;
    const TS* sc = y_t + plane * hw;
    const TS* ss = y_s + plane * hw;
    const TT* tl = label + plane * hw;
    TS* gc = grad_t ? grad_t + plane * hw : nullptr;
    TS* gs = grad_s ? grad_s + plane * hw : nullptr;
    Pack<TS, GC> c_[U];
    Pack<TS, GS> s_[U];
    Pack<TT, GS> t_[U];
    cons_batch_load<TS, T>(c_, sc, 0, ngc);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int gi = u * T + threadIdx.x;
        if (gi < ngs) {
            s_[u].load(ss + GS * gi);
            t_[u].load(tl + GS * gi);
        }
    }
    const float* tab = build_window_table(s_tab, a.gw);
    const float g = a.grad_scale_dev ? __ldg(a.grad_scale_dev) : a.grad_scale;
    const float wgt = load_plane_scalar(a.weight, a.w_dtype, plane);
    const float m = load_plane_scalar(a.tea_mask, a.mask_dtype, plane);
    const float coef_s = g * wgt / (static_cast<float>(a.planes_s) * static_cast<float>(hw));
    const float coef_c = 2.0f * (g * a.lambda_c) * m * m / (static_cast<float>(a.joints) *
                         (static_cast<float>(a.planes_t / a.joints) * static_cast<float>(hw)));
    const RectGeom geom = rect_geometry(a.tea_preds[2 * plane], a.tea_preds[2 * plane + 1], a.h, a.w, a.gw);
    __syncthreads();   // the table
    float acc_c = cons_plane_analytic<TS, TT, T, true>(sc, gc, hw, a.w, geom, a.gw, tab, coef_c, c_);
    float acc_s = 0.0f;
    for (int base = 0; base < ngs; base += U * T) {
        if (base > 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int gi = base + u * T + threadIdx.x;
                if (gi < ngs) {
                    s_[u].load(ss + GS * gi);
                    t_[u].load(tl + GS * gi);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int gi = base + u * T + threadIdx.x;
            if (gi < ngs) {
                float fs[GS], ft[GS];
                s_[u].get(fs);
                t_[u].get(ft);
                float tsum = 0.0f;
#pragma unroll
                for (int e = 0; e < GS; ++e) {
                    const float d = fs[e] - ft[e];
                    tsum = fmaf(d, d, tsum);
                    fs[e] = coef_s * d;
                }
                acc_s += tsum;
                if (gs) Pack<TS, GS>::store(gs + GS * gi, fs);
            }
        }
    }
    acc_s = block_sum<T>(acc_s, red);
    if (threadIdx.x == 0) a.partial[plane] = (0.5f * wgt / static_cast<float>(hw)) * acc_s;
    acc_c = block_sum<T>(acc_c, red);
    if (threadIdx.x == 0) a.partial[a.planes_s + plane] = (m * m) * acc_c;
    if (last_block_done(a.ticket, gridDim.x)) {
        const float sum_s = cta_sum_array(a.partial, a.planes_s, red);
        const float sum_c = cta_sum_array(a.partial + a.planes_s, a.planes_t, red);
        if (threadIdx.x == 0) {
            const float loss_s = sum_s / static_cast<float>(a.planes_s);
            const float loss_c = sum_c / static_cast<float>(a.joints) /
                                 (static_cast<float>(a.planes_t / a.joints) * static_cast<float>(hw));
            a.losses[0] = loss_s + a.lambda_c * loss_c;   // train_human.py:434
            a.losses[1] = loss_s;
            a.losses[2] = loss_c;
        }
    }
}

template <typename TA, typename TB>
static bool pair_vectorizable(const void* a, const void* b, const void* c, const void* vm, int64_t hw) {
    // every plane must start on a 16-byte boundary of the wide operand and hold a whole
    // number of 8-element groups (covers G = 4 and G = 8)
    return (hw % 8) == 0 && aligned16(a) && aligned16(b) && (c == nullptr || aligned16(c)) &&
           (vm == nullptr || aligned_to(vm, 8));
}

static bool plane_scalar_dtype_ok(int d) { return d == UDAPE_F32 || d == UDAPE_F16 || d == UDAPE_BF16 || d == UDAPE_U8; }

}  // namespace udape

using namespace udape;

#define UDAPE_LOSS_COMMON_CHECKS(NAME, A, AD, B, BD, PLANES, HW)                                          \
    UDAPE_REQUIRE((A) && (B), UDAPE_ERR_NULL, NAME ": NULL tensor pointer");                              \
    UDAPE_REQUIRE((PLANES) > 0 && (HW) > 0 && (PLANES) < (1ll << 31) && (HW) < (1ll << 31), UDAPE_ERR_SHAPE, \
                  NAME ": bad extents planes=%lld hw=%lld", (long long)(PLANES), (long long)(HW));        \
    UDAPE_REQUIRE((dtype_size(AD) == 2 || dtype_size(AD) == 4) && (dtype_size(BD) == 2 || dtype_size(BD) == 4), \
                  UDAPE_ERR_DTYPE, NAME ": unsupported dtype codes %d/%d", (int)(AD), (int)(BD));         \
    UDAPE_REQUIRE(aligned_to((A), dtype_size(AD)) && aligned_to((B), dtype_size(BD)), UDAPE_ERR_ALIGN,    \
                  NAME ": misaligned tensor pointer")

extern "C" int udape_joints_mse_fwd(const void* output, int out_dtype, const void* target, int tgt_dtype,
                                    const void* weight, int w_dtype, int64_t planes, int64_t hw,
                                    float* plane_loss, float* loss_mean, uint32_t* ticket, void* stream) {
    UDAPE_LOSS_COMMON_CHECKS("udape_joints_mse_fwd", output, out_dtype, target, tgt_dtype, planes, hw);
    UDAPE_REQUIRE(plane_loss, UDAPE_ERR_NULL, "udape_joints_mse_fwd: plane_loss is NULL");
    UDAPE_REQUIRE(!loss_mean || ticket, UDAPE_ERR_NULL, "udape_joints_mse_fwd: ticket scratch is NULL");
    UDAPE_REQUIRE(!weight || plane_scalar_dtype_ok(w_dtype), UDAPE_ERR_DTYPE, "udape_joints_mse_fwd: bad weight dtype %d", w_dtype);
    cudaStream_t st = as_stream(stream);
    const unsigned grid = static_cast<unsigned>(planes);
    const int ihw = static_cast<int>(hw);
    UDAPE_DISPATCH_FLOAT(out_dtype, TO, UDAPE_DISPATCH_FLOAT(tgt_dtype, TT, {
        const TO* o = static_cast<const TO*>(output);
        const TT* t = static_cast<const TT*>(target);
        if (pair_vectorizable<TO, TT>(output, target, nullptr, nullptr, hw))
            joints_mse_fwd_kernel<TO, TT, true><<<grid, kLossThreads, 0, st>>>(o, t, weight, w_dtype, ihw, plane_loss, loss_mean, ticket);
        else
            joints_mse_fwd_kernel<TO, TT, false><<<grid, kLossThreads, 0, st>>>(o, t, weight, w_dtype, ihw, plane_loss, loss_mean, ticket);
    }));
    return check_launch("udape_joints_mse_fwd");
}

extern "C" int udape_joints_mse_bwd(const void* output, int out_dtype, const void* target, int tgt_dtype,
                                    const void* weight, int w_dtype, int64_t planes, int64_t hw,
                                    const float* grad_out, int grad_per_plane, void* grad_in, void* stream) {
    UDAPE_LOSS_COMMON_CHECKS("udape_joints_mse_bwd", output, out_dtype, target, tgt_dtype, planes, hw);
    UDAPE_REQUIRE(grad_out && grad_in, UDAPE_ERR_NULL, "udape_joints_mse_bwd: NULL gradient pointer");
    UDAPE_REQUIRE(!weight || plane_scalar_dtype_ok(w_dtype), UDAPE_ERR_DTYPE, "udape_joints_mse_bwd: bad weight dtype %d", w_dtype);
    cudaStream_t st = as_stream(stream);
    const unsigned grid = static_cast<unsigned>(planes);
    const int ihw = static_cast<int>(hw);
    // 'mean': g / (planes*hw); 'none': g[p] / hw
    const float inv_count = grad_per_plane ? 1.0f / static_cast<float>(hw)
                                           : 1.0f / (static_cast<float>(planes) * static_cast<float>(hw));
    UDAPE_DISPATCH_FLOAT(out_dtype, TO, UDAPE_DISPATCH_FLOAT(tgt_dtype, TT, {
        const TO* o = static_cast<const TO*>(output);
        const TT* t = static_cast<const TT*>(target);
        TO* gi = static_cast<TO*>(grad_in);
        if (pair_vectorizable<TO, TT>(output, target, grad_in, nullptr, hw))
            joints_mse_bwd_kernel<TO, TT, true><<<grid, kLossThreads, 0, st>>>(o, t, weight, w_dtype, ihw, grad_out, grad_per_plane, inv_count, gi);
        else
            joints_mse_bwd_kernel<TO, TT, false><<<grid, kLossThreads, 0, st>>>(o, t, weight, w_dtype, ihw, grad_out, grad_per_plane, inv_count, gi);
    }));
    return check_launch("udape_joints_mse_bwd");
}

extern "C" int udape_cons_fwd(const void* stu, int stu_dtype, const void* tea, int tea_dtype,
                              const void* tea_mask, int mask_dtype, const uint8_t* valid_mask,
                              int64_t batch, int64_t joints, int64_t hw, float* plane_partial,
                              int32_t* valid_count, float* loss, uint32_t* ticket, void* stream) {
    const int64_t planes = batch * joints;
    UDAPE_REQUIRE(batch > 0 && joints > 0, UDAPE_ERR_SHAPE, "udape_cons_fwd: bad extents B=%lld K=%lld", (long long)batch, (long long)joints);
    UDAPE_LOSS_COMMON_CHECKS("udape_cons_fwd", stu, stu_dtype, tea, tea_dtype, planes, hw);
    UDAPE_REQUIRE(plane_partial && loss && ticket, UDAPE_ERR_NULL, "udape_cons_fwd: NULL scratch/output pointer");
    UDAPE_REQUIRE(!valid_mask || valid_count, UDAPE_ERR_NULL, "udape_cons_fwd: valid_count is NULL");
    UDAPE_REQUIRE(!tea_mask || plane_scalar_dtype_ok(mask_dtype), UDAPE_ERR_DTYPE, "udape_cons_fwd: bad mask dtype %d", mask_dtype);
    cudaStream_t st = as_stream(stream);
    if (valid_mask) {
        cudaError_t e = cudaMemsetAsync(valid_count, 0, sizeof(int32_t), st);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_cons_fwd: memset: %s", cudaGetErrorString(e));
    }
    if (valid_mask) {
        const int64_t n = batch * hw;
        const unsigned g = static_cast<unsigned>((n + kLossThreads * 8 - 1) / (kLossThreads * 8));
        count_valid_kernel<<<g ? g : 1, kLossThreads, 0, st>>>(valid_mask, n, valid_count);
    }
    const unsigned grid = static_cast<unsigned>(planes);
    const int ihw = static_cast<int>(hw), ij = static_cast<int>(joints);
    UDAPE_DISPATCH_FLOAT(stu_dtype, TS, UDAPE_DISPATCH_FLOAT(tea_dtype, TT, {
        const TS* s = static_cast<const TS*>(stu);
        const TT* t = static_cast<const TT*>(tea);
        if (pair_vectorizable<TS, TT>(stu, tea, nullptr, valid_mask, hw))
            cons_fwd_kernel<TS, TT, true><<<grid, kLossThreads, 0, st>>>(s, t, tea_mask, mask_dtype, valid_mask, ij, ihw, plane_partial, valid_count, loss, ticket);
        else
            cons_fwd_kernel<TS, TT, false><<<grid, kLossThreads, 0, st>>>(s, t, tea_mask, mask_dtype, valid_mask, ij, ihw, plane_partial, valid_count, loss, ticket);
    }));
    return check_launch("udape_cons_fwd");
}

extern "C" int udape_cons_bwd(const void* stu, int stu_dtype, const void* tea, int tea_dtype,
                              const void* tea_mask, int mask_dtype, const uint8_t* valid_mask,
                              int64_t batch, int64_t joints, int64_t hw, const float* grad_out,
                              const int32_t* valid_count, void* grad_stu, void* stream) {
    const int64_t planes = batch * joints;
    UDAPE_REQUIRE(batch > 0 && joints > 0, UDAPE_ERR_SHAPE, "udape_cons_bwd: bad extents B=%lld K=%lld", (long long)batch, (long long)joints);
    UDAPE_LOSS_COMMON_CHECKS("udape_cons_bwd", stu, stu_dtype, tea, tea_dtype, planes, hw);
    UDAPE_REQUIRE(grad_out && grad_stu, UDAPE_ERR_NULL, "udape_cons_bwd: NULL gradient pointer");
    UDAPE_REQUIRE(!valid_mask || valid_count, UDAPE_ERR_NULL, "udape_cons_bwd: valid_count is NULL");
    UDAPE_REQUIRE(!tea_mask || plane_scalar_dtype_ok(mask_dtype), UDAPE_ERR_DTYPE, "udape_cons_bwd: bad mask dtype %d", mask_dtype);
    cudaStream_t st = as_stream(stream);
    const unsigned grid = static_cast<unsigned>(planes);
    const int ihw = static_cast<int>(hw), ij = static_cast<int>(joints);
    UDAPE_DISPATCH_FLOAT(stu_dtype, TS, UDAPE_DISPATCH_FLOAT(tea_dtype, TT, {
        const TS* s = static_cast<const TS*>(stu);
        const TT* t = static_cast<const TT*>(tea);
        TS* gs = static_cast<TS*>(grad_stu);
        if (pair_vectorizable<TS, TT>(stu, tea, grad_stu, valid_mask, hw))
            cons_bwd_kernel<TS, TT, true><<<grid, kLossThreads, 0, st>>>(s, t, tea_mask, mask_dtype, valid_mask, ij, ihw, batch, grad_out, valid_count, gs);
        else
            cons_bwd_kernel<TS, TT, false><<<grid, kLossThreads, 0, st>>>(s, t, tea_mask, mask_dtype, valid_mask, ij, ihw, batch, grad_out, valid_count, gs);
    }));
    return check_launch("udape_cons_bwd");
}

extern "C" int udape_loss_step(const void* y_s, const void* label, const void* weight, int w_dtype,
                               int64_t planes_s, const void* y_t_stu, const void* tea,
                               const float* tea_preds, double sigma, const void* tea_mask, int mask_dtype,
                               int64_t batch_t, int64_t joints, int64_t h, int64_t w, int stu_dtype,
                               int tgt_dtype, float lambda_c, float grad_scale, const float* grad_scale_dev,
                               float* partial, float* losses, uint32_t* ticket, void* grad_y_s,
                               void* grad_y_t_stu, void* stream) {
    const int64_t planes_t = batch_t * joints, hw = h * w;
    UDAPE_REQUIRE(planes_s >= 0 && batch_t >= 0 && joints > 0 && h > 0 && w > 0 && hw < (1ll << 31) &&
                      planes_s + planes_t > 0 && planes_s + planes_t < (1ll << 31),
                  UDAPE_ERR_SHAPE, "udape_loss_step: bad extents planes_s=%lld B_t=%lld K=%lld h=%lld w=%lld",
                  (long long)planes_s, (long long)batch_t, (long long)joints, (long long)h, (long long)w);
    const int es = dtype_size(stu_dtype), et = dtype_size(tgt_dtype);
    UDAPE_REQUIRE((es == 2 || es == 4) && (et == 2 || et == 4), UDAPE_ERR_DTYPE,
                  "udape_loss_step: unsupported dtype codes %d/%d", stu_dtype, tgt_dtype);
    UDAPE_REQUIRE(partial && losses && ticket, UDAPE_ERR_NULL, "udape_loss_step: NULL scratch/output pointer");
    UDAPE_REQUIRE(planes_s == 0 || (y_s && label), UDAPE_ERR_NULL, "udape_loss_step: NULL supervised operand");
    UDAPE_REQUIRE(planes_t == 0 || (y_t_stu && (tea || tea_preds)), UDAPE_ERR_NULL,
                  "udape_loss_step: consistency pair needs y_t_stu and either tea or tea_preds");
    UDAPE_REQUIRE(!weight || plane_scalar_dtype_ok(w_dtype), UDAPE_ERR_DTYPE, "udape_loss_step: bad weight dtype %d", w_dtype);
    UDAPE_REQUIRE(!tea_mask || plane_scalar_dtype_ok(mask_dtype), UDAPE_ERR_DTYPE, "udape_loss_step: bad mask dtype %d", mask_dtype);
    UDAPE_REQUIRE(aligned_to(y_s, es) && aligned_to(y_t_stu, es) && aligned_to(grad_y_s, es) &&
                      aligned_to(grad_y_t_stu, es) && aligned_to(label, et) && aligned_to(tea, et) &&
                      aligned_to(tea_preds, 4),
                  UDAPE_ERR_ALIGN, "udape_loss_step: misaligned tensor pointer");
    if (planes_t && !tea) {
        UDAPE_REQUIRE(sigma > 0.0 && sigma < 1e4, UDAPE_ERR_ARG, "udape_loss_step: sigma %g out of range", sigma);
    }
    cudaStream_t st = as_stream(stream);
    LossStepArgs a;
    a.weight = weight; a.w_dtype = w_dtype;
    a.tea_mask = tea_mask; a.mask_dtype = mask_dtype;
    a.tea_preds = tea_preds;
    a.gw = make_window((planes_t && !tea) ? sigma : 1.0);
    a.planes_s = static_cast<int>(planes_s); a.planes_t = static_cast<int>(planes_t);
    a.joints = static_cast<int>(joints); a.hw = static_cast<int>(hw);
    a.w = static_cast<int>(w); a.h = static_cast<int>(h);
    a.lambda_c = lambda_c; a.grad_scale = grad_scale; a.grad_scale_dev = grad_scale_dev;
    a.partial = partial; a.losses = losses; a.ticket = ticket;
    const unsigned grid = static_cast<unsigned>(planes_s + planes_t);
    const bool vec = (hw % 8) == 0 && aligned16(y_s) && aligned16(label) && aligned16(y_t_stu) && aligned16(tea) &&
                     aligned16(grad_y_s) && aligned16(grad_y_t_stu);
    // one supervised + one consistency plane per CTA: the analytic route at equal plane counts (what the step runs)
    // — for batches that fill the GPU: 56.0 -> 53.9 us at C5 (5376 pairs, 72 -> 75 % of the HBM roofline), but fewer and
    // longer CTAs are slower when the grid is a wave or two (C4, 1152 pairs: 16.4 -> 17.6 us; C2: 10.2 -> 10.7 us)
    bool pair = !tea && planes_s == planes_t && planes_s > 0;
    if (const char* e = std::getenv("UDAPE_LOSS_PAIR")) {   // tests / tuning: force the route
        pair = pair && e[0] == '1';
    } else if (pair) {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 148;
        pair = planes_s >= 16ll * sms;
    }
    UDAPE_DISPATCH_FLOAT(stu_dtype, TS, UDAPE_DISPATCH_FLOAT(tgt_dtype, TT, {
        const TS* ys = static_cast<const TS*>(y_s);
        const TS* yt = static_cast<const TS*>(y_t_stu);
        const TT* lb = static_cast<const TT*>(label);
        const TT* te = static_cast<const TT*>(tea);
        TS* g1 = static_cast<TS*>(grad_y_s);
        TS* g2 = static_cast<TS*>(grad_y_t_stu);
        if (vec && pair) loss_step_pair_kernel<TS, TT><<<static_cast<unsigned>(planes_s), kStepThreads, 0, st>>>(ys, lb, yt, g1, g2, a);
        else if (vec) loss_step_kernel<TS, TT, true><<<grid, kStepThreads, 0, st>>>(ys, lb, yt, te, g1, g2, a);
        else loss_step_kernel<TS, TT, false><<<grid, kStepThreads, 0, st>>>(ys, lb, yt, te, g1, g2, a);
    }));
    return check_launch("udape_loss_step");
}
