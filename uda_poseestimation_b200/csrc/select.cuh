// select.cuh — the k-th value select + tea_mask that runs in the LAST CTA of a launch that produced the activations
// (decode.cu: udape_decode_select; rewarp.cu: udape_rewarp_decode_select) or as its own single-CTA kernel
// (udape_mask_select).  train_human.py:427-430.
#pragma once

#include "common.cuh"

namespace udape {

// ---- k-th value + tea_mask ----------------------------------------------------------------------
// 4-pass radix select on the ordered key by ONE CTA of any size (exact; NaN sorts last like
// torch.kthvalue), then tea_mask = (tea_mask_in * activates) > thresh.  Runs as its own single-CTA
// kernel (udape_mask_select) or in the last CTA of the decode launch (udape_decode_select), where
// `act` was written by the other CTAs of the same grid: read through L2 (__ldcg).
struct SelectArgs {
    int kth;                    // 1-based rank
    const float* tm_in;         // optional
    float* thresh_out;          // optional
    uint8_t* tm_out;            // optional
    uint32_t* ticket;           // decode launch only: zeroed, self-resetting; NULL = no select
    int cache_elems;            // decode_kernel only: floats of dynamic shared memory the launch reserved for the select's values
};

__device__ __forceinline__ void select_body(const float* __restrict__ act, int n, const SelectArgs& sa,
                                            float* __restrict__ cache = nullptr, int cache_elems = 0) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_mask, s_k;
    const int nthreads = blockDim.x;
    if (threadIdx.x == 0) { s_prefix = 0u; s_mask = 0u; s_k = static_cast<unsigned>(sa.kth); }
    // The values are fetched from L2 ONCE — four loads of a thread in flight at a time — and kept in shared memory
    // (`cache`, when the caller has room for n words) as ORDERED KEYS; the four passes and the mask read them there.
    // Re-read from L2 in every pass they were 5 x n/T dependent round trips in the LAST CTA of the decode launch —
    // on the teacher chain of the step — while the rest of the GPU had drained.  (key_value(order_key(v)) is v but
    // for -0.0 -> +0.0 and the NaN payload: neither changes `a > thresh`.)
    const bool cached = cache != nullptr && n <= cache_elems;
    uint32_t* __restrict__ keys = reinterpret_cast<uint32_t*>(cache);
    if (cached) {
        constexpr int UN = 4;
        for (int i0 = threadIdx.x; i0 < n; i0 += UN * nthreads) {
            float x[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) x[u] = (i0 + u * nthreads < n) ? __ldcg(act + i0 + u * nthreads) : 0.0f;
#pragma unroll
            for (int u = 0; u < UN; ++u)
                if (i0 + u * nthreads < n) keys[i0 + u * nthreads] = order_key(x[u]);   // (a thread only ever reads what it cached)
        }
    }
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int i = threadIdx.x; i < 256; i += nthreads) hist[i] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        // (Same-address shared-memory atomics are not what this pass waits for: aggregating the lanes of a warp per
        // bin made the select SLOWER both ways it was tried — with match.any 8.8 -> 14.8 us at C5 (profiles/r02ao), with
        // two rounds of ballot + leader add 7.7 -> 13.0 us (profiles/r02bc).)
        if (cached) {
#pragma unroll 2
            for (int i = threadIdx.x; i < n; i += nthreads) {
                const uint32_t key = keys[i];
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
        } else {
            for (int i = threadIdx.x; i < n; i += nthreads) {
                const uint32_t key = order_key(__ldcg(act + i));
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 finds the bin holding rank k: lane l owns bins [8l, 8l+8)
            const unsigned k = s_k;
            unsigned c[8], lane_sum = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[8 * threadIdx.x + j]; lane_sum += c[j]; }
            unsigned incl = lane_sum;  // inclusive prefix over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (threadIdx.x >= o) incl += t;
            }
            const unsigned excl = incl - lane_sum;
            const unsigned owner = __ballot_sync(0xffffffffu, incl >= k);
            if (static_cast<int>(threadIdx.x) == __ffs(owner) - 1) {
                unsigned cum = excl;
                int b = 0;
#pragma unroll
                for (; b < 7; ++b) {
                    if (cum + c[b] >= k) break;
                    cum += c[b];
                }
                s_k = k - cum;
                s_prefix = prefix | (static_cast<unsigned>(8 * threadIdx.x + b) << shift);
                s_mask = mask | (255u << shift);
            }
        }
        __syncthreads();
    }
    const float thresh = key_value(s_prefix);
    if (threadIdx.x == 0 && sa.thresh_out) *sa.thresh_out = thresh;
    if (sa.tm_out) {
#pragma unroll 2
        for (int i = threadIdx.x; i < n; i += nthreads) {
            const float v = cached ? key_value(keys[i]) : __ldcg(act + i);
            const float a = sa.tm_in ? sa.tm_in[i] * v : v;
            sa.tm_out[i] = (a > thresh) ? 1 : 0;
        }
    }
}

}  // namespace udape
