// pipeline.cuh — TMA-staged plane pipeline for the HBM-bound streaming kernels (sm_100a).
//
// Every hot-path reduction reads whole (n,c) / (b,k) planes of a few KB exactly once.  A
// persistent CTA (one per SM) owns the planes  blockIdx.x, blockIdx.x + gridDim.x, ...  and
// moves them through a ring of shared-memory stages:
//
//   producer  (the last warp, one lane per stage): waits for its stage to be released, arms its
//             `full` mbarrier with the byte count and issues one `cp.async.bulk` (the 1-D TMA
//             bulk copy, UBLKCP in SASS) per operand plane — no registers, no per-thread load
//             instructions, up to kMaxStages planes (128-192 KB) in flight per SM;
//   consumers (all other warps; count chosen at launch): every stage is bound to one warp; it waits on the
//             stage's `full` mbarrier, reduces the plane out of shared memory with 128-bit
//             LDS + warp shuffles / redux.sync (no block barrier anywhere), writes its results,
//             and releases the stage through the `empty` mbarrier.
//
// The bulk copy needs 16-byte aligned global addresses and sizes; callers fall back to the
// generic (register-staged) kernels for planes that do not qualify.
#pragma once

#include "common.cuh"

namespace udape {

constexpr int kConsumerWarps = 12;                       // most consumer warps per CTA
constexpr int kPipeThreads = (kConsumerWarps + 1) * 32;  // + the producer warp = 416 threads (<= 152 regs each)
constexpr int kMaxStages = 32;  // one producer lane per stage
constexpr int kPipeSmemBudget = 200 * 1024;  // of the 227 KB a CTA may own on sm_100a

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_addr(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

__device__ __forceinline__ uint4 lds128(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ uint2 lds64(const void* p) { return *reinterpret_cast<const uint2*>(p); }

// NaN-propagating maximum (FMNMX.NAN / FMNMX3.NAN) and its warp-wide form (CREDUX.MAX.F32.NAN):
// numpy/torch argmax treat NaN as the maximum, so "max is NaN" <=> "the plane holds a NaN".
__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float warp_max_nan(float v) {
    float r;
    asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}

struct PipeBarriers {
    uint64_t full[kMaxStages];
    uint64_t empty[kMaxStages];
};

// number of items (planes) this CTA owns
__device__ __forceinline__ int pipe_items(int64_t planes) {
    const int64_t first = blockIdx.x;
    return first < planes ? static_cast<int>((planes - first + gridDim.x - 1) / gridDim.x) : 0;
}

__device__ __forceinline__ void pipe_init(PipeBarriers& b, int stages) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&b.full[s], 1);   // the producer's expect_tx arrival (+ the copies' bytes)
            mbar_init(&b.empty[s], 1);  // lane 0 of the consuming warp
        }
        mbar_init_fence();
    }
    __syncthreads();
}

// Producer side: call from every lane of the producer warp.  Lane l owns stage l (stages <= 32):
// it fills that stage with items l, l+stages, l+2*stages, ... so the per-copy issue latency
// (empty-wait -> expect_tx -> cp.async.bulk) of different stages overlaps.  A stage is only ever
// touched by its owner lane, which keeps the mbarrier parity protocol one phase deep.
// `issue(item, stage, full_barrier)` arms the barrier and issues the bulk copies of that item.
template <typename Issue>
__device__ __forceinline__ void pipe_produce(PipeBarriers& b, int stages, int n_items, Issue issue) {
    const int stage = threadIdx.x & 31;
    if (stage >= stages) return;
    uint32_t phase = 0;
    for (int i = stage; i < n_items; i += stages) {
        if (i >= stages) mbar_wait(&b.empty[stage], phase ^ 1u);  // released after its previous fill
        issue(i, stage, &b.full[stage]);
        phase ^= 1u;
    }
}

// Consumer side: call from every lane of consumer warp `warp` (0 <= warp < consumers, where
// consumers = blockDim.x/32 - 1; the last warp is the producer).
// Stage s is bound to warp s % consumers for the whole kernel: all fills of a stage are
// consumed by the same warp, in order, so a waiter is never more than one mbarrier phase away
// from the barrier's current phase (the parity protocol cannot tell two phases apart).
// `consume(item, stage)` runs with the stage's data visible; the stage is then released and
// `after_release(item)` runs while the producer is already refilling it.
template <typename Consume, typename After>
__device__ __forceinline__ void pipe_consume(PipeBarriers& b, int stages, int n_items, int warp, Consume consume,
                                             After after_release) {
    const int consumers = (blockDim.x >> 5) - 1;
    uint32_t phase = 0;
    for (int base = 0; base < n_items; base += stages, phase ^= 1u) {
        for (int stage = warp; stage < stages; stage += consumers) {
            const int i = base + stage;
            if (i >= n_items) break;
            mbar_wait(&b.full[stage], phase);
            consume(i, stage);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&b.empty[stage]);
            after_release(i);  // work that no longer needs the staged data overlaps the refill
        }
    }
}
template <typename Consume>
__device__ __forceinline__ void pipe_consume(PipeBarriers& b, int stages, int n_items, int warp, Consume consume) {
    pipe_consume(b, stages, n_items, warp, consume, [](int) {});
}

// host: pipeline geometry for planes of `plane_bytes` (all operands of one plane together).
// Small planes are grouped: an item = `group` consecutive planes (contiguous in memory, so still
// one bulk copy per operand), sized towards kPipeItemTarget so that ~200 KB are in flight per SM
// with at most kMaxStages barriers.  stages == 0: does not fit, use the generic kernel.
constexpr int kPipeItemTarget = 8 * 1024;
struct PipeGeom {
    int group;    // planes per item
    int stages;   // ring depth
    int threads;  // CTA size = (consumer warps + 1 producer warp) * 32
};
int pipe_stage_cap();            // api.cu: kMaxStages, UDAPE_PIPE_STAGES overrides (tuning)
inline PipeGeom pipe_geometry(int64_t plane_bytes, int64_t planes, int max_group = 8, int stages_per_warp = 1) {
    PipeGeom g = {1, 0, 0};
    int warps = kConsumerWarps;
    if (plane_bytes <= 0 || plane_bytes > kPipeSmemBudget / 2) return g;
    int64_t grp = kPipeItemTarget / plane_bytes;
    if (grp > max_group) grp = max_group;
    // keep every SM busy: never group so much that there are fewer items than ~4 per SM
    while (grp > 1 && planes / grp < 4 * 148) grp >>= 1;
    if (grp < 1) grp = 1;
    int64_t s = kPipeSmemBudget / (grp * plane_bytes);
    if (s > pipe_stage_cap()) s = pipe_stage_cap();
    // equal number of stages per consumer warp: the most warps (<= kConsumerWarps) that divide the
    // stage count, trimming the ring by up to 3 stages if that buys a better divisor
    int best_s = static_cast<int>(s), best_w = 1;
    for (int cand = static_cast<int>(s); cand >= 2 && cand > static_cast<int>(s) - 4; --cand) {
        int wdiv = 1;
        for (int d = warps; d >= 1; --d)
            if (cand % d == 0 && cand / d >= stages_per_warp) { wdiv = d; break; }
        if (wdiv > best_w) { best_w = wdiv; best_s = cand; }
    }
    s = best_s;
    warps = best_w;
    g.group = static_cast<int>(grp);
    g.stages = static_cast<int>(s);
    g.threads = (warps + 1) * 32;
    return g;
}
// host: opt a kernel instantiation into `bytes` of dynamic shared memory (> 48 KB needs the
// attribute); remembered per device so that steady-state launches (and CUDA-graph capture) make
// no extra runtime call.  The kernel is a non-type template argument, so every instantiation
// owns its own cache:  pipe_reserve_smem<&my_kernel<T, J>>(bytes).
template <auto kernel>
inline cudaError_t pipe_reserve_smem(size_t bytes) {
    static int reserved[64] = {0};  // benign race: monotone, idempotent
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && reserved[dev] >= static_cast<int>(bytes)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e == cudaSuccess && dev >= 0 && dev < 64) reserved[dev] = static_cast<int>(bytes);
    return e;
}

// host: persistent grid — one CTA per SM, fewer when there are not enough planes to go round
int pipe_grid(int64_t planes);   // api.cu (queries the SM count once per device)
bool pipe_enabled();             // api.cu: UDAPE_NO_TMA=1 forces the generic kernels (debugging)

}  // namespace udape
