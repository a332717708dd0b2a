// dp.cu — the data-parallel tail of the train step over peer memory (NVLink 5 / NVSwitch).
//
// Replaces what nn.DataParallel + the optimizers do after backward in the reference
// (train_human.py:145-148 replicate / reduce-add of the student gradients onto GPU 0, :436-438
// scaler.step(stu_optimizer); tea_optimizer.step(), and the host-side PCK bookkeeping of :443-445)
// for one process per GPU.  Every rank holds, in peer-mapped memory,
//     grads[P]   its flat gradient bucket (what backward wrote),
//     params[P]  its flat student parameters,
//     pad        a small signal pad the other ranks write flags into,
// and the step is three kernels that READ / WRITE THE OTHER GPUS' MEMORY DIRECTLY (ld/st on mapped peer
// pointers; NVSwitch gives every pair the full link) instead of an all-reduce followed by a replicated
// optimizer pass:
//   K1 reduce_scatter : rank r sums slice r of every rank's gradient bucket in rank order (deterministic,
//                       identical on every world size that is a power of two after the exact 1/W scaling),
//                       checks the result for non-finite values (GradScaler's found_inf — free here,
//                       the data is in registers) and keeps the averaged slice in place;
//   K2 shard_step     : unscale + Adam | SGD on slice r only — optimizer state is sharded, each rank streams
//                       1/W of exp_avg / exp_avg_sq (ZeRO-1 layout; results are bit-identical to the
//                       replicated update because the update is elementwise);
//   K3 gather_ema     : all-gather of the updated slices by peer loads fused with the teacher EMA
//                       (utils.py:21-25): the pulled student value is written to the local replica and folded
//                       into the teacher in the same pass.
// Bytes over NVLink per rank and step: (W-1)/W * P * 4 in K1 and again in K3 — what a ring all-reduce
// moves — but HBM traffic drops from ~13 P*4 (NCCL in + out, grad check, replicated 9-pass step) to
// ~(4 + 8/W) P*4.  Cross-rank ordering uses monotonic step numbers in the signal pads (release / acquire
// at system scope); waiting is done by single-CTA kernels so that a rank that is late never parks a
// grid of spinning CTAs on the other GPUs, and every wait is bounded (timeout -> error word, no hang).
#include <cstring>

#include "optim.cuh"

namespace udape {

constexpr int kDpThreads = 256;
constexpr int kDpChunk = 4096;    // elements per CTA in K2 / K3 (256 threads x 4 x 128-bit)
constexpr int kRsChunk = 8192;    // elements per CTA in K1 (16 peer loads in flight per thread)

// ---- system-scope access ---------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 128-bit load of data another GPU owns.  A plain (weak) load that does not allocate in L1: the data was complete
// before this kernel started (the flag wait is an earlier kernel of the same stream, and L1 is invalidated at
// kernel boundaries; peer addresses bypass the local L2), so no stale line can serve it.  (ld.relaxed.sys moved
// the same bytes ~20 % slower: 583 GB/s vs the 770 GB/s a peer copy reaches.)
__device__ __forceinline__ uint4 ld_peer(const void* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// pad layout (uint32 words), see include/udape.h
constexpr int kPadInf = UDAPE_DP_PHASES * UDAPE_DP_MAX_RANKS;         // inf[src]
constexpr int kPadErr = kPadInf + UDAPE_DP_MAX_RANKS;                 // err (local)
constexpr int kPadCounts = 64;                                        // counts[parity][src][64]

// Spin until *flag has reached step number e (monotonic, wrap-safe).  Bounded: after timeout_ns the local
// error word is set and the wait gives up, so a missing rank costs a wrong result and an error, not a hang.
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t e, uint64_t timeout_ns, uint32_t* err, uint32_t code) {
    if (static_cast<int32_t>(ld_acquire_sys(flag) - e) >= 0) return true;
    const uint64_t t0 = global_ns();
    unsigned spins = 0;
    while (static_cast<int32_t>(ld_acquire_sys(flag) - e) < 0) {
        if ((++spins & 63u) == 0) {
            __nanosleep(64);
            if (timeout_ns && global_ns() - t0 > timeout_ns) {
                atomicExch(err, code);
                return false;
            }
        }
    }
    return true;
}

// post step number e of `phase` into every rank's pad (threads 0..W-1 of the calling CTA)
__device__ __forceinline__ void post_all(const udape_dp_peers& pr, int phase, uint32_t e) {
    if (threadIdx.x < static_cast<unsigned>(pr.world)) {
        __threadfence_system();
        st_release_sys(pr.pads[threadIdx.x] + phase * UDAPE_DP_MAX_RANKS + pr.rank, e);
    }
}

// ---- barrier / wait (single CTA, one thread per rank) ---------------------------------------------
__global__ void __launch_bounds__(32)
dp_barrier_kernel(udape_dp_peers pr, int phase, const uint32_t* __restrict__ epoch_dev, unsigned long long timeout_ns) {
    const uint32_t e = *epoch_dev + 1u;
    post_all(pr, phase, e);
    if (threadIdx.x < static_cast<unsigned>(pr.world))
        wait_flag(pr.pads[pr.rank] + phase * UDAPE_DP_MAX_RANKS + threadIdx.x, e, timeout_ns, pr.pads[pr.rank] + kPadErr,
                  1u + phase);
}

__global__ void __launch_bounds__(32)
dp_wait_kernel(udape_dp_peers pr, int phase, const uint32_t* __restrict__ epoch_dev, float* __restrict__ found_inf,
               unsigned long long timeout_ns) {
    const uint32_t e = *epoch_dev + 1u;
    uint32_t bad = 0;
    if (threadIdx.x < static_cast<unsigned>(pr.world)) {
        uint32_t* pad = pr.pads[pr.rank];
        wait_flag(pad + phase * UDAPE_DP_MAX_RANKS + threadIdx.x, e, timeout_ns, pad + kPadErr, 1u + phase);
        if (found_inf) bad = ld_relaxed_sys(pad + kPadInf + threadIdx.x);
    }
    const uint32_t any = __reduce_or_sync(0xffffffffu, bad);
    if (found_inf && threadIdx.x == 0) *found_inf = any ? 1.0f : 0.0f;
}

// ---- K1: reduce-scatter of the gradient buckets, averaged, with the non-finite check ------------------
// Rank r owns elements [lo, lo + n).  UN vectors x world ranks = 16 x 128-bit loads in flight per thread.
template <int WMAX>
__global__ void __launch_bounds__(kDpThreads)
dp_reduce_scatter_kernel(udape_dp_peers pr, long long lo, long long n, float inv_world, float* __restrict__ reduced,
                         const uint32_t* __restrict__ epoch_dev, uint32_t* __restrict__ ws) {
    constexpr int UN = 16 / WMAX;
    const int world = pr.world;
    const long long cta_lo = static_cast<long long>(blockIdx.x) * kRsChunk;
    const long long rem = n - cta_lo;
    const int nvec = static_cast<int>((rem < kRsChunk ? rem : kRsChunk) >> 2);
    const long long vbase = (lo + cta_lo) >> 2;
    uint32_t bad = 0;
    for (int base = 0; base < nvec; base += kDpThreads * UN) {
        uint4 x[WMAX][UN];
#pragma unroll
        for (int q = 0; q < WMAX; ++q) {
            if (q < world) {
                const uint4* src = reinterpret_cast<const uint4*>(pr.grads[q]) + vbase;
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = base + u * kDpThreads + threadIdx.x;
                    if (i < nvec) x[q][u] = ld_peer(src + i);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = base + u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                float acc[4], t[4];
                unpack16<float>(x[0][u], acc);
#pragma unroll
                for (int q = 1; q < WMAX; ++q) {       // rank order: ((g0 + g1) + g2) + ...
                    if (q < world) {
                        unpack16<float>(x[q][u], t);
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[k] = __fadd_rn(acc[k], t[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[k] = __fmul_rn(acc[k], inv_world);
                    bad |= (__float_as_uint(acc[k]) & 0x7f800000u) == 0x7f800000u;
                }
                stg_plain(reinterpret_cast<uint4*>(reduced) + (cta_lo >> 2) + i, pack16<float>(acc));
            }
        }
    }
    const int any = __syncthreads_or(static_cast<int>(bad));
    if (threadIdx.x == 0 && any) atomicOr(ws + 1, 1u);
    if (last_block_done(ws, gridDim.x)) {
        // every rank learns (a) that this rank no longer reads its bucket, (b) whether slice r is finite
        const uint32_t e = *epoch_dev + 1u;
        const uint32_t f = *reinterpret_cast<volatile uint32_t*>(ws + 1);
        if (threadIdx.x < static_cast<unsigned>(world)) st_relaxed_sys(pr.pads[threadIdx.x] + kPadInf + pr.rank, f);
        post_all(pr, UDAPE_DP_REDUCED, e);
        __syncthreads();
        if (threadIdx.x == 0) ws[1] = 0u;
    }
}

// ---- K2: the optimizer on this rank's slice -------------------------------------------------------------
// ALGO 0: Adam, 1: SGD, 2: SGD + Nesterov.  m / v are this rank's state shards (index 0 = element lo).
template <int ALGO>
__global__ void __launch_bounds__(kDpThreads)
dp_shard_step_kernel(udape_dp_peers pr, long long lo, long long n, udape_opt_hyper h, const float* __restrict__ lr_dev,
                     const float* __restrict__ grad_scale, const float* __restrict__ found_inf,
                     int32_t* __restrict__ step_dev, const float* __restrict__ reduced, float* __restrict__ m,
                     float* __restrict__ v, const uint32_t* __restrict__ epoch_dev, uint32_t* __restrict__ ticket) {
    __shared__ OptScalars sc;
    __shared__ int first_s;
    if (threadIdx.x == 0) {
        sc = make_opt_scalars<ALGO>(h, lr_dev, grad_scale, found_inf, step_dev);
        first_s = *step_dev == 0;      // the flat bucket gives every parameter a gradient from the first step on
    }
    __syncthreads();
    const OptScalars s = sc;
    const bool first = first_s != 0;
    const bool has_m = m != nullptr;
    if (!s.skip) {
        const long long cta_lo = static_cast<long long>(blockIdx.x) * kDpChunk;
        const long long rem = n - cta_lo;
        const int nvec = static_cast<int>((rem < kDpChunk ? rem : kDpChunk) >> 2);
        uint4* p4 = reinterpret_cast<uint4*>(pr.params[pr.rank]) + ((lo + cta_lo) >> 2);
        const uint4* g4 = reinterpret_cast<const uint4*>(reduced) + (cta_lo >> 2);
        uint4* m4 = has_m ? reinterpret_cast<uint4*>(m) + (cta_lo >> 2) : nullptr;
        uint4* v4 = ALGO == 0 ? reinterpret_cast<uint4*>(v) + (cta_lo >> 2) : nullptr;
        uint4 pv[4], gv[4], mv[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                pv[u] = ldg_cached(p4 + i);
                gv[u] = ldg_cached(g4 + i);
                if (has_m) mv[u] = ldg_cached(m4 + i);
                if (ALGO == 0) vv[u] = ldg_cached(v4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                float fp[4], fg[4], fm[4], fv[4];
                unpack16<float>(pv[u], fp);
                unpack16<float>(gv[u], fg);
                if (has_m) unpack16<float>(mv[u], fm);
                if (ALGO == 0) unpack16<float>(vv[u], fv);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (ALGO == 0) adam_elem(fp[e], fg[e], fm[e], fv[e], s);
                    else sgd_elem<ALGO == 2>(fp[e], fg[e], fm[e], s, has_m, first);
                }
                stg_plain(p4 + i, pack16<float>(fp));
                if (has_m) stg_plain(m4 + i, pack16<float>(fm));
                if (ALGO == 0) stg_plain(v4 + i, pack16<float>(fv));
            }
        }
    }
    // The updated slice must be visible to the peers' loads before they are told it is ready.  Every CTA's stores
    // are ordered before its ticket by a device-scope fence (they are in this GPU's L2, which is where peer loads
    // are served from); the last CTA, having observed every ticket, fences at SYSTEM scope (post_all) before the
    // release store of the flag — fence cumulativity carries the other CTAs' stores along.  (A system-scope
    // fence in every CTA cost ~110 us of the 230 us this kernel took at N=2.)
    __syncthreads();
    if (last_block_done(ticket, gridDim.x)) {
        const uint32_t e = *epoch_dev + 1u;
        post_all(pr, UDAPE_DP_PARAMS, e);
        if (threadIdx.x == 0 && !s.skip) *step_dev = *step_dev + 1;
    }
}

// ---- K3: all-gather of the updated slices by peer loads, fused with the teacher EMA ----------------------
__global__ void __launch_bounds__(kDpThreads)
dp_gather_ema_kernel(udape_dp_peers pr, long long n_total, long long shard_elems, float* __restrict__ teacher,
                     float ema_a, float ema_b, const float* __restrict__ found_inf, uint32_t* __restrict__ epoch_dev,
                     uint32_t* __restrict__ ticket) {
    const bool skip = found_inf && *found_inf != 0.0f;     // nothing was updated anywhere: only the EMA runs
    // Consecutive CTAs take chunks of DIFFERENT owners (chunk j of owner 0, chunk j of owner 1, ...), so at any
    // moment this rank pulls from every peer at once — every link of the switch carries its share and the local
    // slice's HBM-bound chunks overlap the NVLink-bound ones.  (In plain chunk order all ranks would pull from
    // the same owner at the same time and share that one GPU's egress.)
    // The grid is world x (chunks per slice); the few CTAs past the end of the short last slice only take a ticket.
    const int world = pr.world;
    const long long per = shard_elems / kDpChunk;                        // chunks per slice
    const int owner = static_cast<int>(blockIdx.x % world);
    const long long cta_lo = (static_cast<long long>(owner) * per + blockIdx.x / world) * kDpChunk;
    const bool remote = owner != pr.rank && !skip;
    const long long rem = n_total - cta_lo;
    const int nvec = rem > 0 ? static_cast<int>((rem < kDpChunk ? rem : kDpChunk) >> 2) : 0;
    if (nvec > 0 && (remote || teacher)) {
        const uint4* src = reinterpret_cast<const uint4*>(remote ? pr.params[owner] : pr.params[pr.rank]) + (cta_lo >> 2);
        uint4* dst = reinterpret_cast<uint4*>(pr.params[pr.rank]) + (cta_lo >> 2);
        uint4* t4 = teacher ? reinterpret_cast<uint4*>(teacher) + (cta_lo >> 2) : nullptr;
        uint4 sv[4], tv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                sv[u] = remote ? ld_peer(src + i) : ldg_cached(src + i);
                if (t4) tv[u] = ldg_cached(t4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                if (remote) stg_plain(dst + i, sv[u]);
                if (t4) {
                    float fs[4], ft[4];
                    unpack16<float>(sv[u], fs);
                    unpack16<float>(tv[u], ft);
#pragma unroll
                    for (int e = 0; e < 4; ++e) ft[e] = ema_fold(ft[e], fs[e], ema_a, ema_b);
                    stg_plain(t4 + i, pack16<float>(ft));
                }
            }
        }
    }
    if (last_block_done(ticket, gridDim.x) && threadIdx.x == 0) *epoch_dev = *epoch_dev + 1u;   // the step is over on this rank
}

// ---- integer all-reduce of the PCK counts (<= 64 int32) in one single-CTA kernel --------------------------
__global__ void __launch_bounds__(64)
dp_counts_kernel(udape_dp_peers pr, const int32_t* __restrict__ counts, int n, int32_t* __restrict__ out,
                 uint32_t* __restrict__ epoch_dev, unsigned long long timeout_ns) {
    const uint32_t e = *epoch_dev + 1u;
    const int slot = kPadCounts + static_cast<int>(e & 1u) * UDAPE_DP_MAX_RANKS * 64;   // double-buffered by step parity
    const int i = threadIdx.x;
    if (i < n) {
        const uint32_t v = static_cast<uint32_t>(counts[i]);
        for (int q = 0; q < pr.world; ++q) st_relaxed_sys(pr.pads[q] + slot + pr.rank * 64 + i, v);
    }
    __syncthreads();
    post_all(pr, UDAPE_DP_COUNTS, e);
    uint32_t* pad = pr.pads[pr.rank];
    if (i < pr.world) wait_flag(pad + UDAPE_DP_COUNTS * UDAPE_DP_MAX_RANKS + i, e, timeout_ns, pad + kPadErr, 1u + UDAPE_DP_COUNTS);
    __syncthreads();
    if (i < n) {
        int32_t s = 0;
        for (int q = 0; q < pr.world; ++q) s += static_cast<int32_t>(ld_relaxed_sys(pad + slot + q * 64 + i));
        out[i] = s;
    }
    if (i == 0) *epoch_dev = e;
}

// CUDA loads kernels lazily, and loading one may wait for the kernels that are running.  A rank's barrier
// kernel can be spinning on a peer while the host goes on to launch the next kernel of the chain: if that
// launch had to load its code first, the host would stall until the spin ends — and with several ranks driven
// by ONE host thread (PeerGroup.virtual) the peer's kernels are launched by that same thread, so the spin would
// only end by its timeout.  Every entry point therefore makes sure, once per device, that all kernels of this
// file are resident before anything is launched.
template <typename K> static cudaError_t preload(K kernel) {
    cudaFuncAttributes a;
    return cudaFuncGetAttributes(&a, kernel);
}
static int ensure_loaded() {
    static bool done[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return UDAPE_OK;
    if (done[dev]) return UDAPE_OK;
    cudaError_t e = cudaSuccess;
    auto acc = [&e](cudaError_t r) { if (e == cudaSuccess) e = r; };
    acc(preload(dp_barrier_kernel));
    acc(preload(dp_wait_kernel));
    acc(preload(dp_reduce_scatter_kernel<2>));
    acc(preload(dp_reduce_scatter_kernel<4>));
    acc(preload(dp_reduce_scatter_kernel<8>));
    acc(preload(dp_shard_step_kernel<0>));
    acc(preload(dp_shard_step_kernel<1>));
    acc(preload(dp_shard_step_kernel<2>));
    acc(preload(dp_gather_ema_kernel));
    acc(preload(dp_counts_kernel));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_dp: loading the kernels failed: %s", cudaGetErrorString(e));
    done[dev] = true;
    return UDAPE_OK;
}

static int check_peers(const udape_dp_peers* p, const char* what, bool need_data) {
    if (int e = ensure_loaded()) return e;
    UDAPE_REQUIRE(p, UDAPE_ERR_NULL, "%s: peers is NULL", what);
    UDAPE_REQUIRE(p->world >= 1 && p->world <= UDAPE_DP_MAX_RANKS && p->rank >= 0 && p->rank < p->world, UDAPE_ERR_ARG,
                  "%s: bad rank %d / world %d (world <= %d)", what, (int)p->rank, (int)p->world, UDAPE_DP_MAX_RANKS);
    for (int q = 0; q < p->world; ++q) {
        UDAPE_REQUIRE(p->pads[q], UDAPE_ERR_NULL, "%s: pad of rank %d is NULL", what, q);
        if (need_data) {
            UDAPE_REQUIRE(p->grads[q] && p->params[q], UDAPE_ERR_NULL, "%s: grads / params of rank %d is NULL", what, q);
            UDAPE_REQUIRE(aligned16(p->grads[q]) && aligned16(p->params[q]), UDAPE_ERR_ALIGN, "%s: buffers must be 16-byte aligned", what);
        }
    }
    return UDAPE_OK;
}

}  // namespace udape

using namespace udape;

extern "C" int64_t udape_dp_shard_elems(int64_t n_total, int world) {
    if (n_total < 0 || world < 1 || world > UDAPE_DP_MAX_RANKS) return fail(UDAPE_ERR_ARG, "udape_dp_shard_elems: bad n_total / world");
    const int64_t chunks = (n_total + kDpChunk - 1) / kDpChunk;
    const int64_t per = (chunks + world - 1) / world;
    return (per > 0 ? per : 1) * kDpChunk;
}

static inline void shard_range(const udape_dp_peers* p, int64_t n_total, int64_t* lo, int64_t* n) {
    const int64_t s = udape_dp_shard_elems(n_total, p->world);
    *lo = s * p->rank;
    const int64_t hi = *lo + s < n_total ? *lo + s : n_total;
    *n = hi > *lo ? hi - *lo : 0;
}

extern "C" int udape_dp_barrier(const udape_dp_peers* peers, int phase, const uint32_t* epoch_dev,
                                uint64_t timeout_ns, void* stream) {
    if (int e = check_peers(peers, "udape_dp_barrier", false)) return e;
    UDAPE_REQUIRE(phase >= 0 && phase < UDAPE_DP_PHASES && epoch_dev, UDAPE_ERR_ARG, "udape_dp_barrier: bad phase / NULL epoch");
    dp_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(*peers, phase, epoch_dev, timeout_ns);
    return check_launch("udape_dp_barrier");
}

extern "C" int udape_dp_wait(const udape_dp_peers* peers, int phase, const uint32_t* epoch_dev, float* found_inf,
                             uint64_t timeout_ns, void* stream) {
    if (int e = check_peers(peers, "udape_dp_wait", false)) return e;
    UDAPE_REQUIRE(phase >= 0 && phase < UDAPE_DP_PHASES && epoch_dev, UDAPE_ERR_ARG, "udape_dp_wait: bad phase / NULL epoch");
    dp_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(*peers, phase, epoch_dev, found_inf, timeout_ns);
    return check_launch("udape_dp_wait");
}

extern "C" int udape_dp_reduce_scatter(const udape_dp_peers* peers, int64_t n_total, float* reduced,
                                       const uint32_t* epoch_dev, uint32_t* ws, void* stream) {
    if (int e = check_peers(peers, "udape_dp_reduce_scatter", true)) return e;
    UDAPE_REQUIRE(reduced && epoch_dev && ws, UDAPE_ERR_NULL, "udape_dp_reduce_scatter: reduced / epoch / ws is NULL");
    UDAPE_REQUIRE(aligned16(reduced), UDAPE_ERR_ALIGN, "udape_dp_reduce_scatter: reduced must be 16-byte aligned");
    UDAPE_REQUIRE(n_total > 0 && (n_total & 3) == 0, UDAPE_ERR_SHAPE, "udape_dp_reduce_scatter: n_total must be a positive multiple of 4");
    int64_t lo, n;
    shard_range(peers, n_total, &lo, &n);
    cudaStream_t st = as_stream(stream);
    // a rank whose slice is empty (tiny buckets) still takes part in the signalling: one CTA, no elements
    const unsigned grid = static_cast<unsigned>(n > 0 ? (n + kRsChunk - 1) / kRsChunk : 1);
    const float inv = 1.0f / static_cast<float>(peers->world);
    if (peers->world <= 2) dp_reduce_scatter_kernel<2><<<grid, kDpThreads, 0, st>>>(*peers, lo, n, inv, reduced, epoch_dev, ws);
    else if (peers->world <= 4) dp_reduce_scatter_kernel<4><<<grid, kDpThreads, 0, st>>>(*peers, lo, n, inv, reduced, epoch_dev, ws);
    else dp_reduce_scatter_kernel<8><<<grid, kDpThreads, 0, st>>>(*peers, lo, n, inv, reduced, epoch_dev, ws);
    return check_launch("udape_dp_reduce_scatter");
}

extern "C" int udape_dp_shard_step(const udape_dp_peers* peers, int64_t n_total, int algo, const udape_opt_hyper* hyper,
                                   const float* lr_dev, const float* grad_scale, const float* found_inf,
                                   int32_t* step_dev, const float* reduced, float* state1, float* state2,
                                   const uint32_t* epoch_dev, uint32_t* ticket, void* stream) {
    if (int e = check_peers(peers, "udape_dp_shard_step", true)) return e;
    UDAPE_REQUIRE(hyper && step_dev && reduced && epoch_dev && ticket, UDAPE_ERR_NULL, "udape_dp_shard_step: hyper / step_dev / reduced / epoch / ticket is NULL");
    UDAPE_REQUIRE(aligned16(reduced), UDAPE_ERR_ALIGN, "udape_dp_shard_step: reduced must be 16-byte aligned");
    UDAPE_REQUIRE(n_total > 0 && (n_total & 3) == 0, UDAPE_ERR_SHAPE, "udape_dp_shard_step: n_total must be a positive multiple of 4");
    UDAPE_REQUIRE(algo == UDAPE_OPT_ADAM || algo == UDAPE_OPT_SGD, UDAPE_ERR_ARG, "udape_dp_shard_step: algo must be UDAPE_OPT_ADAM or UDAPE_OPT_SGD");
    if (algo == UDAPE_OPT_ADAM) {
        UDAPE_REQUIRE(state1 && state2, UDAPE_ERR_NULL, "udape_dp_shard_step: Adam needs exp_avg and exp_avg_sq shards");
        UDAPE_REQUIRE(hyper->beta1 >= 0.0 && hyper->beta1 < 1.0 && hyper->beta2 >= 0.0 && hyper->beta2 < 1.0 && hyper->eps >= 0.0,
                      UDAPE_ERR_ARG, "udape_dp_shard_step: Adam needs 0 <= beta < 1 and eps >= 0");
    } else {
        UDAPE_REQUIRE((hyper->beta1 == 0.0) == (state1 == nullptr), UDAPE_ERR_ARG, "udape_dp_shard_step: SGD momentum buffer given iff momentum != 0");
        UDAPE_REQUIRE(!(hyper->nesterov && (hyper->beta1 <= 0.0 || hyper->beta2 != 0.0)), UDAPE_ERR_ARG,
                      "udape_dp_shard_step: Nesterov momentum requires a momentum and zero dampening");
    }
    UDAPE_REQUIRE((!state1 || aligned16(state1)) && (!state2 || aligned16(state2)), UDAPE_ERR_ALIGN, "udape_dp_shard_step: state shards must be 16-byte aligned");
    int64_t lo, n;
    shard_range(peers, n_total, &lo, &n);
    cudaStream_t st = as_stream(stream);
    const unsigned grid = static_cast<unsigned>(n > 0 ? (n + kDpChunk - 1) / kDpChunk : 1);
    if (algo == UDAPE_OPT_ADAM)
        dp_shard_step_kernel<0><<<grid, kDpThreads, 0, st>>>(*peers, lo, n, *hyper, lr_dev, grad_scale, found_inf, step_dev, reduced, state1, state2, epoch_dev, ticket);
    else if (hyper->nesterov)
        dp_shard_step_kernel<2><<<grid, kDpThreads, 0, st>>>(*peers, lo, n, *hyper, lr_dev, grad_scale, found_inf, step_dev, reduced, state1, state2, epoch_dev, ticket);
    else
        dp_shard_step_kernel<1><<<grid, kDpThreads, 0, st>>>(*peers, lo, n, *hyper, lr_dev, grad_scale, found_inf, step_dev, reduced, state1, state2, epoch_dev, ticket);
    return check_launch("udape_dp_shard_step");
}

extern "C" int udape_dp_gather_ema(const udape_dp_peers* peers, int64_t n_total, float* teacher, float ema_a, float ema_b,
                                   const float* found_inf, uint32_t* epoch_dev, uint32_t* ticket, void* stream) {
    if (int e = check_peers(peers, "udape_dp_gather_ema", true)) return e;
    UDAPE_REQUIRE(epoch_dev && ticket, UDAPE_ERR_NULL, "udape_dp_gather_ema: epoch / ticket is NULL");
    UDAPE_REQUIRE(n_total > 0 && (n_total & 3) == 0, UDAPE_ERR_SHAPE, "udape_dp_gather_ema: n_total must be a positive multiple of 4");
    UDAPE_REQUIRE(!teacher || aligned16(teacher), UDAPE_ERR_ALIGN, "udape_dp_gather_ema: teacher must be 16-byte aligned");
    const int64_t shard = udape_dp_shard_elems(n_total, peers->world);
    const unsigned grid = static_cast<unsigned>(shard / kDpChunk * peers->world);
    dp_gather_ema_kernel<<<grid, kDpThreads, 0, as_stream(stream)>>>(*peers, n_total, shard, teacher, ema_a, ema_b, found_inf, epoch_dev, ticket);
    return check_launch("udape_dp_gather_ema");
}

extern "C" int udape_dp_allreduce_counts(const udape_dp_peers* peers, const int32_t* counts, int n, int32_t* out,
                                         uint32_t* epoch_dev, uint64_t timeout_ns, void* stream) {
    if (int e = check_peers(peers, "udape_dp_allreduce_counts", false)) return e;
    UDAPE_REQUIRE(counts && out && epoch_dev, UDAPE_ERR_NULL, "udape_dp_allreduce_counts: counts / out / epoch is NULL");
    UDAPE_REQUIRE(n >= 1 && n <= 64, UDAPE_ERR_SHAPE, "udape_dp_allreduce_counts: n=%d (1..64 int32: hits || valid of <= 32 joints)", n);
    dp_counts_kernel<<<1, 64, 0, as_stream(stream)>>>(*peers, counts, n, out, epoch_dev, timeout_ns);
    return check_launch("udape_dp_allreduce_counts");
}

// ---- peer-memory plumbing: one cudaMalloc'ed arena per rank, exported / opened through CUDA IPC -------------
extern "C" int udape_peer_alloc(size_t bytes, void** ptr) {
    UDAPE_REQUIRE(ptr && bytes > 0, UDAPE_ERR_ARG, "udape_peer_alloc: NULL ptr / zero bytes");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        return fail(static_cast<int>(e), "udape_peer_alloc(%zu): %s", bytes, cudaGetErrorString(e));
    }
    *ptr = p;
    return UDAPE_OK;
}

extern "C" int udape_peer_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? UDAPE_OK : fail(static_cast<int>(e), "udape_peer_free: %s", cudaGetErrorString(e));
}

extern "C" int udape_peer_export(const void* ptr, unsigned char* handle64) {
    UDAPE_REQUIRE(ptr && handle64, UDAPE_ERR_NULL, "udape_peer_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_peer_export: %s", cudaGetErrorString(e));
    memcpy(handle64, &h, 64);
    return UDAPE_OK;
}

extern "C" int udape_peer_open(const unsigned char* handle64, void** ptr) {
    UDAPE_REQUIRE(ptr && handle64, UDAPE_ERR_NULL, "udape_peer_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_peer_open: %s", cudaGetErrorString(e));
    *ptr = p;
    return UDAPE_OK;
}

extern "C" int udape_peer_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    return e == cudaSuccess ? UDAPE_OK : fail(static_cast<int>(e), "udape_peer_close: %s", cudaGetErrorString(e));
}
