// dp.cu — the data-parallel tail of the train step over peer memory (NVLink 5 / NVSwitch).
//
// Replaces what nn.DataParallel + the optimizers do after backward in the reference
// (train_human.py:145-148 replicate / reduce-add of the student gradients onto GPU 0, :436-438
// scaler.step(stu_optimizer); tea_optimizer.step(), and the host-side PCK bookkeeping of :443-445)
// for one process per GPU.  Every rank holds, in peer-mapped memory,
//     grads[P]   its flat gradient bucket (what backward wrote),
//     params[P]  its flat student parameters,
//     pad        a small signal pad the other ranks write flags into,
// and the step is two kernels that READ THE OTHER GPUS' MEMORY DIRECTLY (ld/st on mapped peer
// pointers; NVSwitch gives every pair the full link) instead of an all-reduce followed by a replicated
// optimizer pass:
//   K1 reduce_step : rank r sums slice r of every rank's gradient bucket in rank order (deterministic, and the same
//                    on every power-of-two world size after the exact 1/W scaling), tests the result for non-finite
//                    values (GradScaler's found_inf — free here, the data is in registers) and applies unscale +
//                    Adam | SGD to slice r in the same registers, speculatively, into a shadow slice; optimizer
//                    state is sharded (ZeRO-1 layout: each rank streams 1/W of exp_avg / exp_avg_sq; bit-identical
//                    to the replicated update because the update is elementwise);
//   K2 gather_ema  : once every rank's verdict is in: all-gather of the shadow slices by peer loads = the commit,
//                    fused with the teacher EMA (utils.py:21-25): the pulled value is written to the local replica
//                    and folded into the teacher in the same pass.
// Bytes over NVLink per rank and step: (W-1)/W * P * 4 in K1 and again in K2 — what a ring all-reduce moves —
// but HBM traffic drops from ~13 P*4 (NCCL in + out, grad check, replicated 9-pass step) to ~(4 + 8/W) P*4 and the
// step has two cross-rank waits.  Cross-rank ordering uses monotonic step numbers in the signal pads (release /
// acquire at system scope); waiting is done by single-CTA kernels so that a rank that is late never parks a grid
// of spinning CTAs on the other GPUs, and every wait is bounded (timeout -> error word, no hang).
#include <cstring>

#include "optim.cuh"

namespace udape {

constexpr int kDpThreads = 256;
constexpr int kDpChunk = 4096;    // elements per CTA in K2 / K3 (256 threads x 4 x 128-bit)
constexpr int kRsChunk = 8192;    // elements per CTA in K1

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
// the same bytes ~20 % slower.  What 128-bit loads can pull at all, two GPUs pulling from each other at once, is
// 621 GB/s whatever the grid / loads in flight — copy engines 718, cp.async.bulk 652, stores to the peer 671:
// tools/probe/peer_probe.py, profiles/r02au_peer_probe_n2.json.)
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

// ---- K1: reduce-scatter of the gradient buckets fused with the optimizer --------------------------------
// Rank r owns elements [lo, lo + n).  Per element: g = ((g_0 + g_1) + ...) * (1/W) summed in rank order from every
// rank's bucket (UN vectors x world ranks = 8 x 128-bit peer loads in flight per thread), non-finite test of g
// (GradScaler's found_inf — whether ANY rank's slice trips it is only known after the cross-rank wait that
// follows), and — without waiting for that verdict — the update itself, SPECULATIVELY: the new parameters go
// to this rank's `shadow` slice and the new optimizer state to the other half of the double-buffered state
// shards.  The gather kernel commits (copies shadow -> parameters everywhere, flips the state halves by
// advancing *step_dev) only if no rank saw a non-finite value; otherwise everything written here is simply
// never looked at again — exactly `scaler.step()` skipping the update.  Against reduce-scatter -> wait ->
// separate optimizer pass this saves the write + re-read of the reduced slice, one pass over p / m / v
// (742 MB per rank at N=2) and one cross-rank round trip.
// ALGO 0: Adam, 1: SGD, 2: SGD + Nesterov.  m / v: [2][S] shards; half (*step_dev & 1) is current.
template <int WMAX, int ALGO>
__global__ void __launch_bounds__(kDpThreads, 2)
dp_reduce_step_kernel(udape_dp_peers pr, long long lo, long long n, long long shard_elems, float inv_world,
                      udape_opt_hyper h, const float* __restrict__ lr_dev, const float* __restrict__ grad_scale,
                      const int32_t* __restrict__ step_dev, float* __restrict__ m, float* __restrict__ v,
                      const uint32_t* __restrict__ epoch_dev, uint32_t* __restrict__ ws) {
    // vectors per thread and pass: 8 gradient loads (peer) + 3 * UN local loads (p, m, v) in flight per thread,
    // at <= 128 registers so that two CTAs share an SM (loads of one overlap the arithmetic / stores of the other)
    constexpr int UN = 8 / WMAX;
    __shared__ OptScalars sc;
    __shared__ int cur_s;
    if (threadIdx.x == 0) {
        sc = make_opt_scalars<ALGO>(h, lr_dev, grad_scale, nullptr, step_dev);
        cur_s = *step_dev;
    }
    __syncthreads();
    const OptScalars s = sc;
    const bool first = cur_s == 0;        // the flat bucket gives every parameter a gradient from the first step on
    const long long half_in = static_cast<long long>(cur_s & 1) * shard_elems;
    const long long half_out = static_cast<long long>((cur_s & 1) ^ 1) * shard_elems;
    const bool has_m = m != nullptr;
    const int world = pr.world;
    const long long cta_lo = static_cast<long long>(blockIdx.x) * kRsChunk;
    const long long rem = n - cta_lo;
    const int nvec = rem > 0 ? static_cast<int>((rem < kRsChunk ? rem : kRsChunk) >> 2) : 0;
    const long long vbase = (lo + cta_lo) >> 2;
    const uint4* p4 = reinterpret_cast<const uint4*>(pr.params[pr.rank]) + vbase;
    uint4* out4 = reinterpret_cast<uint4*>(pr.shadow[pr.rank]) + (cta_lo >> 2);
    const uint4* mi = has_m ? reinterpret_cast<const uint4*>(m + half_in) + (cta_lo >> 2) : nullptr;
    uint4* mo = has_m ? reinterpret_cast<uint4*>(m + half_out) + (cta_lo >> 2) : nullptr;
    const uint4* vi = ALGO == 0 ? reinterpret_cast<const uint4*>(v + half_in) + (cta_lo >> 2) : nullptr;
    uint4* vo = ALGO == 0 ? reinterpret_cast<uint4*>(v + half_out) + (cta_lo >> 2) : nullptr;
    uint32_t bad = 0;
    for (int base = 0; base < nvec; base += kDpThreads * UN) {
        uint4 x[WMAX][UN], pv[UN], mv[UN], vv[UN];
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
                pv[u] = ldg_cached(p4 + i);
                if (has_m) mv[u] = ldg_stream(mi + i);
                if (ALGO == 0) vv[u] = ldg_stream(vi + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int i = base + u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                float g[4], t[4], fp[4], fm[4], fv[4];
                unpack16<float>(x[0][u], g);
#pragma unroll
                for (int q = 1; q < WMAX; ++q) {       // rank order: ((g0 + g1) + g2) + ...
                    if (q < world) {
                        unpack16<float>(x[q][u], t);
#pragma unroll
                        for (int k = 0; k < 4; ++k) g[k] = __fadd_rn(g[k], t[k]);
                    }
                }
                unpack16<float>(pv[u], fp);
                if (has_m) unpack16<float>(mv[u], fm);
                if (ALGO == 0) unpack16<float>(vv[u], fv);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    g[k] = __fmul_rn(g[k], inv_world);
                    bad |= (__float_as_uint(g[k]) & 0x7f800000u) == 0x7f800000u;
                    if (ALGO == 0) adam_elem(fp[k], g[k], fm[k], fv[k], s);
                    else sgd_elem<ALGO == 2>(fp[k], g[k], fm[k], s, has_m, first);
                }
                stg_plain(out4 + i, pack16<float>(fp));
                if (has_m) stg_stream(mo + i, pack16<float>(fm));
                if (ALGO == 0) stg_stream(vo + i, pack16<float>(fv));
            }
        }
    }
    const int any = __syncthreads_or(static_cast<int>(bad));
    if (threadIdx.x == 0 && any) atomicOr(ws + 1, 1u);
    // Every CTA's shadow stores are ordered before its ticket by a device-scope fence (they sit in this GPU's L2,
    // where peer loads are served from); the last CTA, having observed every ticket, fences at SYSTEM scope
    // (post_all) before the release store of the flag — fence cumulativity carries the other CTAs' stores along.
    // (A system-scope fence in every CTA cost ~110 us of a 230 us kernel at N=2.)
    if (last_block_done(ws, gridDim.x)) {
        // every rank learns (a) that this rank no longer reads its bucket, (b) that slice r's candidate values
        // are in place, (c) whether slice r's gradient is finite
        const uint32_t e = *epoch_dev + 1u;
        const uint32_t f = *reinterpret_cast<volatile uint32_t*>(ws + 1);
        if (threadIdx.x < static_cast<unsigned>(world)) st_relaxed_sys(pr.pads[threadIdx.x] + kPadInf + pr.rank, f);
        post_all(pr, UDAPE_DP_REDUCED, e);
        __syncthreads();
        if (threadIdx.x == 0) ws[1] = 0u;
    }
}

// ---- K2: all-gather of the updated slices by peer loads — the commit — fused with the teacher EMA ----------
// params_local[i] = shadow_owner(i)[i - owner's lo] for EVERY slice (the own one too: K1 left its candidates in
// the shadow), and teacher[i] = fl(fl(teacher[i]*a) + fl(params[i]*b)) in the same pass.  *found_inf != 0: nothing
// is committed anywhere, the EMA runs with the unchanged parameters — what scaler.step() + tea_optimizer.step()
// do.  The last CTA advances *step_dev (which also flips the optimizer-state halves) and *epoch_dev.
__global__ void __launch_bounds__(kDpThreads)
dp_gather_ema_kernel(udape_dp_peers pr, long long n_total, long long shard_elems, float* __restrict__ teacher,
                     float ema_a, float ema_b, const float* __restrict__ found_inf, int32_t* __restrict__ step_dev,
                     uint32_t* __restrict__ epoch_dev, uint32_t* __restrict__ ticket) {
    const bool skip = found_inf && *found_inf != 0.0f;
    // Consecutive CTAs take chunks of DIFFERENT owners (chunk j of owner 0, chunk j of owner 1, ...), so at any
    // moment this rank pulls from every peer at once — every link of the switch carries its share and the local
    // slice's HBM-bound chunks overlap the NVLink-bound ones.  (In plain chunk order all ranks would pull from
    // the same owner at the same time and share that one GPU's egress.)
    // The grid is world x (chunks per slice); the few CTAs past the end of the short last slice only take a ticket.
    const int world = pr.world;
    const long long per = shard_elems / kDpChunk;                        // chunks per slice
    const int owner = static_cast<int>(blockIdx.x % world);
    const long long in_slice = (blockIdx.x / world) * static_cast<long long>(kDpChunk);
    const long long cta_lo = static_cast<long long>(owner) * per * kDpChunk + in_slice;
    const long long rem = n_total - cta_lo;
    const int nvec = rem > 0 ? static_cast<int>((rem < kDpChunk ? rem : kDpChunk) >> 2) : 0;
    if (nvec > 0 && (!skip || teacher)) {
        const bool remote = owner != pr.rank;
        uint4* dst = reinterpret_cast<uint4*>(pr.params[pr.rank]) + (cta_lo >> 2);
        const uint4* src = skip ? dst : reinterpret_cast<const uint4*>(pr.shadow[owner]) + (in_slice >> 2);
        uint4* t4 = teacher ? reinterpret_cast<uint4*>(teacher) + (cta_lo >> 2) : nullptr;
        uint4 sv[4], tv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                sv[u] = (remote && !skip) ? ld_peer(src + i) : ldg_cached(src + i);
                if (t4) tv[u] = ldg_cached(t4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = u * kDpThreads + threadIdx.x;
            if (i < nvec) {
                if (!skip) stg_plain(dst + i, sv[u]);
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
    if (last_block_done(ticket, gridDim.x) && threadIdx.x == 0) {
        if (!skip) *step_dev = *step_dev + 1;      // the update counts — and the state halves flip
        *epoch_dev = *epoch_dev + 1u;              // the step is over on this rank
    }
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
    acc(preload(dp_reduce_step_kernel<2, 0>));
    acc(preload(dp_reduce_step_kernel<4, 0>));
    acc(preload(dp_reduce_step_kernel<8, 0>));
    acc(preload(dp_reduce_step_kernel<2, 1>));
    acc(preload(dp_reduce_step_kernel<4, 1>));
    acc(preload(dp_reduce_step_kernel<8, 1>));
    acc(preload(dp_reduce_step_kernel<2, 2>));
    acc(preload(dp_reduce_step_kernel<4, 2>));
    acc(preload(dp_reduce_step_kernel<8, 2>));
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
            UDAPE_REQUIRE(p->grads[q] && p->params[q] && p->shadow[q], UDAPE_ERR_NULL, "%s: grads / params / shadow of rank %d is NULL", what, q);
            UDAPE_REQUIRE(aligned16(p->grads[q]) && aligned16(p->params[q]) && aligned16(p->shadow[q]), UDAPE_ERR_ALIGN,
                          "%s: buffers must be 16-byte aligned", what);
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

template <int WMAX>
static void launch_reduce_step(int variant, unsigned grid, cudaStream_t st, const udape_dp_peers& pr, int64_t lo, int64_t n,
                               int64_t shard, float inv, const udape_opt_hyper& h, const float* lr_dev, const float* grad_scale,
                               const int32_t* step_dev, float* m, float* v, const uint32_t* epoch_dev, uint32_t* ws) {
    if (variant == 0) dp_reduce_step_kernel<WMAX, 0><<<grid, kDpThreads, 0, st>>>(pr, lo, n, shard, inv, h, lr_dev, grad_scale, step_dev, m, v, epoch_dev, ws);
    else if (variant == 1) dp_reduce_step_kernel<WMAX, 1><<<grid, kDpThreads, 0, st>>>(pr, lo, n, shard, inv, h, lr_dev, grad_scale, step_dev, m, v, epoch_dev, ws);
    else dp_reduce_step_kernel<WMAX, 2><<<grid, kDpThreads, 0, st>>>(pr, lo, n, shard, inv, h, lr_dev, grad_scale, step_dev, m, v, epoch_dev, ws);
}

extern "C" int udape_dp_reduce_step(const udape_dp_peers* peers, int64_t n_total, int algo, const udape_opt_hyper* hyper,
                                    const float* lr_dev, const float* grad_scale, const int32_t* step_dev, float* state1,
                                    float* state2, const uint32_t* epoch_dev, uint32_t* ws, void* stream) {
    if (int e = check_peers(peers, "udape_dp_reduce_step", true)) return e;
    UDAPE_REQUIRE(hyper && step_dev && epoch_dev && ws, UDAPE_ERR_NULL, "udape_dp_reduce_step: hyper / step_dev / epoch / ws is NULL");
    UDAPE_REQUIRE(n_total > 0 && (n_total & 3) == 0, UDAPE_ERR_SHAPE, "udape_dp_reduce_step: n_total must be a positive multiple of 4");
    UDAPE_REQUIRE(algo == UDAPE_OPT_ADAM || algo == UDAPE_OPT_SGD, UDAPE_ERR_ARG, "udape_dp_reduce_step: algo must be UDAPE_OPT_ADAM or UDAPE_OPT_SGD");
    if (algo == UDAPE_OPT_ADAM) {
        UDAPE_REQUIRE(state1 && state2, UDAPE_ERR_NULL, "udape_dp_reduce_step: Adam needs exp_avg and exp_avg_sq shards");
        UDAPE_REQUIRE(hyper->beta1 >= 0.0 && hyper->beta1 < 1.0 && hyper->beta2 >= 0.0 && hyper->beta2 < 1.0 && hyper->eps >= 0.0,
                      UDAPE_ERR_ARG, "udape_dp_reduce_step: Adam needs 0 <= beta < 1 and eps >= 0");
    } else {
        UDAPE_REQUIRE((hyper->beta1 == 0.0) == (state1 == nullptr), UDAPE_ERR_ARG, "udape_dp_reduce_step: SGD momentum buffer given iff momentum != 0");
        UDAPE_REQUIRE(!(hyper->nesterov && (hyper->beta1 <= 0.0 || hyper->beta2 != 0.0)), UDAPE_ERR_ARG,
                      "udape_dp_reduce_step: Nesterov momentum requires a momentum and zero dampening");
    }
    UDAPE_REQUIRE((!state1 || aligned16(state1)) && (!state2 || aligned16(state2)), UDAPE_ERR_ALIGN, "udape_dp_reduce_step: state shards must be 16-byte aligned");
    int64_t lo, n;
    shard_range(peers, n_total, &lo, &n);
    const int64_t shard = udape_dp_shard_elems(n_total, peers->world);
    cudaStream_t st = as_stream(stream);
    // a rank whose slice is empty (tiny buckets) still takes part in the signalling: one CTA, no elements
    const unsigned grid = static_cast<unsigned>(n > 0 ? (n + kRsChunk - 1) / kRsChunk : 1);
    const float inv = 1.0f / static_cast<float>(peers->world);
    const int variant = algo == UDAPE_OPT_ADAM ? 0 : (hyper->nesterov ? 2 : 1);
    if (peers->world <= 2) launch_reduce_step<2>(variant, grid, st, *peers, lo, n, shard, inv, *hyper, lr_dev, grad_scale, step_dev, state1, state2, epoch_dev, ws);
    else if (peers->world <= 4) launch_reduce_step<4>(variant, grid, st, *peers, lo, n, shard, inv, *hyper, lr_dev, grad_scale, step_dev, state1, state2, epoch_dev, ws);
    else launch_reduce_step<8>(variant, grid, st, *peers, lo, n, shard, inv, *hyper, lr_dev, grad_scale, step_dev, state1, state2, epoch_dev, ws);
    return check_launch("udape_dp_reduce_step");
}

extern "C" int udape_dp_gather_ema(const udape_dp_peers* peers, int64_t n_total, float* teacher, float ema_a, float ema_b,
                                   const float* found_inf, int32_t* step_dev, uint32_t* epoch_dev, uint32_t* ticket,
                                   void* stream) {
    if (int e = check_peers(peers, "udape_dp_gather_ema", true)) return e;
    UDAPE_REQUIRE(step_dev && epoch_dev && ticket, UDAPE_ERR_NULL, "udape_dp_gather_ema: step_dev / epoch / ticket is NULL");
    UDAPE_REQUIRE(n_total > 0 && (n_total & 3) == 0, UDAPE_ERR_SHAPE, "udape_dp_gather_ema: n_total must be a positive multiple of 4");
    UDAPE_REQUIRE(!teacher || aligned16(teacher), UDAPE_ERR_ALIGN, "udape_dp_gather_ema: teacher must be 16-byte aligned");
    const int64_t shard = udape_dp_shard_elems(n_total, peers->world);
    const unsigned grid = static_cast<unsigned>(shard / kDpChunk * peers->world);
    dp_gather_ema_kernel<<<grid, kDpThreads, 0, as_stream(stream)>>>(*peers, n_total, shard, teacher, ema_a, ema_b, found_inf, step_dev, epoch_dev, ticket);
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
