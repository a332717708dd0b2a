// decode.cu — heatmap arg-max decoding, confidence masks, rectified pseudo-labels,
// PCK hit counting and the k-th-value consistency mask.
//
// Replaces (reference file:line)
//   lib/keypoint_detection.py:9-37   get_max_preds        (numpy, forces a D2H copy)
//   utils.py:54-75                   get_max_preds_torch
//   train_human.py:376-383,427-430   conf / pred_position / conf_table / activates /
//                                    kthvalue threshold / tea_mask   (inline fragments)
//   utils.py:77-109                  rectify              (B*K Python loop, 4 syncs each)
//   lib/keypoint_detection.py:40-94  calc_dists / dist_acc / accuracy (B*K Python loop)
//
// One CTA owns one (b,k) heatmap plane.  The plane is read once with 128-bit streaming
// loads (all of a thread's loads issued before the first use); every element is mapped
// to an order-preserving u32 key and the (key, ~index) pair is max-reduced as a u64 —
// exact, order-independent, first-index tie-break, NaN-is-max — so indices, masks and
// PCK counts are bit-identical to numpy/torch by construction.
#include <cmath>

#include "common.cuh"
#include "gauss.cuh"
#include "pipeline.cuh"
#include "select.cuh"

namespace udape {

constexpr int kDecThreads = 256;
constexpr int kDecUnroll = 4;
constexpr int64_t kTmaMinPlanes = 2048;

template <typename T> __device__ __forceinline__ float vec_max_nan(const uint4& v);
template <> __device__ __forceinline__ float vec_max_nan<float>(const uint4& v) {
    return fmax_nan(fmax_nan(__uint_as_float(v.x), __uint_as_float(v.y)),
                    fmax_nan(__uint_as_float(v.z), __uint_as_float(v.w)));
}
template <> __device__ __forceinline__ float vec_max_nan<__half>(const uint4& v) {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    const __half2 m = __hmax2_nan(__hmax2_nan(h[0], h[1]), __hmax2_nan(h[2], h[3]));
    return fmax_nan(__low2float(m), __high2float(m));
}
template <> __device__ __forceinline__ float vec_max_nan<__nv_bfloat16>(const uint4& v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    const __nv_bfloat162 m = __hmax2_nan(__hmax2_nan(h[0], h[1]), __hmax2_nan(h[2], h[3]));
    return fmax_nan(__low2float(m), __high2float(m));
}

// One thread's share of a plane -> packed (ordered key of its maximum, ~index of the first
// occurrence).  The scan itself works at vector granularity with plain float maxima (about 1.5
// instructions per element): `best` is the NaN-ignoring maximum with strict > (keeps the first
// vector), `all` the NaN-propagating one.  Only the winning vector is looked at element by
// element, and only then is the exact ordered key formed — so the block reduction stays the
// exact, order-independent u64 maximum (first index wins, NaN is the maximum, -0.0 == +0.0).
template <typename T, bool VEC>
__device__ __forceinline__ unsigned long long thread_plane_argmax(const T* __restrict__ p, int hw) {
    uint32_t best_key = 0u, best_idx = 0xffffffffu;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const int nvec = hw / EPV;
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
        float best = -INFINITY, all = -INFINITY;
        int bv = -1;
        for (int base = 0; base < nvec; base += kDecThreads * kDecUnroll) {
            uint4 v[kDecUnroll];
#pragma unroll
            for (int u = 0; u < kDecUnroll; ++u) {
                const int i = base + u * kDecThreads + threadIdx.x;
                v[u] = (i < nvec) ? ldg_stream(p4 + i) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kDecUnroll; ++u) {
                const int i = base + u * kDecThreads + threadIdx.x;
                if (i < nvec) {
                    const float vm = vec_max_nan<T>(v[u]);
                    all = fmax_nan(all, vm);
                    if (vm > best) { best = vm; bv = i; }
                }
            }
        }
        if (all != all) {
            // rare: this thread's first NaN
            for (int i = threadIdx.x; i < nvec && best_idx == 0xffffffffu; i += kDecThreads) {
                float f[EPV];
                unpack16<T>(ldg_cached(p4 + i), f);
#pragma unroll
                for (int e = EPV - 1; e >= 0; --e)
                    if (f[e] != f[e]) best_idx = static_cast<uint32_t>(i * EPV + e);
            }
            best_key = 0xffffffffu;
        } else if (bv >= 0) {
            float f[EPV];
            unpack16<T>(ldg_cached(p4 + bv), f);
#pragma unroll
            for (int e = EPV - 1; e >= 0; --e)
                if (f[e] == best) best_idx = static_cast<uint32_t>(bv * EPV + e);
            best_key = order_key(best);
        } else if (static_cast<int>(threadIdx.x) < nvec) {
            best_key = order_key(-INFINITY);  // everything this thread saw is -inf: its first element
            best_idx = threadIdx.x * EPV;
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kDecThreads) {
            const uint32_t k = order_key(to_f32<T>(p[i]));
            if (k > best_key) { best_key = k; best_idx = i; }
        }
    }
    return pack_arg(best_key, best_idx);
}

// One 16-byte vector (EPV elements starting at flat index `flat`) of the rectified plane.
// Almost every vector of a plane lies outside the (6*sigma+1)^2 window: when the vector does not
// straddle rows and misses the window's clipped range it is zeros without any per-element work.
template <typename T>
__device__ __forceinline__ uint4 rectified_vector_at(int x, int y, int w, const RectGeom& geom, const GaussWindow& gw,
                                                     const float* __restrict__ tab) {
    constexpr int EPV = Vec16<T>::EPV;
    if (x + EPV <= w && (y < geom.y0i || y >= geom.y1i || x + EPV <= geom.x0i || x >= geom.x1i))
        return make_uint4(0u, 0u, 0u, 0u);
    float f[EPV];
#pragma unroll
    for (int e = 0; e < EPV; ++e) {
        f[e] = rectified_value(x, y, geom, gw, tab);
        if (++x == w) { x = 0; ++y; }
    }
    return pack16<T>(f);
}
constexpr int kSelThreads = 1024;
constexpr int kSelCache = 8192;     // floats of static shared memory of the stand-alone kernel (B*K of every config)

__global__ void __launch_bounds__(kSelThreads)
mask_select_kernel(const float* __restrict__ act, int n, SelectArgs sa) {
    __shared__ float s_cache[kSelCache];
    select_body(act, n, sa, s_cache, kSelCache);
}

// (8 / 6 CTAs per SM: the select / flag tails are separate functions whose registers must not cost the scan its occupancy)
template <typename T, bool VEC>
__global__ void __launch_bounds__(kDecThreads, 8)
decode_kernel(const T* __restrict__ hm, int hw, int w, int h, int32_t* __restrict__ idx_out,
              float* __restrict__ preds, T* __restrict__ maxvals, float* __restrict__ maxvals_f32,
              int64_t* __restrict__ position, float occlude_thresh, uint8_t* __restrict__ conf_table,
              GaussWindow gw, T* __restrict__ rect, SelectArgs sel) {
    __shared__ unsigned long long red[32];
    const int64_t plane = blockIdx.x;
    const T* p = hm + plane * hw;
    // one plane per CTA: tabulating the window would cost as many expf as evaluating it in place
    // (measured: 40.7 -> 44.1 us at C5 with a per-CTA table); the persistent kernel below tabulates
    const float* tab = nullptr;
    const unsigned long long best =
        block_max_u64<kDecThreads>(thread_plane_argmax<T, VEC>(p, hw), red);
    const uint32_t idx = arg_idx(best);
    const float mv = key_value(arg_key(best));
    const int ix = static_cast<int>(idx % static_cast<uint32_t>(w));
    const int iy = static_cast<int>(idx / static_cast<uint32_t>(w));
    const bool positive = mv > 0.0f;  // false for NaN, like np.greater / torch.gt
    if (threadIdx.x == 0) {
        if (idx_out) idx_out[plane] = static_cast<int32_t>(idx);
        if (preds) {
            preds[2 * plane] = positive ? static_cast<float>(ix) : 0.0f;
            preds[2 * plane + 1] = positive ? static_cast<float>(iy) : 0.0f;
        }
        if (maxvals) maxvals[plane] = from_f32<T>(mv);
        if (maxvals_f32) maxvals_f32[plane] = mv;
        if (position) { position[2 * plane] = ix; position[2 * plane + 1] = iy; }
        if (conf_table) conf_table[plane] = (mv >= occlude_thresh) ? 1 : 0;
    }
    if (rect != nullptr) {
    // ---- rectify: utils.py:84-107 ----
    const float mu_x = positive ? static_cast<float>(ix) : 0.0f;
    const float mu_y = positive ? static_cast<float>(iy) : 0.0f;
    const RectGeom geom = rect_geometry(mu_x, mu_y, h, w, gw);
    T* r = rect + plane * hw;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const int nvec = hw / EPV;
        uint4* r4 = reinterpret_cast<uint4*>(r);
        int y = (threadIdx.x * EPV) / w, x = threadIdx.x * EPV - y * w;  // one division, then incremental
        const int step_y = (kDecThreads * EPV) / w, step_x = kDecThreads * EPV - step_y * w;
        for (int i = threadIdx.x; i < nvec; i += kDecThreads) {
            stg_stream(r4 + i, rectified_vector_at<T>(x, y, w, geom, gw, tab));
            x += step_x; y += step_y;
            if (x >= w) { x -= w; ++y; }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kDecThreads) {
            const int y = i / w, x = i - y * w;
            r[i] = from_f32<T>(rectified_value(x, y, geom, gw, tab));
        }
    }
    }
    // train_human.py:427-430 in the same launch: the CTA that finishes last selects the k-th activation
    if (sel.ticket != nullptr && last_block_done(sel.ticket, gridDim.x)) {
        extern __shared__ __align__(16) float sel_cache[];
        select_body(maxvals_f32, static_cast<int>(gridDim.x), sel, sel.cache_elems > 0 ? sel_cache : nullptr, sel.cache_elems);
    }
}

// ---- TMA-staged warp-per-plane arg-max (pipeline.cuh) ---------------------------------------
// The plane sits in shared memory; a warp scans it once with 128-bit LDS keeping, per lane, the
// NaN-ignoring running maximum at vector granularity (strict > keeps the first vector) and a
// NaN-propagating maximum of everything (FMNMX.NAN).  One CREDUX.MAX.F32.NAN gives the plane
// maximum m (NaN iff the plane holds a NaN); lanes whose maximum equals m look up the first
// matching element of their winning vector and one integer redux.min yields the first index.
// Floating-point == makes -0.0 and +0.0 tie, so the first of them wins, like numpy/torch.
struct PlaneMax {
    float val;     // plane maximum (NaN if the plane holds a NaN)
    uint32_t idx;  // flat index of its first occurrence
};

template <typename T>
__device__ __forceinline__ PlaneMax warp_argmax_smem(const uint8_t* __restrict__ src, int nvec, int lane) {
    constexpr int EPV = Vec16<T>::EPV;
    float best = -INFINITY, all = -INFINITY;
    int bv = lane;
#pragma unroll 4
    for (int v = lane; v < nvec; v += 32) {
        const float vm = vec_max_nan<T>(lds128(src + 16 * v));
        all = fmax_nan(all, vm);
        if (vm > best) { best = vm; bv = v; }
    }
    PlaneMax r;
    r.val = warp_max_nan(all);
    uint32_t idx = 0xffffffffu;
    if (r.val != r.val) {
        // rare: first NaN of the plane
        for (int v = lane; v < nvec && idx == 0xffffffffu; v += 32) {
            float f[EPV];
            unpack16<T>(lds128(src + 16 * v), f);
#pragma unroll
            for (int e = EPV - 1; e >= 0; --e)
                if (f[e] != f[e]) idx = static_cast<uint32_t>(v * EPV + e);
        }
    } else if (best == r.val && bv < nvec) {
        float f[EPV];
        unpack16<T>(lds128(src + 16 * bv), f);
#pragma unroll
        for (int e = EPV - 1; e >= 0; --e)
            if (f[e] == r.val) idx = static_cast<uint32_t>(bv * EPV + e);
    }
    r.idx = __reduce_min_sync(0xffffffffu, idx);
    return r;
}

template <typename T>
__global__ void __launch_bounds__(kPipeThreads, 1)
decode_tma_kernel(const T* __restrict__ hm, int64_t planes, int hw, int w, int h, int32_t* __restrict__ idx_out,
                  float* __restrict__ preds, T* __restrict__ maxvals, float* __restrict__ maxvals_f32,
                  int64_t* __restrict__ position, float occlude_thresh, uint8_t* __restrict__ conf_table,
                  GaussWindow gw, T* __restrict__ rect, int group, int stages, SelectArgs sel) {
    constexpr int EPV = Vec16<T>::EPV;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ PipeBarriers bars;
    __shared__ float s_tab[kWinTabN * kWinTabN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nvec = hw / EPV;
    const uint32_t plane_bytes = 16u * nvec, stage_bytes = plane_bytes * group;
    const float* tab = rect ? build_window_table(s_tab, gw) : nullptr;  // published by pipe_init's barrier
    pipe_init(bars, stages);
    const int n_items = pipe_items((planes + group - 1) / group);
    if (warp == static_cast<int>(blockDim.x >> 5) - 1) {
        pipe_produce(bars, stages, n_items, [&](int i, int s, uint64_t* full) {
            const int64_t first = (blockIdx.x + static_cast<int64_t>(i) * gridDim.x) * group;
            const uint32_t np = static_cast<uint32_t>(min(static_cast<int64_t>(group), planes - first));
            mbar_expect_tx(full, np * plane_bytes);
            bulk_load(smem + static_cast<size_t>(s) * stage_bytes, hm + first * hw, np * plane_bytes, full);
        });
    } else {
    constexpr int kMaxGroup = 8;  // pipe_geometry's max_group
    PlaneMax pm[kMaxGroup];
    pipe_consume(
        bars, stages, n_items, warp,
        [&](int i, int s) {
            const int64_t first = (blockIdx.x + static_cast<int64_t>(i) * gridDim.x) * group;
            const int np = static_cast<int>(min(static_cast<int64_t>(group), planes - first));
#pragma unroll
            for (int g = 0; g < kMaxGroup; ++g)
                if (g < np)
                    pm[g] = warp_argmax_smem<T>(smem + static_cast<size_t>(s) * stage_bytes +
                                                    static_cast<size_t>(g) * plane_bytes, nvec, lane);
        },
        [&](int i) {
            // everything below needs only the arg-max: the stage is already being refilled
            const int64_t first = (blockIdx.x + static_cast<int64_t>(i) * gridDim.x) * group;
            const int np = static_cast<int>(min(static_cast<int64_t>(group), planes - first));
#pragma unroll
            for (int g = 0; g < kMaxGroup; ++g) {
                if (g >= np) break;
                const int64_t plane = first + g;
                const float mv = pm[g].val;
                const uint32_t idx = pm[g].idx;
                const int ix = static_cast<int>(idx % static_cast<uint32_t>(w));
                const int iy = static_cast<int>(idx / static_cast<uint32_t>(w));
                const bool positive = mv > 0.0f;  // false for NaN, like np.greater / torch.gt
                if (lane == 0) {
                    if (idx_out) idx_out[plane] = static_cast<int32_t>(idx);
                    if (preds) {
                        preds[2 * plane] = positive ? static_cast<float>(ix) : 0.0f;
                        preds[2 * plane + 1] = positive ? static_cast<float>(iy) : 0.0f;
                    }
                    if (maxvals) maxvals[plane] = from_f32<T>(mv);
                    if (maxvals_f32) maxvals_f32[plane] = mv;
                    if (position) { position[2 * plane] = ix; position[2 * plane + 1] = iy; }
                    if (conf_table) conf_table[plane] = (mv >= occlude_thresh) ? 1 : 0;
                }
                if (rect != nullptr) {
                    // ---- rectify: utils.py:84-107 ----
                    const RectGeom geom = rect_geometry(positive ? static_cast<float>(ix) : 0.0f,
                                                        positive ? static_cast<float>(iy) : 0.0f, h, w, gw);
                    uint4* r4 = reinterpret_cast<uint4*>(rect + plane * hw);
                    int y = (lane * EPV) / w, x = lane * EPV - y * w;  // one division per plane, then incremental
                    for (int v = lane; v < nvec; v += 32) {
                        stg_stream(r4 + v, rectified_vector_at<T>(x, y, w, geom, gw, tab));
                        x += 32 * EPV;
                        while (x >= w) { x -= w; ++y; }
                    }
                }
            }
        });
    }
    // (the ring is idle by now — every consumer warp has passed the ticket's barrier: it caches the select's values)
    if (sel.ticket != nullptr && last_block_done(sel.ticket, gridDim.x))
        select_body(maxvals_f32, static_cast<int>(planes), sel, reinterpret_cast<float*>(smem),
                    static_cast<int>(static_cast<size_t>(stages) * stage_bytes / sizeof(float)));
}

// ---- PCK -------------------------------------------------------------------------------------
// calc_dists + dist_acc for one (b,k) pair (keypoint_detection.py:40-62): returns bit0 = valid,
// bit1 = hit, and writes the decoded coordinates.
__device__ __forceinline__ uint8_t pck_flags(uint32_t io, float vo, uint32_t it, float vt, int w, double norm_x,
                                             double norm_y, double thr, int64_t plane, float* __restrict__ pred_out,
                                             float* __restrict__ tgt_out) {
    const bool po = vo > 0.0f, pt = vt > 0.0f;
    // get_max_preds: float32 coordinates, zeroed when max <= 0 (keypoint_detection.py:28-36)
    const float px = po ? static_cast<float>(io % static_cast<uint32_t>(w)) : 0.0f;
    const float py = po ? static_cast<float>(io / static_cast<uint32_t>(w)) : 0.0f;
    const float tx = pt ? static_cast<float>(it % static_cast<uint32_t>(w)) : 0.0f;
    const float ty = pt ? static_cast<float>(it / static_cast<uint32_t>(w)) : 0.0f;
    if (pred_out) { pred_out[2 * plane] = px; pred_out[2 * plane + 1] = py; }
    if (tgt_out) { tgt_out[2 * plane] = tx; tgt_out[2 * plane + 1] = ty; }
    uint8_t f = 0;
    // calc_dists (keypoint_detection.py:40-52): float64, no contraction
    if (tx > 1.0f && ty > 1.0f) {
        const double d0 = __dsub_rn(__ddiv_rn(static_cast<double>(px), norm_x),
                                    __ddiv_rn(static_cast<double>(tx), norm_x));
        const double d1 = __dsub_rn(__ddiv_rn(static_cast<double>(py), norm_y),
                                    __ddiv_rn(static_cast<double>(ty), norm_y));
        const double dist = sqrt(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)));
        f = 1;
        if (dist < thr) f |= 2;  // dist_acc: strict < (keypoint_detection.py:60)
    }
    return f;
}

// last CTA: per-joint integer sums of the per-plane flags (exact, order-independent)
__device__ __forceinline__ void pck_reduce_flags(const uint8_t* flags, int64_t planes, int joints,
                                                 int32_t* __restrict__ hits, int32_t* __restrict__ valid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int64_t batch = planes / joints;
    constexpr int UN = 4;   // flags of other CTAs, from L2, four loads of a lane in flight (not a chain of round trips)
    for (int k = warp; k < joints; k += warps) {
        int nh = 0, nv = 0;
        for (int64_t b = lane; b < batch; b += 32 * UN) {
            uint32_t f[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) f[u] = (b + 32 * u < batch) ? __ldcg(flags + (b + 32 * u) * joints + k) : 0u;
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                nv += f[u] & 1;
                nh += (f[u] >> 1) & 1;
            }
        }
        nh = __reduce_add_sync(0xffffffffu, nh);
        nv = __reduce_add_sync(0xffffffffu, nv);
        if (lane == 0) { hits[k] = nh; valid[k] = nv; }
    }
}

template <typename TO, typename TT, bool VEC_O, bool VEC_T>
__global__ void __launch_bounds__(kDecThreads, 6)
pck_kernel(const TO* __restrict__ output, const TT* __restrict__ target, int joints, int hw, int w,
           double norm_x, double norm_y, double thr, float* __restrict__ pred_out,
           float* __restrict__ tgt_out, int32_t* __restrict__ hits, int32_t* __restrict__ valid,
           uint8_t* __restrict__ flags, uint32_t* __restrict__ ticket) {
    __shared__ unsigned long long red[32];
    const int64_t plane = blockIdx.x;
    const unsigned long long bo =
        block_max_u64<kDecThreads>(thread_plane_argmax<TO, VEC_O>(output + plane * hw, hw), red);
    const unsigned long long bt =
        block_max_u64<kDecThreads>(thread_plane_argmax<TT, VEC_T>(target + plane * hw, hw), red);
    if (threadIdx.x == 0)
        flags[plane] = pck_flags(arg_idx(bo), key_value(arg_key(bo)), arg_idx(bt), key_value(arg_key(bt)), w, norm_x,
                                 norm_y, thr, plane, pred_out, tgt_out);
    if (last_block_done(ticket, gridDim.x)) pck_reduce_flags(flags, gridDim.x, joints, hits, valid);
}

// TMA-staged: item = [output plane | target plane]; one warp decodes both planes of a pair
template <typename TO, typename TT>
__global__ void __launch_bounds__(kPipeThreads, 1)
pck_tma_kernel(const TO* __restrict__ output, const TT* __restrict__ target, int64_t planes, int joints, int hw, int w,
               double norm_x, double norm_y, double thr, float* __restrict__ pred_out, float* __restrict__ tgt_out,
               int32_t* __restrict__ hits, int32_t* __restrict__ valid, uint8_t* __restrict__ flags,
               uint32_t* __restrict__ ticket, int stages) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ PipeBarriers bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bytes_o = static_cast<uint32_t>(hw * sizeof(TO)), bytes_t = static_cast<uint32_t>(hw * sizeof(TT));
    const uint32_t stage_bytes = bytes_o + bytes_t;
    pipe_init(bars, stages);
    const int n_items = pipe_items(planes);
    if (warp == static_cast<int>(blockDim.x >> 5) - 1) {
        pipe_produce(bars, stages, n_items, [&](int i, int s, uint64_t* full) {
            const int64_t plane = blockIdx.x + static_cast<int64_t>(i) * gridDim.x;
            uint8_t* dst = smem + static_cast<size_t>(s) * stage_bytes;
            mbar_expect_tx(full, stage_bytes);
            bulk_load(dst, output + plane * hw, bytes_o, full);
            bulk_load(dst + bytes_o, target + plane * hw, bytes_t, full);
        });
    } else {
        PlaneMax mo, mt;
        pipe_consume(
            bars, stages, n_items, warp,
            [&](int i, int s) {
                const uint8_t* src = smem + static_cast<size_t>(s) * stage_bytes;
                mo = warp_argmax_smem<TO>(src, hw / Vec16<TO>::EPV, lane);
                mt = warp_argmax_smem<TT>(src + bytes_o, hw / Vec16<TT>::EPV, lane);
            },
            [&](int i) {
                const int64_t plane = blockIdx.x + static_cast<int64_t>(i) * gridDim.x;
                if (lane == 0)
                    flags[plane] = pck_flags(mo.idx, mo.val, mt.idx, mt.val, w, norm_x, norm_y, thr, plane, pred_out, tgt_out);
            });
    }
    __syncthreads();
    if (last_block_done(ticket, gridDim.x)) pck_reduce_flags(flags, planes, joints, hits, valid);
}

template <typename T>
static int launch_decode(const void* hm, int64_t planes, int64_t h, int64_t w, int32_t* idx,
                         float* preds, void* maxvals, float* maxvals_f32, int64_t* position,
                         float occlude_thresh, uint8_t* conf_table, double sigma, void* rect,
                         const SelectArgs& sel, cudaStream_t st) {
    const int hw = static_cast<int>(h * w);
    const bool vec = aligned16(hm) && (hw % Vec16<T>::EPV) == 0 && (rect == nullptr || aligned16(rect));
    const GaussWindow gw = make_window(rect ? sigma : 1.0);
    const unsigned grid = static_cast<unsigned>(planes);
    PipeGeom pg = {1, 0, 0};
    // TMA staging pays for pure scans over many planes (measured cross-over ~2k planes on B200); with a
    // rectified map to write, or few planes, the many-warp register path below is faster
    if (vec && pipe_enabled() && rect == nullptr && planes >= kTmaMinPlanes && hw * static_cast<int>(sizeof(T)) >= 2048)
        pg = pipe_geometry(hw * static_cast<int64_t>(sizeof(T)), planes);
    if (pg.stages >= 2) {
        const size_t smem = static_cast<size_t>(pg.stages) * pg.group * hw * sizeof(T);
        cudaError_t e = pipe_reserve_smem<&decode_tma_kernel<T>>(smem);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_decode: smem opt-in: %s", cudaGetErrorString(e));
        decode_tma_kernel<T><<<pipe_grid((planes + pg.group - 1) / pg.group), pg.threads, smem, st>>>(
            static_cast<const T*>(hm), planes, hw, static_cast<int>(w), static_cast<int>(h), idx, preds,
            static_cast<T*>(maxvals), maxvals_f32, position, occlude_thresh, conf_table, gw,
            static_cast<T*>(rect), pg.group, pg.stages, sel);
        return check_launch("udape_decode");
    }
    // the fused select caches its B*K values in dynamic shared memory (every CTA reserves it, the last one uses it)
    SelectArgs sel2 = sel;
    const size_t sel_smem = (sel.ticket != nullptr && planes * sizeof(float) <= 32 * 1024) ? ((planes * sizeof(float) + 15) & ~size_t(15)) : 0;
    sel2.cache_elems = static_cast<int>(sel_smem / sizeof(float));
    if (vec)
        decode_kernel<T, true><<<grid, kDecThreads, sel_smem, st>>>(
            static_cast<const T*>(hm), hw, static_cast<int>(w), static_cast<int>(h), idx, preds,
            static_cast<T*>(maxvals), maxvals_f32, position, occlude_thresh, conf_table, gw,
            static_cast<T*>(rect), sel2);
    else
        decode_kernel<T, false><<<grid, kDecThreads, sel_smem, st>>>(
            static_cast<const T*>(hm), hw, static_cast<int>(w), static_cast<int>(h), idx, preds,
            static_cast<T*>(maxvals), maxvals_f32, position, occlude_thresh, conf_table, gw,
            static_cast<T*>(rect), sel2);
    return check_launch("udape_decode");
}

template <typename TO, typename TT>
static int launch_pck(const void* output, const void* target, int64_t planes, int64_t joints,
                      int64_t h, int64_t w, double thr, float* pred, float* tgt, int32_t* hits,
                      int32_t* valid, uint8_t* flags, uint32_t* ticket, cudaStream_t st) {
    const int hw = static_cast<int>(h * w);
    const bool vo = aligned16(output) && (hw % Vec16<TO>::EPV) == 0;
    const bool vt = aligned16(target) && (hw % Vec16<TT>::EPV) == 0;
    // accuracy(): norm = ones((B,2)) * [h, w] / 10, applied to (x, y)  (keypoint_detection.py:79)
    const double norm_x = static_cast<double>(h) / 10.0, norm_y = static_cast<double>(w) / 10.0;
    const unsigned grid = static_cast<unsigned>(planes);
    const TO* o = static_cast<const TO*>(output);
    const TT* t = static_cast<const TT*>(target);
    const int ij = static_cast<int>(joints), iw = static_cast<int>(w);
    PipeGeom pg = {1, 0, 0};
    if (vo && vt && pipe_enabled() && planes >= kTmaMinPlanes && hw * static_cast<int64_t>(sizeof(TO)) >= 2048)
        pg = pipe_geometry(hw * static_cast<int64_t>(sizeof(TO) + sizeof(TT)), planes, 1);
    if (pg.stages >= 2) {
        const size_t smem = static_cast<size_t>(pg.stages) * hw * (sizeof(TO) + sizeof(TT));
        cudaError_t e = pipe_reserve_smem<&pck_tma_kernel<TO, TT>>(smem);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_pck_counts: smem opt-in: %s", cudaGetErrorString(e));
        pck_tma_kernel<TO, TT><<<pipe_grid(planes), pg.threads, smem, st>>>(o, t, planes, ij, hw, iw, norm_x, norm_y, thr,
                                                                           pred, tgt, hits, valid, flags, ticket, pg.stages);
        return check_launch("udape_pck_counts");
    }
#define UDAPE_PCK_LAUNCH(VO, VT) \
    pck_kernel<TO, TT, VO, VT><<<grid, kDecThreads, 0, st>>>(o, t, ij, hw, iw, norm_x, norm_y, thr, pred, tgt, hits, valid, flags, ticket)
    if (vo && vt) UDAPE_PCK_LAUNCH(true, true);
    else if (vo) UDAPE_PCK_LAUNCH(true, false);
    else if (vt) UDAPE_PCK_LAUNCH(false, true);
    else UDAPE_PCK_LAUNCH(false, false);
#undef UDAPE_PCK_LAUNCH
    return check_launch("udape_pck_counts");
}

}  // namespace udape

using namespace udape;

static int decode_entry(const void* hm, int dtype, int64_t planes, int64_t h, int64_t w,
                        int32_t* idx, float* preds, void* maxvals, float* maxvals_f32,
                        int64_t* position, float occlude_thresh, uint8_t* conf_table,
                        double sigma, void* rectified, const SelectArgs& sel, void* stream) {
    UDAPE_REQUIRE(hm, UDAPE_ERR_NULL, "udape_decode: hm is NULL");
    UDAPE_REQUIRE(planes > 0 && h > 0 && w > 0 && planes < (1ll << 31) && h * w < (1ll << 31),
                  UDAPE_ERR_SHAPE, "udape_decode: bad extents planes=%lld h=%lld w=%lld",
                  (long long)planes, (long long)h, (long long)w);
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(es == 2 || es == 4, UDAPE_ERR_DTYPE, "udape_decode: unsupported dtype code %d", dtype);
    UDAPE_REQUIRE(aligned_to(hm, es) && (!maxvals || aligned_to(maxvals, es)) &&
                      (!rectified || aligned_to(rectified, es)) && (!position || aligned_to(position, 8)) &&
                      (!preds || aligned_to(preds, 4)) && (!idx || aligned_to(idx, 4)) &&
                      (!maxvals_f32 || aligned_to(maxvals_f32, 4)),
                  UDAPE_ERR_ALIGN, "udape_decode: misaligned pointer");
    if (rectified) {
        UDAPE_REQUIRE(sigma > 0.0 && sigma < 1e4, UDAPE_ERR_ARG, "udape_decode: sigma %g out of range", sigma);
    }
    UDAPE_DISPATCH_FLOAT(dtype, T, return launch_decode<T>(hm, planes, h, w, idx, preds, maxvals, maxvals_f32, position, occlude_thresh, conf_table, sigma, rectified, sel, as_stream(stream)));
    return UDAPE_OK;
}

extern "C" int udape_decode(const void* hm, int dtype, int64_t planes, int64_t h, int64_t w,
                            int32_t* idx, float* preds, void* maxvals, float* maxvals_f32,
                            int64_t* position, float occlude_thresh, uint8_t* conf_table,
                            double sigma, void* rectified, void* stream) {
    const SelectArgs none = {0, nullptr, nullptr, nullptr, nullptr};
    return decode_entry(hm, dtype, planes, h, w, idx, preds, maxvals, maxvals_f32, position, occlude_thresh, conf_table,
                        sigma, rectified, none, stream);
}

extern "C" int udape_decode_select(const void* hm, int dtype, int64_t planes, int64_t h, int64_t w,
                                   int32_t* idx, float* preds, void* maxvals, float* maxvals_f32,
                                   int64_t* position, float occlude_thresh, uint8_t* conf_table,
                                   double sigma, void* rectified, int64_t kth, const float* tea_mask_in,
                                   float* thresh_out, uint8_t* tea_mask_out, uint32_t* ticket, void* stream) {
    UDAPE_REQUIRE(maxvals_f32 && ticket, UDAPE_ERR_NULL, "udape_decode_select: maxvals_f32 and ticket are required");
    UDAPE_REQUIRE(aligned_to(ticket, 4) && (!thresh_out || aligned_to(thresh_out, 4)) && (!tea_mask_in || aligned_to(tea_mask_in, 4)),
                  UDAPE_ERR_ALIGN, "udape_decode_select: misaligned pointer");
    // torch.kthvalue raises for k outside [1, n]
    UDAPE_REQUIRE(kth >= 1 && kth <= planes, UDAPE_ERR_ARG, "udape_decode_select: kth=%lld outside [1,%lld]",
                  (long long)kth, (long long)planes);
    const SelectArgs sel = {static_cast<int>(kth), tea_mask_in, thresh_out, tea_mask_out, ticket, 0};
    return decode_entry(hm, dtype, planes, h, w, idx, preds, maxvals, maxvals_f32, position, occlude_thresh, conf_table,
                        sigma, rectified, sel, stream);
}

extern "C" int udape_mask_select(const float* activates, int64_t n, int64_t kth,
                                 const float* tea_mask_in, float* thresh_out,
                                 uint8_t* tea_mask_out, void* stream) {
    UDAPE_REQUIRE(activates, UDAPE_ERR_NULL, "udape_mask_select: activates is NULL");
    UDAPE_REQUIRE(n > 0 && n < (1ll << 31), UDAPE_ERR_SHAPE, "udape_mask_select: bad n=%lld", (long long)n);
    // torch.kthvalue raises for k outside [1, n]
    UDAPE_REQUIRE(kth >= 1 && kth <= n, UDAPE_ERR_ARG, "udape_mask_select: kth=%lld outside [1,%lld]",
                  (long long)kth, (long long)n);
    const SelectArgs sa = {static_cast<int>(kth), tea_mask_in, thresh_out, tea_mask_out, nullptr, 0};
    mask_select_kernel<<<1, kSelThreads, 0, as_stream(stream)>>>(activates, static_cast<int>(n), sa);
    return check_launch("udape_mask_select");
}

extern "C" int udape_pck_counts(const void* output, int out_dtype, const void* target, int tgt_dtype,
                                int64_t batch, int64_t joints, int64_t h, int64_t w, double thr,
                                float* pred, float* tgt, int32_t* hits, int32_t* valid, uint8_t* flags,
                                uint32_t* ticket, void* stream) {
    UDAPE_REQUIRE(output && target && hits && valid && flags && ticket, UDAPE_ERR_NULL, "udape_pck_counts: NULL pointer");
    UDAPE_REQUIRE(batch > 0 && joints > 0 && h > 0 && w > 0 && batch * joints < (1ll << 31) &&
                      h * w < (1ll << 31),
                  UDAPE_ERR_SHAPE, "udape_pck_counts: bad extents B=%lld K=%lld h=%lld w=%lld",
                  (long long)batch, (long long)joints, (long long)h, (long long)w);
    const int eo = dtype_size(out_dtype), et = dtype_size(tgt_dtype);
    UDAPE_REQUIRE((eo == 2 || eo == 4) && (et == 2 || et == 4), UDAPE_ERR_DTYPE,
                  "udape_pck_counts: unsupported dtype codes %d/%d", out_dtype, tgt_dtype);
    UDAPE_REQUIRE(aligned_to(output, eo) && aligned_to(target, et) && aligned_to(ticket, 4) && aligned_to(hits, 4) &&
                      aligned_to(valid, 4),
                  UDAPE_ERR_ALIGN, "udape_pck_counts: misaligned pointer");
    cudaStream_t st = as_stream(stream);
    const int64_t planes = batch * joints;
    UDAPE_DISPATCH_FLOAT(out_dtype, TO,
        UDAPE_DISPATCH_FLOAT(tgt_dtype, TT,
            return launch_pck<TO, TT>(output, target, planes, joints, h, w, thr, pred, tgt, hits, valid, flags, ticket, st)));
    return UDAPE_OK;
}
