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

namespace udape {

constexpr int kDecThreads = 256;
constexpr int kDecUnroll = 4;

// window geometry of the unit-peak Gaussian, derived on the host from sigma exactly as
// utils.py:81,93-98 derives it in Python
struct GaussWindow {
    float tmp;    // 3*sigma                       (utils.py:81)
    int n;        // len(arange(0, 2*tmp+1, 1))    (utils.py:93-94)
    float x0;     // (2*tmp+1) // 2                (utils.py:96)
    float denom;  // 2*sigma**2                    (utils.py:98)
};

template <typename T, bool VEC>
__device__ __forceinline__ unsigned long long thread_plane_argmax(const T* __restrict__ p, int hw) {
    uint32_t best_key = 0u, best_idx = 0xffffffffu;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const int nvec = hw / EPV;
        const uint4* p4 = reinterpret_cast<const uint4*>(p);
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
                    float f[EPV];
                    unpack16<T>(v[u], f);
#pragma unroll
                    for (int e = 0; e < EPV; ++e) {
                        const uint32_t k = order_key(f[e]);
                        // indices visited by one thread are increasing: strict > keeps the first
                        if (k > best_key) { best_key = k; best_idx = i * EPV + e; }
                    }
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kDecThreads) {
            const uint32_t k = order_key(to_f32<T>(p[i]));
            if (k > best_key) { best_key = k; best_idx = i; }
        }
    }
    return pack_arg(best_key, best_idx);
}

// zeros + clipped Gaussian window, following utils.py:84-107 (including its use of h for
// the x bound and w for the y bound)
template <typename T>
__device__ __forceinline__ float rectified_value(int x, int y, int ul_x, int ul_y, int x0i, int x1i,
                                                 int y0i, int y1i, const GaussWindow& g) {
    if (x < x0i || x >= x1i || y < y0i || y >= y1i) return 0.0f;
    const int gx = x - ul_x, gy = y - ul_y;
    if (gx >= g.n || gy >= g.n) return 0.0f;
    const float dx = static_cast<float>(gx) - g.x0, dy = static_cast<float>(gy) - g.x0;
    const float d2 = dx * dx + dy * dy;
    return expf(-(d2 / g.denom));
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kDecThreads)
decode_kernel(const T* __restrict__ hm, int hw, int w, int h, int32_t* __restrict__ idx_out,
              float* __restrict__ preds, T* __restrict__ maxvals, float* __restrict__ maxvals_f32,
              int64_t* __restrict__ position, float occlude_thresh, uint8_t* __restrict__ conf_table,
              GaussWindow gw, T* __restrict__ rect) {
    __shared__ unsigned long long red[32];
    const int64_t plane = blockIdx.x;
    const T* p = hm + plane * hw;
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
    if (rect == nullptr) return;

    // ---- rectify: utils.py:84-107 ----
    const float mu_x = positive ? static_cast<float>(ix) : 0.0f;
    const float mu_y = positive ? static_cast<float>(iy) : 0.0f;
    const int ul_x = static_cast<int>(mu_x - gw.tmp), ul_y = static_cast<int>(mu_y - gw.tmp);
    const int br_x = static_cast<int>(mu_x + gw.tmp + 1.0f), br_y = static_cast<int>(mu_y + gw.tmp + 1.0f);
    const bool skip = (mu_x >= static_cast<float>(h)) || (mu_y >= static_cast<float>(w));
    int x0i = max(0, ul_x), x1i = min(min(br_x, h), w);
    int y0i = max(0, ul_y), y1i = min(min(br_y, w), h);
    if (skip) { x1i = x0i = 0; y1i = y0i = 0; }
    T* r = rect + plane * hw;
    if (VEC) {
        constexpr int EPV = Vec16<T>::EPV;
        const int nvec = hw / EPV;
        uint4* r4 = reinterpret_cast<uint4*>(r);
        for (int i = threadIdx.x; i < nvec; i += kDecThreads) {
            const int flat = i * EPV;
            int y = flat / w, x = flat - y * w;
            float f[EPV];
#pragma unroll
            for (int e = 0; e < EPV; ++e) {
                f[e] = rectified_value<T>(x, y, ul_x, ul_y, x0i, x1i, y0i, y1i, gw);
                if (++x == w) { x = 0; ++y; }
            }
            stg_stream(r4 + i, pack16<T>(f));
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kDecThreads) {
            const int y = i / w, x = i - y * w;
            r[i] = from_f32<T>(rectified_value<T>(x, y, ul_x, ul_y, x0i, x1i, y0i, y1i, gw));
        }
    }
}

// ---- PCK -------------------------------------------------------------------------------------
template <typename TO, typename TT, bool VEC_O, bool VEC_T>
__global__ void __launch_bounds__(kDecThreads)
pck_kernel(const TO* __restrict__ output, const TT* __restrict__ target, int joints, int hw, int w,
           double norm_x, double norm_y, double thr, float* __restrict__ pred_out,
           float* __restrict__ tgt_out, int32_t* __restrict__ hits, int32_t* __restrict__ valid) {
    __shared__ unsigned long long red[32];
    const int64_t plane = blockIdx.x;
    const unsigned long long bo =
        block_max_u64<kDecThreads>(thread_plane_argmax<TO, VEC_O>(output + plane * hw, hw), red);
    const unsigned long long bt =
        block_max_u64<kDecThreads>(thread_plane_argmax<TT, VEC_T>(target + plane * hw, hw), red);
    if (threadIdx.x != 0) return;
    const uint32_t io = arg_idx(bo), it = arg_idx(bt);
    const bool po = key_value(arg_key(bo)) > 0.0f, pt = key_value(arg_key(bt)) > 0.0f;
    // get_max_preds: float32 coordinates, zeroed when max <= 0 (keypoint_detection.py:28-36)
    const float px = po ? static_cast<float>(io % static_cast<uint32_t>(w)) : 0.0f;
    const float py = po ? static_cast<float>(io / static_cast<uint32_t>(w)) : 0.0f;
    const float tx = pt ? static_cast<float>(it % static_cast<uint32_t>(w)) : 0.0f;
    const float ty = pt ? static_cast<float>(it / static_cast<uint32_t>(w)) : 0.0f;
    if (pred_out) { pred_out[2 * plane] = px; pred_out[2 * plane + 1] = py; }
    if (tgt_out) { tgt_out[2 * plane] = tx; tgt_out[2 * plane + 1] = ty; }
    // calc_dists (keypoint_detection.py:40-52): float64, no contraction
    if (tx > 1.0f && ty > 1.0f) {
        const double d0 = __dsub_rn(__ddiv_rn(static_cast<double>(px), norm_x),
                                    __ddiv_rn(static_cast<double>(tx), norm_x));
        const double d1 = __dsub_rn(__ddiv_rn(static_cast<double>(py), norm_y),
                                    __ddiv_rn(static_cast<double>(ty), norm_y));
        const double dist = sqrt(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)));
        const int k = static_cast<int>(plane % joints);
        atomicAdd(valid + k, 1);
        if (dist < thr) atomicAdd(hits + k, 1);  // dist_acc: strict < (keypoint_detection.py:60)
    }
}

// ---- k-th value + tea_mask ----------------------------------------------------------------------
// Single-CTA 4-pass radix select on the ordered key (exact; NaN sorts last like
// torch.kthvalue), then tea_mask = (tea_mask_in * activates) > thresh.
constexpr int kSelThreads = 1024;

__global__ void __launch_bounds__(kSelThreads)
mask_select_kernel(const float* __restrict__ act, int n, int kth, const float* __restrict__ tm_in,
                   float* __restrict__ thresh_out, uint8_t* __restrict__ tm_out) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_mask, s_k;
    if (threadIdx.x == 0) { s_prefix = 0u; s_mask = 0u; s_k = static_cast<unsigned>(kth); }
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = pass * 8;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix, mask = s_mask;
        for (int i = threadIdx.x; i < n; i += kSelThreads) {
            const uint32_t key = order_key(act[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
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
    if (threadIdx.x == 0 && thresh_out) *thresh_out = thresh;
    if (tm_out) {
        for (int i = threadIdx.x; i < n; i += kSelThreads) {
            const float a = tm_in ? tm_in[i] * act[i] : act[i];
            tm_out[i] = (a > thresh) ? 1 : 0;
        }
    }
}

static GaussWindow make_window(double sigma) {
    GaussWindow g;
    const double tmp = 3.0 * sigma;
    const double size = 2.0 * tmp + 1.0;
    g.tmp = static_cast<float>(tmp);
    g.n = static_cast<int>(std::ceil(size));
    g.x0 = static_cast<float>(std::floor(size / 2.0));
    g.denom = static_cast<float>(2.0 * sigma * sigma);
    return g;
}

template <typename T>
static int launch_decode(const void* hm, int64_t planes, int64_t h, int64_t w, int32_t* idx,
                         float* preds, void* maxvals, float* maxvals_f32, int64_t* position,
                         float occlude_thresh, uint8_t* conf_table, double sigma, void* rect,
                         cudaStream_t st) {
    const int hw = static_cast<int>(h * w);
    const bool vec = aligned16(hm) && (hw % Vec16<T>::EPV) == 0 && (rect == nullptr || aligned16(rect));
    const GaussWindow gw = make_window(rect ? sigma : 1.0);
    const unsigned grid = static_cast<unsigned>(planes);
    if (vec)
        decode_kernel<T, true><<<grid, kDecThreads, 0, st>>>(
            static_cast<const T*>(hm), hw, static_cast<int>(w), static_cast<int>(h), idx, preds,
            static_cast<T*>(maxvals), maxvals_f32, position, occlude_thresh, conf_table, gw,
            static_cast<T*>(rect));
    else
        decode_kernel<T, false><<<grid, kDecThreads, 0, st>>>(
            static_cast<const T*>(hm), hw, static_cast<int>(w), static_cast<int>(h), idx, preds,
            static_cast<T*>(maxvals), maxvals_f32, position, occlude_thresh, conf_table, gw,
            static_cast<T*>(rect));
    return check_launch("udape_decode");
}

template <typename TO, typename TT>
static int launch_pck(const void* output, const void* target, int64_t planes, int64_t joints,
                      int64_t h, int64_t w, double thr, float* pred, float* tgt, int32_t* hits,
                      int32_t* valid, cudaStream_t st) {
    const int hw = static_cast<int>(h * w);
    const bool vo = aligned16(output) && (hw % Vec16<TO>::EPV) == 0;
    const bool vt = aligned16(target) && (hw % Vec16<TT>::EPV) == 0;
    // accuracy(): norm = ones((B,2)) * [h, w] / 10, applied to (x, y)  (keypoint_detection.py:79)
    const double norm_x = static_cast<double>(h) / 10.0, norm_y = static_cast<double>(w) / 10.0;
    const unsigned grid = static_cast<unsigned>(planes);
    const TO* o = static_cast<const TO*>(output);
    const TT* t = static_cast<const TT*>(target);
    const int ij = static_cast<int>(joints), iw = static_cast<int>(w);
#define UDAPE_PCK_LAUNCH(VO, VT) \
    pck_kernel<TO, TT, VO, VT><<<grid, kDecThreads, 0, st>>>(o, t, ij, hw, iw, norm_x, norm_y, thr, pred, tgt, hits, valid)
    if (vo && vt) UDAPE_PCK_LAUNCH(true, true);
    else if (vo) UDAPE_PCK_LAUNCH(true, false);
    else if (vt) UDAPE_PCK_LAUNCH(false, true);
    else UDAPE_PCK_LAUNCH(false, false);
#undef UDAPE_PCK_LAUNCH
    return check_launch("udape_pck_counts");
}

}  // namespace udape

using namespace udape;

extern "C" int udape_decode(const void* hm, int dtype, int64_t planes, int64_t h, int64_t w,
                            int32_t* idx, float* preds, void* maxvals, float* maxvals_f32,
                            int64_t* position, float occlude_thresh, uint8_t* conf_table,
                            double sigma, void* rectified, void* stream) {
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
    UDAPE_DISPATCH_FLOAT(dtype, T, return launch_decode<T>(hm, planes, h, w, idx, preds, maxvals, maxvals_f32, position, occlude_thresh, conf_table, sigma, rectified, as_stream(stream)));
    return UDAPE_OK;
}

extern "C" int udape_mask_select(const float* activates, int64_t n, int64_t kth,
                                 const float* tea_mask_in, float* thresh_out,
                                 uint8_t* tea_mask_out, void* stream) {
    UDAPE_REQUIRE(activates, UDAPE_ERR_NULL, "udape_mask_select: activates is NULL");
    UDAPE_REQUIRE(n > 0 && n < (1ll << 31), UDAPE_ERR_SHAPE, "udape_mask_select: bad n=%lld", (long long)n);
    // torch.kthvalue raises for k outside [1, n]
    UDAPE_REQUIRE(kth >= 1 && kth <= n, UDAPE_ERR_ARG, "udape_mask_select: kth=%lld outside [1,%lld]",
                  (long long)kth, (long long)n);
    mask_select_kernel<<<1, kSelThreads, 0, as_stream(stream)>>>(
        activates, static_cast<int>(n), static_cast<int>(kth), tea_mask_in, thresh_out, tea_mask_out);
    return check_launch("udape_mask_select");
}

extern "C" int udape_pck_counts(const void* output, int out_dtype, const void* target, int tgt_dtype,
                                int64_t batch, int64_t joints, int64_t h, int64_t w, double thr,
                                float* pred, float* tgt, int32_t* hits, int32_t* valid, void* stream) {
    UDAPE_REQUIRE(output && target && hits && valid, UDAPE_ERR_NULL, "udape_pck_counts: NULL pointer");
    UDAPE_REQUIRE(batch > 0 && joints > 0 && h > 0 && w > 0 && batch * joints < (1ll << 31) &&
                      h * w < (1ll << 31),
                  UDAPE_ERR_SHAPE, "udape_pck_counts: bad extents B=%lld K=%lld h=%lld w=%lld",
                  (long long)batch, (long long)joints, (long long)h, (long long)w);
    const int eo = dtype_size(out_dtype), et = dtype_size(tgt_dtype);
    UDAPE_REQUIRE((eo == 2 || eo == 4) && (et == 2 || et == 4), UDAPE_ERR_DTYPE,
                  "udape_pck_counts: unsupported dtype codes %d/%d", out_dtype, tgt_dtype);
    UDAPE_REQUIRE(aligned_to(output, eo) && aligned_to(target, et), UDAPE_ERR_ALIGN,
                  "udape_pck_counts: misaligned pointer");
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(hits, 0, sizeof(int32_t) * joints, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(valid, 0, sizeof(int32_t) * joints, st);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_pck_counts: memset: %s", cudaGetErrorString(e));
    const int64_t planes = batch * joints;
    UDAPE_DISPATCH_FLOAT(out_dtype, TO,
        UDAPE_DISPATCH_FLOAT(tgt_dtype, TT,
            return launch_pck<TO, TT>(output, target, planes, joints, h, w, thr, pred, tgt, hits, valid, st)));
    return UDAPE_OK;
}
