// heatmap.cu — batched Gaussian target-heatmap writers.
//
// Replaces lib/datasets/util.py:12-70 (generate_target, human/hand datasets) and
// :326-363 (draw_labelmap_ori, animal datasets) of the reference, which run per sample
// and per joint in numpy inside DataLoader workers.  Here one launch writes a whole
// [planes, H, W] float32 batch: pure write bandwidth (planes*H*W*4 bytes), 128-bit
// streaming stores, one CTA per plane.  Integer placement follows the reference's
// conventions exactly (float64 truncation toward zero, out-of-bounds -> weight 0,
// `v > 0.5` gate, border-touching windows rejected by the animal variant).
#include <cmath>

#include "common.cuh"

namespace udape {

constexpr int kHmThreads = 256;
// every plane of a launch pastes the same window: tabulated once per CTA in shared memory when it
// is at most kTabN wide (sigma <= 3), so a window element costs one LDS instead of an exp
constexpr int kTabN = 19;

struct TargetWindow {
    double tmp;   // sigma*3                         (util.py:33)
    int n;        // len(arange(0, 2*tmp+1, 1))      (util.py:51-52)
    float x0;     // (2*tmp+1) // 2                  (util.py:54)
    float denom;  // 2*sigma**2 as float32           (util.py:56)
};

// Python int(x): truncation toward zero; huge/NaN values are pushed out of bounds
__device__ __forceinline__ int trunc_to_int(double v) {
    if (!(v > -1.0e9 && v < 1.0e9)) return v < 0.0 ? -1000000000 : 1000000000;
    return static_cast<int>(v);
}

__device__ __forceinline__ void
gauss_target_plane(const int64_t plane, const double* __restrict__ joints, const float* __restrict__ vis, int hm_w,
                   int hm_h, double stride_x, double stride_y, const TargetWindow& tw,
                   float* __restrict__ target, float* __restrict__ weight) {
    __shared__ int s_geom[6];  // ul_x, ul_y, x0i, x1i, y0i, y1i
    auto window = [&](int gx, int gy) -> float {
        const float dx = static_cast<float>(gx) - tw.x0, dy = static_cast<float>(gy) - tw.x0;
        return expf(-((dx * dx + dy * dy) / tw.denom));  // float32 like np.exp on float32
    };
    // (float32 expf in place is as cheap as a per-CTA table here — measured 17.5 vs 18.4 us at C5;
    // the float64 window of labelmap_kernel below is the one worth tabulating)
    if (threadIdx.x == 0) {
        // util.py:38-39  mu = int(joint / feat_stride + 0.5)   (float64)
        const int mu_x = trunc_to_int(joints[2 * plane] / stride_x + 0.5);
        const int mu_y = trunc_to_int(joints[2 * plane + 1] / stride_y + 0.5);
        const int ul_x = trunc_to_int(mu_x - tw.tmp), ul_y = trunc_to_int(mu_y - tw.tmp);
        const int br_x = trunc_to_int(mu_x + tw.tmp + 1), br_y = trunc_to_int(mu_y + tw.tmp + 1);
        float wgt = vis[plane];
        const bool oob = mu_x >= hm_w || mu_y >= hm_h || mu_x < 0 || mu_y < 0;  // util.py:43-47
        if (oob) wgt = 0.0f;
        weight[plane] = wgt;
        const bool paste = !oob && wgt > 0.5f;  // util.py:65-66
        s_geom[0] = ul_x;
        s_geom[1] = ul_y;
        s_geom[2] = paste ? max(0, ul_x) : 0;
        s_geom[3] = paste ? min(br_x, hm_w) : 0;
        s_geom[4] = paste ? max(0, ul_y) : 0;
        s_geom[5] = paste ? min(br_y, hm_h) : 0;
    }
    __syncthreads();
    const int ul_x = s_geom[0], ul_y = s_geom[1];
    const int x0i = s_geom[2], x1i = s_geom[3], y0i = s_geom[4], y1i = s_geom[5];
    const int hw = hm_w * hm_h;
    float* t = target + plane * static_cast<int64_t>(hw);
    auto value = [&](int x, int y) -> float {
        if (x < x0i || x >= x1i || y < y0i || y >= y1i) return 0.0f;
        const int gx = x - ul_x, gy = y - ul_y;
        if (gx >= tw.n || gy >= tw.n) return 0.0f;
        return window(gx, gy);
    };
    if ((hw & 3) == 0 && aligned16(target)) {
        uint4* t4 = reinterpret_cast<uint4*>(t);
        // position of this thread's first vector (one division), then incremental: no per-vector division
        int y = (threadIdx.x * 4) / hm_w, x = threadIdx.x * 4 - y * hm_w;
        const int step_y = (kHmThreads * 4) / hm_w, step_x = kHmThreads * 4 - step_y * hm_w;
        for (int i = threadIdx.x; i < (hw >> 2); i += kHmThreads) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            // almost every vector misses the (6*sigma+1)^2 window: zeros without per-element work
            if (!(x + 4 <= hm_w && (y < y0i || y >= y1i || x + 4 <= x0i || x >= x1i))) {
                int xx = x, yy = y;
                float f[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    f[e] = value(xx, yy);
                    if (++xx == hm_w) { xx = 0; ++yy; }
                }
                v = pack16<float>(f);
            }
            stg_stream(t4 + i, v);
            x += step_x; y += step_y;
            if (x >= hm_w) { x -= hm_w; ++y; }
        }
    } else {
        for (int i = threadIdx.x; i < hw; i += kHmThreads) {
            const int y = i / hm_w, x = i - y * hm_w;
            t[i] = value(x, y);
        }
    }
}

__global__ void __launch_bounds__(kHmThreads)
gauss_target_kernel(const double* __restrict__ joints, const float* __restrict__ vis, int hm_w,
                    int hm_h, double stride_x, double stride_y, TargetWindow tw,
                    float* __restrict__ target, float* __restrict__ weight) {
    gauss_target_plane(blockIdx.x, joints, vis, hm_w, hm_h, stride_x, stride_y, tw, target, weight);
}

// All target sets a loader builds per sample in ONE launch (rendered_hand_pose_mt.py:99,103,115,134,147: the
// student's, the un-augmented and the teacher view's 64x64 targets plus two 8x8 "small" targets — five
// generate_target calls per sample): blockIdx.y picks the set, blockIdx.x the (sample, joint) plane.
struct TargetJobs {
    udape_target_job job[UDAPE_MAX_TARGET_JOBS];
    double stride_x[UDAPE_MAX_TARGET_JOBS], stride_y[UDAPE_MAX_TARGET_JOBS];
};

__global__ void __launch_bounds__(kHmThreads)
gauss_target_multi_kernel(TargetJobs jobs, TargetWindow tw) {
    const udape_target_job& j = jobs.job[blockIdx.y];
    gauss_target_plane(blockIdx.x, j.joints, j.vis, j.hm_w, j.hm_h, jobs.stride_x[blockIdx.y], jobs.stride_y[blockIdx.y], tw,
                       j.target, j.weight);
}

struct LabelWindow {
    float tmp;     // 3*sigma in float32 (int32 tensor - python float -> float32, util.py:333-334)
    int n;         // len(arange(0, 6*sigma+1, 1))
    double x0;     // (6*sigma+1) // 2
    double sigma;
    int kind;      // 0 Gaussian, 1 Cauchy
};

// gate (optional): 0 = the caller skips this joint (real_animal_all_mt.py:275 `if tpts[i, 1] > 0`): the plane
// stays as it is (zero when zero_fill) and vis_out is 1 — the weight the caller multiplies is left alone
__device__ __forceinline__ void
labelmap_plane(const int64_t plane, const int32_t* __restrict__ pts, const uint8_t* __restrict__ gate, int h, int w,
               const LabelWindow& lw, int zero_fill, float* __restrict__ img, int32_t* __restrict__ vis_out) {
    __shared__ float s_tab[kTabN * kTabN];
    auto window = [&](int gx, int gy) -> float {
        const double dx = static_cast<double>(gx) - lw.x0, dy = static_cast<double>(gy) - lw.x0;
        const double d2 = dx * dx + dy * dy;
        const double s2 = lw.sigma * lw.sigma;
        // float64 like numpy, rounded to float32 on store (util.py:349-352,362)
        const double g = lw.kind == 0 ? exp(-d2 / (2.0 * s2)) : lw.sigma / pow(d2 + s2, 1.5);
        return static_cast<float>(g);
    };
    const bool tabulated = lw.n <= kTabN;
    if (tabulated) {
        for (int i = threadIdx.x; i < lw.n * lw.n; i += kHmThreads) s_tab[i] = window(i % lw.n, i / lw.n);
        __syncthreads();
    }
    const int px = pts[2 * plane], py = pts[2 * plane + 1];
    // util.py:333-334: int(pt - 3*sigma), int(pt + 3*sigma + 1) evaluated in float32
    const int ul_x = static_cast<int>(static_cast<float>(px) - lw.tmp);
    const int ul_y = static_cast<int>(static_cast<float>(py) - lw.tmp);
    const int br_x = static_cast<int>(static_cast<float>(px) + lw.tmp + 1.0f);
    const int br_y = static_cast<int>(static_cast<float>(py) + lw.tmp + 1.0f);
    // util.py:337-340: reject any window that touches the border
    const bool gated_off = gate != nullptr && gate[plane] == 0;
    const bool reject = gated_off || br_x >= w || br_y >= h || ul_x < 0 || ul_y < 0;
    if (threadIdx.x == 0 && vis_out) vis_out[plane] = (reject && !gated_off) ? 0 : 1;
    const int x0i = reject ? 0 : max(0, ul_x), x1i = reject ? 0 : min(br_x, w);
    const int y0i = reject ? 0 : max(0, ul_y), y1i = reject ? 0 : min(br_y, h);
    const int hw = h * w;
    float* t = img + plane * static_cast<int64_t>(hw);
    auto value = [&](int x, int y) -> float {  // only called where inside(x, y)
        return tabulated ? s_tab[(y - ul_y) * lw.n + (x - ul_x)] : window(x - ul_x, y - ul_y);
    };
    auto inside = [&](int x, int y) -> bool {
        return x >= x0i && x < x1i && y >= y0i && y < y1i && (x - ul_x) < lw.n && (y - ul_y) < lw.n;
    };
    if (zero_fill) {
        if ((hw & 3) == 0 && aligned16(img)) {
            uint4* t4 = reinterpret_cast<uint4*>(t);
            int y = (threadIdx.x * 4) / w, x = threadIdx.x * 4 - y * w;
            const int step_y = (kHmThreads * 4) / w, step_x = kHmThreads * 4 - step_y * w;
            for (int i = threadIdx.x; i < (hw >> 2); i += kHmThreads) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                // the float64 window is evaluated only where a vector overlaps it
                if (!(x + 4 <= w && (y < y0i || y >= y1i || x + 4 <= x0i || x >= x1i))) {
                    int xx = x, yy = y;
                    float f[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        f[e] = inside(xx, yy) ? value(xx, yy) : 0.0f;
                        if (++xx == w) { xx = 0; ++yy; }
                    }
                    v = pack16<float>(f);
                }
                stg_stream(t4 + i, v);
                x += step_x; y += step_y;
                if (x >= w) { x -= w; ++y; }
            }
        } else {
            for (int i = threadIdx.x; i < hw; i += kHmThreads) {
                const int y = i / w, x = i - y * w;
                t[i] = inside(x, y) ? value(x, y) : 0.0f;
            }
        }
    } else if (!reject) {
        const int ww = x1i - x0i, wh = y1i - y0i;
        for (int i = threadIdx.x; i < ww * wh; i += kHmThreads) {
            const int y = y0i + i / ww, x = x0i + i % ww;
            if (inside(x, y)) t[y * w + x] = value(x, y);
        }
    }
}

__global__ void __launch_bounds__(kHmThreads)
labelmap_kernel(const int32_t* __restrict__ pts, int h, int w, LabelWindow lw, int zero_fill,
                float* __restrict__ img, int32_t* __restrict__ vis_out) {
    labelmap_plane(blockIdx.x, pts, nullptr, h, w, lw, zero_fill, img, vis_out);
}

// The un-augmented, the student's and the teacher view's label maps of a batch in one launch
// (real_animal_all_mt.py:275-283,306-311: draw_labelmap_ori per joint and per view inside `if tpts[i, 1] > 0`).
struct LabelJobs {
    udape_labelmap_job job[UDAPE_MAX_TARGET_JOBS];
};

__global__ void __launch_bounds__(kHmThreads)
labelmap_multi_kernel(LabelJobs jobs, int h, int w, LabelWindow lw) {
    const udape_labelmap_job& j = jobs.job[blockIdx.y];
    labelmap_plane(blockIdx.x, j.pts, j.gate, h, w, lw, 1, j.img, j.vis_out);
}

}  // namespace udape

using namespace udape;

static TargetWindow make_target_window(double sigma) {
    TargetWindow tw;
    tw.tmp = sigma * 3.0;
    const double size = 2.0 * tw.tmp + 1.0;
    tw.n = static_cast<int>(std::ceil(size));
    tw.x0 = static_cast<float>(std::floor(size / 2.0));
    tw.denom = static_cast<float>(2.0 * sigma * sigma);
    return tw;
}

static LabelWindow make_label_window(double sigma, int kind) {
    LabelWindow lw;
    lw.tmp = static_cast<float>(3.0 * sigma);
    const double size = 6.0 * sigma + 1.0;
    lw.n = static_cast<int>(std::ceil(size));
    lw.x0 = std::floor(size / 2.0);
    lw.sigma = sigma;
    lw.kind = kind;
    return lw;
}

extern "C" int udape_gauss_target_multi(const udape_target_job* jobs, int n_jobs, int64_t planes, double sigma,
                                        double image_w, double image_h, void* stream) {
    UDAPE_REQUIRE(jobs, UDAPE_ERR_NULL, "udape_gauss_target_multi: jobs is NULL");
    UDAPE_REQUIRE(n_jobs >= 1 && n_jobs <= UDAPE_MAX_TARGET_JOBS, UDAPE_ERR_ARG, "udape_gauss_target_multi: n_jobs=%d (1..%d)", n_jobs, UDAPE_MAX_TARGET_JOBS);
    UDAPE_REQUIRE(planes > 0 && planes < (1ll << 31), UDAPE_ERR_SHAPE, "udape_gauss_target_multi: bad planes=%lld", (long long)planes);
    UDAPE_REQUIRE(sigma > 0.0 && sigma < 1e4 && image_w > 0.0 && image_h > 0.0, UDAPE_ERR_ARG,
                  "udape_gauss_target_multi: sigma/image size out of range");
    TargetJobs tj;
    for (int i = 0; i < n_jobs; ++i) {
        const udape_target_job& j = jobs[i];
        UDAPE_REQUIRE(j.joints && j.vis && j.target && j.weight, UDAPE_ERR_NULL, "udape_gauss_target_multi: job %d has a NULL pointer", i);
        UDAPE_REQUIRE(j.hm_w > 0 && j.hm_h > 0 && static_cast<int64_t>(j.hm_w) * j.hm_h < (1ll << 31), UDAPE_ERR_SHAPE,
                      "udape_gauss_target_multi: job %d has bad extents w=%d h=%d", i, (int)j.hm_w, (int)j.hm_h);
        UDAPE_REQUIRE(aligned_to(j.joints, 8) && aligned_to(j.vis, 4) && aligned_to(j.target, 4) && aligned_to(j.weight, 4),
                      UDAPE_ERR_ALIGN, "udape_gauss_target_multi: job %d has a misaligned pointer", i);
        tj.job[i] = j;
        tj.stride_x[i] = image_w / static_cast<double>(j.hm_w);   // util.py:37: feat_stride = image_size / heatmap_size
        tj.stride_y[i] = image_h / static_cast<double>(j.hm_h);
    }
    const dim3 grid(static_cast<unsigned>(planes), static_cast<unsigned>(n_jobs));
    gauss_target_multi_kernel<<<grid, kHmThreads, 0, as_stream(stream)>>>(tj, make_target_window(sigma));
    return check_launch("udape_gauss_target_multi");
}

extern "C" int udape_labelmap_multi(const udape_labelmap_job* jobs, int n_jobs, int64_t planes, int64_t h, int64_t w,
                                    double sigma, int kind, void* stream) {
    UDAPE_REQUIRE(jobs, UDAPE_ERR_NULL, "udape_labelmap_multi: jobs is NULL");
    UDAPE_REQUIRE(n_jobs >= 1 && n_jobs <= UDAPE_MAX_TARGET_JOBS, UDAPE_ERR_ARG, "udape_labelmap_multi: n_jobs=%d (1..%d)", n_jobs, UDAPE_MAX_TARGET_JOBS);
    UDAPE_REQUIRE(planes > 0 && h > 0 && w > 0 && planes < (1ll << 31) && h * w < (1ll << 31), UDAPE_ERR_SHAPE,
                  "udape_labelmap_multi: bad extents planes=%lld h=%lld w=%lld", (long long)planes, (long long)h, (long long)w);
    UDAPE_REQUIRE(sigma > 0.0 && sigma < 1e4, UDAPE_ERR_ARG, "udape_labelmap_multi: sigma %g out of range", sigma);
    UDAPE_REQUIRE(kind == 0 || kind == 1, UDAPE_ERR_ARG, "udape_labelmap_multi: kind must be 0 (Gaussian) or 1 (Cauchy)");
    LabelJobs lj;
    for (int i = 0; i < n_jobs; ++i) {
        const udape_labelmap_job& j = jobs[i];
        UDAPE_REQUIRE(j.pts && j.img, UDAPE_ERR_NULL, "udape_labelmap_multi: job %d has a NULL pointer", i);
        UDAPE_REQUIRE(aligned_to(j.pts, 4) && aligned_to(j.img, 4) && (!j.vis_out || aligned_to(j.vis_out, 4)), UDAPE_ERR_ALIGN,
                      "udape_labelmap_multi: job %d has a misaligned pointer", i);
        lj.job[i] = j;
    }
    const dim3 grid(static_cast<unsigned>(planes), static_cast<unsigned>(n_jobs));
    labelmap_multi_kernel<<<grid, kHmThreads, 0, as_stream(stream)>>>(lj, static_cast<int>(h), static_cast<int>(w),
                                                                     make_label_window(sigma, kind));
    return check_launch("udape_labelmap_multi");
}

extern "C" int udape_gauss_target(const double* joints, const float* vis, int64_t planes, int64_t hm_w,
                                  int64_t hm_h, double sigma, double image_w, double image_h,
                                  float* target, float* weight, void* stream) {
    UDAPE_REQUIRE(joints && vis && target && weight, UDAPE_ERR_NULL, "udape_gauss_target: NULL pointer");
    UDAPE_REQUIRE(planes > 0 && hm_w > 0 && hm_h > 0 && planes < (1ll << 31) && hm_w * hm_h < (1ll << 31),
                  UDAPE_ERR_SHAPE, "udape_gauss_target: bad extents planes=%lld w=%lld h=%lld",
                  (long long)planes, (long long)hm_w, (long long)hm_h);
    UDAPE_REQUIRE(sigma > 0.0 && sigma < 1e4 && image_w > 0.0 && image_h > 0.0, UDAPE_ERR_ARG,
                  "udape_gauss_target: sigma/image size out of range");
    UDAPE_REQUIRE(aligned_to(joints, 8) && aligned_to(vis, 4) && aligned_to(target, 4) && aligned_to(weight, 4),
                  UDAPE_ERR_ALIGN, "udape_gauss_target: misaligned pointer");
    const TargetWindow tw = make_target_window(sigma);
    // util.py:37: feat_stride = image_size / heatmap_size (float64)
    const double stride_x = image_w / static_cast<double>(hm_w), stride_y = image_h / static_cast<double>(hm_h);
    gauss_target_kernel<<<static_cast<unsigned>(planes), kHmThreads, 0, as_stream(stream)>>>(
        joints, vis, static_cast<int>(hm_w), static_cast<int>(hm_h), stride_x, stride_y, tw, target, weight);
    return check_launch("udape_gauss_target");
}

extern "C" int udape_labelmap(const int32_t* pts, int64_t planes, int64_t h, int64_t w, double sigma,
                              int kind, int zero_fill, float* img, int32_t* vis_out, void* stream) {
    UDAPE_REQUIRE(pts && img, UDAPE_ERR_NULL, "udape_labelmap: NULL pointer");
    UDAPE_REQUIRE(planes > 0 && h > 0 && w > 0 && planes < (1ll << 31) && h * w < (1ll << 31), UDAPE_ERR_SHAPE,
                  "udape_labelmap: bad extents planes=%lld h=%lld w=%lld", (long long)planes, (long long)h, (long long)w);
    UDAPE_REQUIRE(sigma > 0.0 && sigma < 1e4, UDAPE_ERR_ARG, "udape_labelmap: sigma %g out of range", sigma);
    UDAPE_REQUIRE(kind == 0 || kind == 1, UDAPE_ERR_ARG, "udape_labelmap: kind must be 0 (Gaussian) or 1 (Cauchy)");
    UDAPE_REQUIRE(aligned_to(pts, 4) && aligned_to(img, 4) && (!vis_out || aligned_to(vis_out, 4)), UDAPE_ERR_ALIGN,
                  "udape_labelmap: misaligned pointer");
    const LabelWindow lw = make_label_window(sigma, kind);
    labelmap_kernel<<<static_cast<unsigned>(planes), kHmThreads, 0, as_stream(stream)>>>(
        pts, static_cast<int>(h), static_cast<int>(w), lw, zero_fill, img, vis_out);
    return check_launch("udape_labelmap");
}
