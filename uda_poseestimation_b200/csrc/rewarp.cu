// rewarp.cu — batched multi-stage nearest-neighbour affine re-warp of heatmaps / images.
//
// Replaces the per-sample Python loops of the reference's train():
//   teacher recon   train_human.py:361-372  (k views, three tF.affine calls each, mean over views)
//   student recon   train_human.py:418-423  (three tF.affine calls, under autocast; needs backward)
//   occlusion       train_human.py:385-412  (three-stage warp, patch paste, one-stage warp back)
// (identical in train_animal.py:386-397,443-448,410-437).  Each tF.affine(nearest) call is
// torchvision's `_gen_affine_grid` + `grid_sample(mode="nearest", padding_mode="zeros",
// align_corners=False)`; a chain of nearest-neighbour resamplings is a composition of integer
// source-index maps, so the whole chain is ONE gather:   out[p] = in[s1(s2(s3(p)))]   (0 when any
// stage leaves the image).  The per-stage arithmetic is reproduced in the reference's float32 op
// order (oracle/reference_port.py::affine_source_index restates it and is pinned against
// torchvision on the CPU):
//
//   x = i + (-W/2 + 0.5), y = j + (-H/2 + 0.5)                       (torch.linspace, exact)
//   g = fma(y, r1, x*r0) + r2           per axis, r = theta / (0.5*size) rounded in the grid dtype
//   [g rounded to fp16/bf16 when the stage's theta was built from a half image under autocast]
//   ix = ((g + 1) * W - 1) / 2 ; nearest = rint(ix) (ties to even) ; valid iff 0 <= nearest <= W-1
//
// The source index of a pixel depends only on the sample, not on the channel: every thread
// computes the composed index of its PX consecutive pixels once (registers) and then loops over
// the channels of its CTA — gathers from the L1/L2-resident source plane, 128-bit coalesced
// streaming stores.  HBM-bound: one read and one write of the tensor.
//
// Backward (student recon): grad_in[s] = sum of grad_out[p] over {p : src(p) = s}.  No float
// atomics: the CTA inverts the composed map in shared memory (integer counting sort, lists sorted
// by p) once per sample and every channel then sums its list in ascending p — deterministic.
#include "common.cuh"

namespace udape {

constexpr int kRwThreads = 256;
constexpr int kRwMaxStages = 4;
constexpr int kRwMaxViews = 4;

struct RewarpView {
    const void* in;      // [B, C, H, W]
    const float* theta;  // [B, stages, 6] rescaled thetas in EVALUATION order (last applied stage first)
};

struct RewarpArgs {
    RewarpView view[kRwMaxViews];
    const int32_t* paste;   // [B, 6] = row0,row1,col0,col1 (destination), src_row0, src_col0; or NULL
    const uint8_t* active;  // [B] 0 = copy the sample through unchanged; or NULL (all active)
    int views, stages;
    int half_mask;          // bit s: stage s (evaluation order) rounds its grid to `grid_dtype`
    int grid_dtype;         // UDAPE_F16 or UDAPE_BF16
    int paste_after;        // the paste remap is applied after this many evaluated stages
    int B, C, H, W;
    int cpc;                // channels per CTA
};

__device__ __forceinline__ float round_grid(float v, int grid_dtype) {
    return grid_dtype == UDAPE_F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v));
}

// one tF.affine(nearest) stage: source pixel of output pixel (i, j), or false when out of bounds
__device__ __forceinline__ bool stage_source(int& i, int& j, const float* __restrict__ r, int W, int H, bool half,
                                             int grid_dtype) {
    float x = static_cast<float>(i) + (0.5f - 0.5f * static_cast<float>(W));
    float y = static_cast<float>(j) + (0.5f - 0.5f * static_cast<float>(H));
    if (half) { x = round_grid(x, grid_dtype); y = round_grid(y, grid_dtype); }
    // bmm over k = 3 with an FMA chain:  x*r0 -> fma(y, r1, .) -> + 1*r2
    float gx = __fadd_rn(__fmaf_rn(y, r[1], __fmul_rn(x, r[0])), r[2]);
    float gy = __fadd_rn(__fmaf_rn(y, r[4], __fmul_rn(x, r[3])), r[5]);
    if (half) { gx = round_grid(gx, grid_dtype); gy = round_grid(gy, grid_dtype); }
    // grid_sampler_unnormalize (align_corners=False) and nearbyint
    const float fx = rintf(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), static_cast<float>(W)), 1.0f), 0.5f));
    const float fy = rintf(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), static_cast<float>(H)), 1.0f), 0.5f));
    if (!(fx >= 0.0f && fx <= static_cast<float>(W - 1) && fy >= 0.0f && fy <= static_cast<float>(H - 1))) return false;
    i = static_cast<int>(fx);
    j = static_cast<int>(fy);
    return true;
}

// composed source index of output pixel p of sample b through one view's stage table (-1: zero)
__device__ __forceinline__ int composed_source(int p, int b, const float* __restrict__ theta, const RewarpArgs& a) {
    int j = p / a.W, i = p - j * a.W;
    const float* r = theta + static_cast<int64_t>(b) * a.stages * 6;
    for (int s = 0; s < a.stages; ++s) {
        if (a.paste && s == a.paste_after) {
            // temp[:, row0:row1, col0:col1] = temp[:, srow0:.., scol0:..]   (train_human.py:409)
            const int32_t* q = a.paste + 6 * b;
            if (j >= q[0] && j < q[1] && i >= q[2] && i < q[3]) { j += q[4] - q[0]; i += q[5] - q[2]; }
        }
        if (!stage_source(i, j, r + 6 * s, a.W, a.H, (a.half_mask >> s) & 1, a.grid_dtype)) return -1;
    }
    return j * a.W + i;
}

template <typename T, int PX>
__global__ void __launch_bounds__(kRwThreads)
rewarp_fwd_kernel(const RewarpArgs a, T* __restrict__ out) {
    const int hw = a.H * a.W;
    const int bands = (hw + kRwThreads * PX - 1) / (kRwThreads * PX);
    const int cgroups = (a.C + a.cpc - 1) / a.cpc;
    int bid = blockIdx.x;
    const int band = bid % bands; bid /= bands;
    const int cg = bid % cgroups;
    const int b = bid / cgroups;
    const int p0 = (band * kRwThreads + threadIdx.x) * PX;
    if (p0 >= hw) return;
    const int c0 = cg * a.cpc, c1 = min(a.C, c0 + a.cpc);
    const bool through = a.active && a.active[b] == 0;
    int src[kRwMaxViews][PX];
#pragma unroll
    for (int v = 0; v < kRwMaxViews; ++v) {
        if (v < a.views) {
#pragma unroll
            for (int e = 0; e < PX; ++e)
                src[v][e] = through ? p0 + e : (p0 + e < hw ? composed_source(p0 + e, b, a.view[v].theta, a) : -1);
        }
    }
    const float nviews = static_cast<float>(a.views);
#pragma unroll 2
    for (int c = c0; c < c1; ++c) {
        const int64_t base = (static_cast<int64_t>(b) * a.C + c) * hw;
        float f[PX];
#pragma unroll
        for (int e = 0; e < PX; ++e) f[e] = 0.0f;
#pragma unroll
        for (int v = 0; v < kRwMaxViews; ++v) {
            if (v < a.views) {
                const T* in = static_cast<const T*>(a.view[v].in) + base;
#pragma unroll
                for (int e = 0; e < PX; ++e) {
                    const float val = src[v][e] >= 0 ? to_f32<T>(in[src[v][e]]) : 0.0f;
                    f[e] = v == 0 ? val : f[e] + val;   // torch.mean over the k views: sequential sum ...
                }
            }
        }
        if (a.views > 1) {
#pragma unroll
            for (int e = 0; e < PX; ++e) f[e] = __fdiv_rn(f[e], nviews);  // ... then one division (CPU mean)
        }
        if (PX > 1) {
            stg_stream(out + base + p0, pack16<T>(f));
        } else {
            out[base + p0] = from_f32<T>(f[0]);
        }
    }
}

// ---- backward --------------------------------------------------------------------------------
// shared memory: uint32 off[hw + 1] | uint16 map[hw] | uint16 lst[hw]
template <typename T>
__global__ void __launch_bounds__(kRwThreads)
rewarp_bwd_kernel(const RewarpArgs a, const T* __restrict__ gout, T* __restrict__ gin) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int hw = a.H * a.W;
    uint32_t* off = reinterpret_cast<uint32_t*>(smem_raw);
    uint16_t* map = reinterpret_cast<uint16_t*>(off + hw + 1 + ((hw + 1) & 1));
    uint16_t* lst = map + hw;
    __shared__ uint32_t s_scan[kRwThreads];
    const int cgroups = (a.C + a.cpc - 1) / a.cpc;
    const int cg = blockIdx.x % cgroups, b = blockIdx.x / cgroups;
    const int c0 = cg * a.cpc, c1 = min(a.C, c0 + a.cpc);
    // 1. composed map and per-source counts (integer atomics: order-independent)
    for (int s = threadIdx.x; s <= hw; s += kRwThreads) off[s] = 0u;
    __syncthreads();
    for (int p = threadIdx.x; p < hw; p += kRwThreads) {
        const int s = composed_source(p, b, a.view[0].theta, a);
        map[p] = s < 0 ? 0xffffu : static_cast<uint16_t>(s);
        if (s >= 0) atomicAdd(&off[s], 1u);
    }
    __syncthreads();
    // 2. exclusive scan of the counts (each thread owns a contiguous run)
    const int per = (hw + kRwThreads - 1) / kRwThreads;
    const int lo = min(hw, static_cast<int>(threadIdx.x) * per), hi = min(hw, lo + per);
    uint32_t run = 0;
    for (int s = lo; s < hi; ++s) run += off[s];
    s_scan[threadIdx.x] = run;
    __syncthreads();
    if (threadIdx.x < 32) {
        // 256 partials: 8 per lane, warp scan
        uint32_t part[kRwThreads / 32], tot = 0;
#pragma unroll
        for (int q = 0; q < kRwThreads / 32; ++q) { part[q] = tot; tot += s_scan[threadIdx.x * (kRwThreads / 32) + q]; }
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (static_cast<int>(threadIdx.x) >= o) incl += t;
        }
        const uint32_t excl = incl - tot;
#pragma unroll
        for (int q = 0; q < kRwThreads / 32; ++q) s_scan[threadIdx.x * (kRwThreads / 32) + q] = excl + part[q];
    }
    __syncthreads();
    run = s_scan[threadIdx.x];
    for (int s = lo; s < hi; ++s) { const uint32_t c = off[s]; off[s] = run; run += c; }
    __syncthreads();
    // 3. fill the lists (slot order is arbitrary here ...); afterwards off[s] = end of list s
    for (int p = threadIdx.x; p < hw; p += kRwThreads) {
        const uint16_t s = map[p];
        if (s != 0xffffu) lst[atomicAdd(&off[s], 1u)] = static_cast<uint16_t>(p);
    }
    __syncthreads();
    // 4. ... so sort every list by p (insertion sort; lists hold ~1/scale^2 entries)
    for (int s = threadIdx.x; s < hw; s += kRwThreads) {
        const int st = s == 0 ? 0 : static_cast<int>(off[s - 1]), en = static_cast<int>(off[s]);
        for (int q = st + 1; q < en; ++q) {
            const uint16_t v = lst[q];
            int r = q - 1;
            while (r >= st && lst[r] > v) { lst[r + 1] = lst[r]; --r; }
            lst[r + 1] = v;
        }
    }
    __syncthreads();
    // 5. every channel: grad_in[s] = sum over its list in ascending p (fp32, one rounding to T)
    for (int c = c0; c < c1; ++c) {
        const int64_t base = (static_cast<int64_t>(b) * a.C + c) * hw;
        const T* go = gout + base;
        for (int s = threadIdx.x; s < hw; s += kRwThreads) {
            const int st = s == 0 ? 0 : static_cast<int>(off[s - 1]), en = static_cast<int>(off[s]);
            float acc = 0.0f;
            for (int q = st; q < en; ++q) acc += to_f32<T>(go[lst[q]]);
            gin[base + s] = from_f32<T>(acc);
        }
    }
}

static int channels_per_cta(int64_t units, int64_t c) {
    // aim at ~16 CTAs per SM worth of work items, amortising the index computation over channels
    int64_t cpc = units * c / (148 * 16);
    if (cpc < 1) cpc = 1;
    if (cpc > c) cpc = c;
    const int64_t groups = (c + cpc - 1) / cpc;
    return static_cast<int>((c + groups - 1) / groups);  // even split
}

static int fill_common(RewarpArgs& a, const char* name, int views, int stages, int half_mask, int grid_dtype,
                       int64_t B, int64_t C, int64_t H, int64_t W, int dtype) {
    UDAPE_REQUIRE(views >= 1 && views <= kRwMaxViews, UDAPE_ERR_ARG, "%s: views must be in [1, %d]", name, kRwMaxViews);
    UDAPE_REQUIRE(stages >= 1 && stages <= kRwMaxStages, UDAPE_ERR_ARG, "%s: stages must be in [1, %d]", name, kRwMaxStages);
    UDAPE_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && H * W < (1ll << 30) && B * C < (1ll << 31) && H <= 1024 && W <= 1024,
                  UDAPE_ERR_SHAPE, "%s: bad extents B=%lld C=%lld H=%lld W=%lld (H, W <= 1024)", name, (long long)B,
                  (long long)C, (long long)H, (long long)W);
    UDAPE_REQUIRE(dtype_size(dtype) == 2 || dtype_size(dtype) == 4, UDAPE_ERR_DTYPE, "%s: unsupported dtype code %d", name, dtype);
    UDAPE_REQUIRE(half_mask == 0 || grid_dtype == UDAPE_F16 || grid_dtype == UDAPE_BF16, UDAPE_ERR_DTYPE,
                  "%s: half-precision grid stages need grid_dtype F16 or BF16", name);
    UDAPE_REQUIRE(half_mask >= 0 && half_mask < (1 << stages), UDAPE_ERR_ARG, "%s: half_mask has bits beyond the stages", name);
    a.views = views; a.stages = stages; a.half_mask = half_mask; a.grid_dtype = grid_dtype;
    a.B = static_cast<int>(B); a.C = static_cast<int>(C); a.H = static_cast<int>(H); a.W = static_cast<int>(W);
    a.paste = nullptr; a.active = nullptr; a.paste_after = 0;
    return UDAPE_OK;
}

}  // namespace udape

using namespace udape;

extern "C" int udape_rewarp_fwd(const void* const* in, const float* const* theta, int views, int stages,
                                int half_mask, int grid_dtype, const int32_t* paste, int paste_after,
                                const uint8_t* active, int64_t B, int64_t C, int64_t H, int64_t W, int dtype,
                                void* out, void* stream) {
    RewarpArgs a = {};
    UDAPE_REQUIRE(in && theta && out, UDAPE_ERR_NULL, "udape_rewarp_fwd: NULL pointer");
    const int rc = fill_common(a, "udape_rewarp_fwd", views, stages, half_mask, grid_dtype, B, C, H, W, dtype);
    if (rc) return rc;
    const int es = dtype_size(dtype);
    for (int v = 0; v < views; ++v) {
        UDAPE_REQUIRE(in[v] && theta[v], UDAPE_ERR_NULL, "udape_rewarp_fwd: NULL view %d", v);
        UDAPE_REQUIRE(aligned_to(in[v], es) && aligned_to(theta[v], 4), UDAPE_ERR_ALIGN, "udape_rewarp_fwd: misaligned view %d", v);
        UDAPE_REQUIRE(in[v] != out, UDAPE_ERR_ARG, "udape_rewarp_fwd: the gather cannot run in place");
        a.view[v].in = in[v];
        a.view[v].theta = theta[v];
    }
    UDAPE_REQUIRE(aligned_to(out, es) && (!paste || aligned_to(paste, 4)), UDAPE_ERR_ALIGN, "udape_rewarp_fwd: misaligned pointer");
    UDAPE_REQUIRE(!paste || (paste_after >= 0 && paste_after < stages), UDAPE_ERR_ARG,
                  "udape_rewarp_fwd: paste_after must name an evaluated stage");
    UDAPE_REQUIRE(!active || views == 1, UDAPE_ERR_ARG, "udape_rewarp_fwd: pass-through samples need a single view");
    a.paste = paste; a.paste_after = paste_after; a.active = active;
    const int64_t hw = H * W;
    cudaStream_t st = as_stream(stream);
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        constexpr int EPV = Vec16<T>::EPV;
        const bool vec = (hw % EPV) == 0 && aligned16(out);
        const int px = vec ? EPV : 1;
        const int64_t bands = (hw + kRwThreads * px - 1) / (kRwThreads * px);
        a.cpc = channels_per_cta(B * bands, C);
        const int64_t grid = B * bands * ((C + a.cpc - 1) / a.cpc);
        UDAPE_REQUIRE(grid < (1ll << 31), UDAPE_ERR_SHAPE, "udape_rewarp_fwd: grid too large");
        if (vec) rewarp_fwd_kernel<T, EPV><<<static_cast<unsigned>(grid), kRwThreads, 0, st>>>(a, static_cast<T*>(out));
        else rewarp_fwd_kernel<T, 1><<<static_cast<unsigned>(grid), kRwThreads, 0, st>>>(a, static_cast<T*>(out));
    });
    return check_launch("udape_rewarp_fwd");
}

extern "C" int udape_rewarp_bwd(const void* grad_out, const float* theta, int stages, int half_mask, int grid_dtype,
                                int64_t B, int64_t C, int64_t H, int64_t W, int dtype, void* grad_in, void* stream) {
    RewarpArgs a = {};
    UDAPE_REQUIRE(grad_out && theta && grad_in, UDAPE_ERR_NULL, "udape_rewarp_bwd: NULL pointer");
    const int rc = fill_common(a, "udape_rewarp_bwd", 1, stages, half_mask, grid_dtype, B, C, H, W, dtype);
    if (rc) return rc;
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(aligned_to(grad_out, es) && aligned_to(grad_in, es) && aligned_to(theta, 4), UDAPE_ERR_ALIGN,
                  "udape_rewarp_bwd: misaligned pointer");
    UDAPE_REQUIRE(grad_out != grad_in, UDAPE_ERR_ARG, "udape_rewarp_bwd: cannot run in place");
    const int64_t hw = H * W;
    // the inverted map lives in shared memory: 8 bytes per pixel
    UDAPE_REQUIRE(hw <= 25600, UDAPE_ERR_SHAPE, "udape_rewarp_bwd: planes above 25600 pixels are not supported (H*W=%lld)",
                  (long long)hw);
    a.view[0].in = grad_out;
    a.view[0].theta = theta;
    a.cpc = channels_per_cta(B, C);
    const int64_t grid = B * ((C + a.cpc - 1) / a.cpc);
    const size_t smem = sizeof(uint32_t) * (hw + 1 + ((hw + 1) & 1)) + 2 * sizeof(uint16_t) * hw;
    cudaStream_t st = as_stream(stream);
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(rewarp_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem));
            if (e != cudaSuccess) return fail(static_cast<int>(e), "udape_rewarp_bwd: cannot reserve %zu bytes of shared memory", smem);
        }
        rewarp_bwd_kernel<T><<<static_cast<unsigned>(grid), kRwThreads, smem, st>>>(
            a, static_cast<const T*>(grad_out), static_cast<T*>(grad_in));
    });
    return check_launch("udape_rewarp_bwd");
}
