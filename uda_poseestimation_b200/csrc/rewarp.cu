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
// computes the composed index of its pixels once (16-bit offsets packed in registers) and then
// loops over the channels of its CTA.  HBM-bound: one read and one write of the tensor.
//
//   heatmap route (rewarp_smem_kernel, planes up to 4096 px fp32 / 8192 px half): a warp-wide
//     gather along a rotated line touches up to 32 different L1 lines, so gathering from global
//     memory is bound by L1 wavefronts, not by HBM.  Instead the source plane is staged into
//     shared memory with coalesced 128-bit loads (register double buffering: the next plane's loads
//     are in flight while the current one is gathered), rows padded to a stride of +-4 (mod 32)
//     banks — the sign is chosen per sample from the composed map's direction so that the 32 lanes
//     of a gather (consecutive output pixels) spread over the banks for every rotation angle —
//     and consecutive lanes write consecutive pixels (128-byte coalesced stores).
//   general route (rewarp_fwd_kernel): gathers straight from global memory; used for image-sized
//     planes (occlusion, 3x256x256), ragged widths and the paste / pass-through options.
//
// Backward (student recon): grad_in[s] = sum of grad_out[p] over {p : src(p) = s}, float32 sums in ascending p,
// one rounding; no float atomics anywhere.
//   with a plan (what autograd and the step use): the forward side builds, once per batch, a PUSH PLAN — a
//     shared-memory slot for every output pixel — and the backward is a branch-free scatter into those slots,
//     one ordered fold per source pixel that has several contributors, and a read-out (rewarp_push_plan_kernel,
//     rewarp_bwd_push2_kernel / rewarp_bwd_push_kernel below);
//   without one: the CTA inverts the composed map in shared memory (integer counting sort, lists sorted by p)
//     once per sample and every channel sums its lists, staged through the padded buffers (the fallback).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"
#include "pipeline.cuh"
#include "select.cuh"

namespace cg = cooperative_groups;

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
    int tile_log2;          // wide2 route, 2-byte planes: a warp's 32 output words form a tile 2^tile_log2 words wide (0: a row walk)
};

__device__ __forceinline__ float round_grid(float v, int grid_dtype) {
    return grid_dtype == UDAPE_F16 ? __half2float(__float2half_rn(v)) : __bfloat162float(__float2bfloat16_rn(v));
}
template <int GD> __device__ __forceinline__ float round_grid_t(float v, int grid_dtype) {
    if (GD == UDAPE_F16) return __half2float(__float2half_rn(v));
    if (GD == UDAPE_BF16) return __bfloat162float(__float2bfloat16_rn(v));
    return round_grid(v, grid_dtype);
}

// one tF.affine(nearest) stage: source pixel of output pixel (i, j), or false when out of bounds.
// HM: 0 = no stage rounds its grid, 1 = this stage does, 2 = decided at run time by `half`;
// GD: the grid dtype when it is known at compile time (UDAPE_F16 / UDAPE_BF16), -1 = run time.
// The composed map costs ~75 instructions per stage and pixel when both variants are compiled in under
// predicates; it was 59 % of the gather kernel's executed instructions at 21 channels per sample
// (profiles/r01v_rewarp_ncu_summary.txt), so the common cases are specialised.
template <int HM, int GD>
__device__ __forceinline__ bool stage_source_t(int& i, int& j, const float* __restrict__ r, int W, int H, bool half,
                                               int grid_dtype) {
    const bool rnd = HM == 2 ? half : (HM == 1);
    float x = static_cast<float>(i) + (0.5f - 0.5f * static_cast<float>(W));
    float y = static_cast<float>(j) + (0.5f - 0.5f * static_cast<float>(H));
    if (rnd) { x = round_grid_t<GD>(x, grid_dtype); y = round_grid_t<GD>(y, grid_dtype); }
    // bmm over k = 3 with an FMA chain:  x*r0 -> fma(y, r1, .) -> + 1*r2
    float gx = __fadd_rn(__fmaf_rn(y, r[1], __fmul_rn(x, r[0])), r[2]);
    float gy = __fadd_rn(__fmaf_rn(y, r[4], __fmul_rn(x, r[3])), r[5]);
    if (rnd) { gx = round_grid_t<GD>(gx, grid_dtype); gy = round_grid_t<GD>(gy, grid_dtype); }
    // grid_sampler_unnormalize (align_corners=False), then nearbyint: F2I.RN rounds ties to even like
    // rint and saturates, so the range test on the integers equals ATen's test on the rounded floats
    const float fx = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), static_cast<float>(W)), 1.0f), 0.5f);
    const float fy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), static_cast<float>(H)), 1.0f), 0.5f);
    const int xi = __float2int_rn(fx), yi = __float2int_rn(fy);
    if (static_cast<unsigned>(xi) >= static_cast<unsigned>(W) || static_cast<unsigned>(yi) >= static_cast<unsigned>(H) ||
        fx != fx || fy != fy)
        return false;
    i = xi;
    j = yi;
    return true;
}
__device__ __forceinline__ bool stage_source(int& i, int& j, const float* __restrict__ r, int W, int H, bool half,
                                             int grid_dtype) {
    return stage_source_t<2, -1>(i, j, r, W, H, half, grid_dtype);
}

// The same stage with the pixel carried as FLOATS (exact small integers): conversions run at a quarter of the
// FP32 rate on the SM and were most of the ~350 issue slots a pixel of a half-grid map cost (per stage two
// int->float, two float->int and eight half roundings), so
//   * the nearest index stays a float (rintf == F2I.RN for every in-range value; the range test on the floats
//     rejects NaN / inf / out-of-range exactly like the test on the saturated integers),
//   * the rounding of x = i + (0.5 - W/2) to the grid dtype is skipped where it is the identity: x is a
//     half-integer below 512 in magnitude (W, H <= 1024), i.e. 2x has at most 10 bits — exact in fp16 (11-bit
//     significand) always, in bf16 (8 bits) when W, H <= 256 (`xy_exact`).
// Bit-identical to stage_source_t (tests/test_gpu_rewarp.py compares every route against torchvision).
template <int HM, int GD>
__device__ __forceinline__ bool stage_source_f(float& fi, float& fj, const float* __restrict__ r, float Wf, float Hf, float cx,
                                               float cy, bool half, int grid_dtype, bool xy_exact) {
    const bool rnd = HM == 2 ? half : (HM == 1);
    float x = fi + cx, y = fj + cy;
    if (rnd && !xy_exact) { x = round_grid_t<GD>(x, grid_dtype); y = round_grid_t<GD>(y, grid_dtype); }
    float gx = __fadd_rn(__fmaf_rn(y, r[1], __fmul_rn(x, r[0])), r[2]);
    float gy = __fadd_rn(__fmaf_rn(y, r[4], __fmul_rn(x, r[3])), r[5]);
    if (rnd) { gx = round_grid_t<GD>(gx, grid_dtype); gy = round_grid_t<GD>(gy, grid_dtype); }
    const float fx = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.0f), Wf), 1.0f), 0.5f);
    const float fy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.0f), Hf), 1.0f), 0.5f);
    const float xr = rintf(fx), yr = rintf(fy);
    if (!(xr >= 0.0f && xr <= Wf - 1.0f && yr >= 0.0f && yr <= Hf - 1.0f)) return false;   // NaN fails every comparison
    fi = xr;
    fj = yr;
    return true;
}
__device__ __forceinline__ bool grid_xy_exact(int W, int H, int grid_dtype) {
    return grid_dtype == UDAPE_F16 || (W <= 256 && H <= 256);
}

// a sample's stage table held in registers (rows beyond `stages` are never read)
struct StageRegs { float r[kRwMaxStages][6]; };
__device__ __forceinline__ void load_stages(StageRegs& R, const float* __restrict__ r, int stages) {
#pragma unroll
    for (int s = 0; s < kRwMaxStages; ++s)
#pragma unroll
        for (int k = 0; k < 6; ++k) R.r[s][k] = s < stages ? r[6 * s + k] : 0.0f;
}
// composed source pixel of output pixel (i, j), in place (false: zero fill); stages unrolled over the
// register table
template <int HM, int GD>
__device__ __forceinline__ bool composed_source_ij_t(int& i, int& j, const StageRegs& R, const RewarpArgs& a) {
    const float Wf = static_cast<float>(a.W), Hf = static_cast<float>(a.H);
    const float cx = 0.5f - 0.5f * Wf, cy = 0.5f - 0.5f * Hf;
    const bool xy_exact = GD == UDAPE_F16 ? true : grid_xy_exact(a.W, a.H, a.grid_dtype);
    float fi = static_cast<float>(i), fj = static_cast<float>(j);
#pragma unroll
    for (int s = 0; s < kRwMaxStages; ++s) {
        if (s < a.stages) {
            if (!stage_source_f<HM, GD>(fi, fj, R.r[s], Wf, Hf, cx, cy, (a.half_mask >> s) & 1, a.grid_dtype, xy_exact)) return false;
        }
    }
    i = static_cast<int>(fi);
    j = static_cast<int>(fj);
    return true;
}
// kernel-uniform dispatch: no rounding (float32 images), every stage on a half grid (the student under
// autocast), or the per-stage mask (anything else)
__device__ __forceinline__ bool composed_source_ij(int& i, int& j, const StageRegs& R, const RewarpArgs& a) {
    if (a.half_mask == 0) return composed_source_ij_t<0, -1>(i, j, R, a);
    if (a.half_mask == (1 << a.stages) - 1) {
        if (a.grid_dtype == UDAPE_F16) return composed_source_ij_t<1, UDAPE_F16>(i, j, R, a);
        return composed_source_ij_t<1, UDAPE_BF16>(i, j, R, a);
    }
    return composed_source_ij_t<2, -1>(i, j, R, a);
}

// composed source index of output pixel p through one sample's stage table r[stages][6] (-1: zero).
// q (optional) is the sample's paste record: temp[:, row0:row1, col0:col1] = temp[:, srow0:.., scol0:..]
__device__ __forceinline__ int composed_source(int p, const float* __restrict__ r, const int32_t* __restrict__ q,
                                               const RewarpArgs& a) {
    int j = p / a.W, i = p - j * a.W;
    for (int s = 0; s < a.stages; ++s) {
        if (q && s == a.paste_after) {   // train_human.py:409
            if (j >= q[0] && j < q[1] && i >= q[2] && i < q[3]) { j += q[4] - q[0]; i += q[5] - q[2]; }
        }
        if (!stage_source(i, j, r + 6 * s, a.W, a.H, (a.half_mask >> s) & 1, a.grid_dtype)) return -1;
    }
    return j * a.W + i;
}

// ---- general route: gather from global memory ---------------------------------------------------
template <typename T, int PX>
__global__ void __launch_bounds__(kRwThreads)
rewarp_fwd_kernel(const RewarpArgs a, T* __restrict__ out) {
    __shared__ float s_theta[kRwMaxViews][kRwMaxStages * 6];
    __shared__ int32_t s_paste[6];
    const int hw = a.H * a.W;
    const int bands = (hw + kRwThreads * PX - 1) / (kRwThreads * PX);
    const int cgroups = (a.C + a.cpc - 1) / a.cpc;
    int bid = blockIdx.x;
    const int band = bid % bands; bid /= bands;
    const int cg = bid % cgroups;
    const int b = bid / cgroups;
    if (threadIdx.x < a.views * a.stages * 6) {
        const int v = threadIdx.x / (a.stages * 6), k = threadIdx.x - v * a.stages * 6;
        s_theta[v][k] = a.view[v].theta[static_cast<int64_t>(b) * a.stages * 6 + k];
    }
    if (a.paste && threadIdx.x >= 32 && threadIdx.x < 38) s_paste[threadIdx.x - 32] = a.paste[6 * b + threadIdx.x - 32];
    __syncthreads();
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
                src[v][e] = through ? p0 + e
                                    : (p0 + e < hw ? composed_source(p0 + e, s_theta[v], a.paste ? s_paste : nullptr, a) : -1);
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

// ---- heatmap route: planes staged through padded shared memory -----------------------------------
constexpr int kRwPix = 16;               // pixels per thread and plane: planes up to 256*16 = 4096 px
constexpr int kRwMaxVec = 4;             // 16-byte staging copies per thread and plane
constexpr int kRwRing = 3;               // plane buffers per CTA: two planes in flight behind the one being gathered
constexpr int kWideRingHalf = 6;         // ... five for the 8 KB planes of the 2-byte types where few CTAs share an SM (rewarp_wide_kernel)
constexpr int kRwBufBytes = 32 * 1024;   // padded plane budget per buffer

// smallest row stride (32-bit words) >= words with stride % 32 == rem (rem % 4 == 0: 16-byte aligned rows)
__host__ __device__ inline int padded_stride(int words, int rem) {
    const int s = ((words - rem + 31) / 32) * 32 + rem;
    return s >= words ? s : s + 32;
}

// Jacobian of the composed output->source pixel map (evaluation order), from the rescaled thetas
__device__ __forceinline__ void composed_jacobian(const float* __restrict__ r, const RewarpArgs& a, float J[4]) {
    J[0] = 1.0f; J[1] = 0.0f; J[2] = 0.0f; J[3] = 1.0f;
    const float hx = 0.5f * a.W, hy = 0.5f * a.H;
    for (int s = 0; s < a.stages; ++s) {
        const float m0 = r[6 * s] * hx, m1 = r[6 * s + 1] * hx, m2 = r[6 * s + 3] * hy, m3 = r[6 * s + 4] * hy;
        const float n0 = m0 * J[0] + m1 * J[2], n1 = m0 * J[1] + m1 * J[3];
        const float n2 = m2 * J[0] + m3 * J[2], n3 = m2 * J[1] + m3 * J[3];
        J[0] = n0; J[1] = n1; J[2] = n2; J[3] = n3;
    }
}

// Row padding of the staged plane.  The 32 lanes of a gather read consecutive pixels, i.e. they walk
// the source plane in steps of (dcol, drow) elements; with a row stride of S words lane l lands in
// bank (l*dcol/EPW + S*l*drow) mod 32.  No single 16-byte-aligned stride is good for every direction
// (S = 4 mod 32 puts the whole warp into one bank when dcol = -4 drow), so every warp plays the walk
// for two candidates (S = 4 and S = 12 mod 32) and keeps the one with the lower conflict degree:
// average 2-way, worst 5-way over all rotations and scales, against 8-9-way for a fixed stride.
// Deterministic in (dcol, drow): all CTAs of a cluster agree.
__device__ __forceinline__ int pick_stride(int row_words, float dcol, float drow, int epw_log2) {
    const int lane = threadIdx.x & 31;
    const int px = __float2int_rn(static_cast<float>(lane << epw_log2) * dcol) >> epw_log2;
    const int py = __float2int_rn(static_cast<float>(lane << epw_log2) * drow);
    const int sa = padded_stride(row_words, 4), sb = padded_stride(row_words, 12);
    const int da = __reduce_max_sync(0xffffffffu, __popc(__match_any_sync(0xffffffffu, (py * sa + px) & 31)));
    const int db = __reduce_max_sync(0xffffffffu, __popc(__match_any_sync(0xffffffffu, (py * sb + px) & 31)));
    return da <= db ? sa : sb;
}

// asynchronous 16-byte global -> shared copies (LDGSTS, L2 only) and their group fences
__device__ __forceinline__ void cp_async16(uint32_t smem_byte_addr, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_byte_addr), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// padded byte offset of each of the thread's 16-byte vectors (-1: none), computed once per CTA
__device__ __forceinline__ void stage_offsets(int (&so)[kRwMaxVec], int nvec, int vpr, int stride) {
#pragma unroll
    for (int q = 0; q < kRwMaxVec; ++q) {
        const int v = q * kRwThreads + threadIdx.x;
        const int row = v / vpr;
        so[q] = v < nvec ? (row * stride + 4 * (v - row * vpr)) * 4 : -1;
    }
}
// one plane, global -> padded shared buffer: coalesced, no registers, completion through the group fence
template <typename T>
__device__ __forceinline__ void stage_issue(uint32_t buf_addr, const int (&so)[kRwMaxVec], const T* __restrict__ plane) {
    const uint4* src = reinterpret_cast<const uint4*>(plane) + threadIdx.x;
#pragma unroll
    for (int q = 0; q < kRwMaxVec; ++q)
        if (so[q] >= 0) cp_async16(buf_addr + so[q], src + q * kRwThreads);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// thread t owns the 32-bit words t + 256*slot of every plane (1 fp32 or 2 halves per word):
// consecutive lanes, consecutive pixels -> spread gathers, 128-byte coalesced stores
template <typename T> __device__ __forceinline__ void store_word(T* __restrict__ out, int word, const float* f);
template <> __device__ __forceinline__ void store_word<float>(float* __restrict__ out, int word, const float* f) {
    out[word] = f[0];
}
template <> __device__ __forceinline__ void store_word<__half>(__half* __restrict__ out, int word, const float* f) {
    reinterpret_cast<__half2*>(out)[word] = __floats2half2_rn(f[0], f[1]);
}
template <> __device__ __forceinline__ void store_word<__nv_bfloat16>(__nv_bfloat16* __restrict__ out, int word,
                                                                      const float* f) {
    reinterpret_cast<__nv_bfloat162*>(out)[word] = __floats2bfloat162_rn(f[0], f[1]);
}

__device__ __forceinline__ int ceil_log2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

// Every CTA of the cluster owns the slice [rank << slice_log2, (rank+1) << slice_log2) of a uint16 array
// that is laid out identically in all CTAs: pull the peers' slices through distributed shared memory
// with 128-bit loads (scalar 16-bit DSMEM loads cost ~2 us per 4096 on B200).  n % 8 == 0, slices >= 8.
__device__ __forceinline__ void pull_slices(cg::cluster_group& cluster, uint16_t* arr, int n, int slice_log2, int rank) {
    for (int vec = threadIdx.x; vec < n / 8; vec += kRwThreads) {
        const int p = vec * 8, owner = p >> slice_log2;
        if (owner != rank)
            reinterpret_cast<uint4*>(arr)[vec] = *reinterpret_cast<const uint4*>(cluster.map_shared_rank(arr + p, owner));
    }
}

// The composed map of a sample is shared by all of its channels.  The CTAs that split a sample's
// channels form a thread-block CLUSTER: each computes 1/n-th of the map into its own shared memory,
// and after one cluster barrier every CTA reads the offsets of its pixels from its peers through
// distributed shared memory — the index arithmetic is done once per sample, not once per CTA.
// dynamic smem: RING padded plane buffers | uint16 map[VIEWS][hw].  RING - 1 planes are in flight behind the
// one being gathered: 2 when a CTA owns few planes, 5 when it walks all channels of a sample (one CTA per sample
// from batch 32 up: 32 CTAs must keep enough bytes in flight on their own)
template <typename T, int VIEWS, int RING>
__global__ void __launch_bounds__(kRwThreads)
rewarp_smem_kernel(const RewarpArgs a, T* __restrict__ out, int buf_words) {
    constexpr int EPW = 4 / static_cast<int>(sizeof(T));  // elements per 32-bit word
    constexpr int SLOTS = kRwPix / EPW;                    // words per thread and plane
    extern __shared__ __align__(16) uint32_t rw_smem[];
    __shared__ float s_theta[VIEWS][kRwMaxStages * 6];
    cg::cluster_group cluster = cg::this_cluster();
    const int nrank = static_cast<int>(cluster.num_blocks()), rank = static_cast<int>(cluster.block_rank());
    const int hw = a.H * a.W, nwords = hw / EPW;
    const int b = blockIdx.x / nrank;                      // one cluster per sample
    const int c0 = min(a.C, rank * a.cpc), c1 = min(a.C, c0 + a.cpc);
    uint16_t* s_map = reinterpret_cast<uint16_t*>(rw_smem + RING * buf_words);
    if (threadIdx.x < VIEWS * a.stages * 6) {
        const int v = threadIdx.x / (a.stages * 6), k = threadIdx.x - v * a.stages * 6;
        s_theta[v][k] = a.view[v].theta[static_cast<int64_t>(b) * a.stages * 6 + k];
    }
    // the word behind every padded plane stays zero: out-of-bounds pixels gather from it (no select)
    const uint32_t zero_byte = static_cast<uint32_t>(buf_words - 4) * 4u;
    if (threadIdx.x < RING) rw_smem[threadIdx.x * buf_words + buf_words - 4] = 0u;
    __syncthreads();
    const int row_words = a.W / EPW, vpr = row_words / 4, nvec = nwords / 4;
    int stride[VIEWS], so[VIEWS][kRwMaxVec];
#pragma unroll
    for (int v = 0; v < VIEWS; ++v) {
        float J[4];
        composed_jacobian(s_theta[v], a, J);
        stride[v] = pick_stride(row_words, J[0], J[2], EPW - 1);  // one output column = (J[0], J[2]) in the source
        stage_offsets(so[v], nvec, vpr, stride[v]);
    }
    auto issue = [&](int it) {   // item = (channel, view)
        const int c = c0 + it / VIEWS, v = it - (it / VIEWS) * VIEWS;
        const T* plane = static_cast<const T*>(a.view[v].in) + (static_cast<int64_t>(b) * a.C + c) * hw;
        const uint32_t dst = smem_u32(rw_smem + (it % RING) * buf_words);
        if constexpr (VIEWS == 1) {
            stage_issue<T>(dst, so[0], plane);
        } else {
#pragma unroll
            for (int u = 0; u < VIEWS; ++u)
                if (u == v) stage_issue<T>(dst, so[u], plane);
        }
    };
    const int nitems = (c1 - c0) * VIEWS;
#pragma unroll
    for (int it = 0; it < RING - 1; ++it) {   // in flight while the map is built
        if (it < nitems) issue(it);
        cp_async_commit();
    }
    // this CTA's slice of the map: BYTE offsets into the padded plane, 16 bits each
    const int slice_log2 = ceil_log2((hw + nrank - 1) / nrank);
    {
        const int p_first = (rank << slice_log2) + threadIdx.x, p_end = min(hw, (rank + 1) << slice_log2);
        const int dj = kRwThreads / a.W, di = kRwThreads - dj * a.W;
#pragma unroll
        for (int v = 0; v < VIEWS; ++v) {
            StageRegs R;
            load_stages(R, s_theta[v], a.stages);
            int j0 = p_first / a.W, i0 = p_first - j0 * a.W;
            for (int p = p_first; p < p_end; p += kRwThreads) {
                int i = i0, j = j0;
                uint32_t o = zero_byte;
                if (composed_source_ij(i, j, R, a)) o = static_cast<uint32_t>(j * stride[v] * 4 + i * static_cast<int>(sizeof(T)));
                s_map[v * hw + p] = static_cast<uint16_t>(o);
                i0 += di; j0 += dj;
                if (i0 >= a.W) { i0 -= a.W; ++j0; }
            }
        }
    }
    cluster.sync();
#pragma unroll
    for (int v = 0; v < VIEWS; ++v) pull_slices(cluster, s_map + v * hw, hw, slice_log2, rank);
    cluster.sync();   // nobody leaves (or reuses its slice) while a peer is still reading it; also a CTA barrier
    // offsets of this thread's pixels, two per register
    uint32_t idx[VIEWS][kRwPix / 2];
#pragma unroll
    for (int v = 0; v < VIEWS; ++v) {
#pragma unroll
        for (int k = 0; k < kRwPix; ++k) {
            // pixel k of this thread: word (k / EPW) * 256 + t, element k % EPW of that word
            const int word = (k / EPW) * kRwThreads + threadIdx.x;
            const uint32_t o = word < nwords ? s_map[v * hw + word * EPW + (k % EPW)] : zero_byte;
            if (k & 1) idx[v][k >> 1] |= o << 16;
            else idx[v][k >> 1] = o;
        }
    }
    for (int c = c0; c < c1; ++c) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * a.C + c) * hw) + threadIdx.x;
        float acc[VIEWS == 1 ? 1 : kRwPix];
#pragma unroll
        for (int v = 0; v < VIEWS; ++v) {
            const int it = (c - c0) * VIEWS + v;
            cp_async_wait<RING - 2>();   // this thread's copies of item `it` have landed ...
            __syncthreads();                // ... everybody's have, and everybody is done with item it-1
            if (it + RING - 1 < nitems) issue(it + RING - 1);   // refills the buffer of item it-1
            cp_async_commit();
            const uint8_t* bytes = reinterpret_cast<const uint8_t*>(rw_smem + (it % RING) * buf_words);
            if constexpr (VIEWS == 1) {
                // single view: the values are moved, never converted
#pragma unroll
                for (int sl = 0; sl < SLOTS; ++sl) {
                    uint32_t w32;
                    if constexpr (EPW == 1) {
                        const uint32_t o = (idx[0][sl >> 1] >> (16 * (sl & 1))) & 0xffffu;
                        w32 = *reinterpret_cast<const uint32_t*>(bytes + o);
                    } else {
                        const uint32_t lo = *reinterpret_cast<const uint16_t*>(bytes + (idx[0][sl] & 0xffffu));
                        const uint32_t hi = *reinterpret_cast<const uint16_t*>(bytes + (idx[0][sl] >> 16));
                        w32 = lo | (hi << 16);
                    }
                    if (sl * kRwThreads + static_cast<int>(threadIdx.x) < nwords) o32[sl * kRwThreads] = w32;
                }
            } else {
#pragma unroll
                for (int k = 0; k < kRwPix; ++k) {
                    const uint32_t o = (idx[v][k >> 1] >> (16 * (k & 1))) & 0xffffu;
                    const float val = to_f32<T>(*reinterpret_cast<const T*>(bytes + o));
                    acc[k] = v == 0 ? val : acc[k] + val;  // torch.mean over the k views: sequential sum ...
                }
            }
        }
        if constexpr (VIEWS > 1) {
            const float nviews = static_cast<float>(VIEWS);
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl) {
                float f[EPW];
#pragma unroll
                for (int e = 0; e < EPW; ++e) f[e] = __fdiv_rn(acc[sl * EPW + e], nviews);  // ... one division (CPU mean)
                if (sl * kRwThreads + static_cast<int>(threadIdx.x) < nwords)
                    store_word<T>(reinterpret_cast<T*>(o32 - threadIdx.x), sl * kRwThreads + threadIdx.x, f);
            }
        }
    }
    cp_async_wait<0>();
}

// ---- backward ------------------------------------------------------------------------------------
// Inverts the composed map of one sample (map[p] = source pixel of output pixel p) in shared memory:
// off[s] = end of the list of source pixel s (start = off[s-1]),  lst[] = the output pixels of every
// list in ascending order, stored as `code(p)`.  Integer atomics only (16-bit counters packed two per
// word), so the result is independent of scheduling.  Call from all threads.
template <typename Code>
__device__ __forceinline__ void invert_map(const uint16_t* map, int hw, int s_lo, int s_hi, uint16_t* off, uint16_t* lst,
                                           uint32_t* s_scan, Code code) {
    // only the source pixels [s_lo, s_hi) are inverted (a cluster splits the range); off[] is indexed by
    // s - s_lo and must be zero on entry (barrier passed)
    const int ns = s_hi - s_lo;
    uint32_t* off32 = reinterpret_cast<uint32_t*>(off);
    // 1. per-source counts
    for (int p = threadIdx.x; p < hw; p += kRwThreads) {
        const int s = static_cast<int>(map[p]) - s_lo;
        if (s >= 0 && s < ns) atomicAdd(&off32[s >> 1], 1u << (16 * (s & 1)));
    }
    __syncthreads();
    // 2. exclusive scan of the counts (each thread owns a contiguous run)
    const int per = (ns + kRwThreads - 1) / kRwThreads;
    const int lo = min(ns, static_cast<int>(threadIdx.x) * per), hi = min(ns, lo + per);
    uint32_t run = 0;
    for (int s = lo; s < hi; ++s) run += off[s];
    s_scan[threadIdx.x] = run;
    __syncthreads();
    if (threadIdx.x < 32) {
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
    for (int s = lo; s < hi; ++s) { const uint32_t c = off[s]; off[s] = static_cast<uint16_t>(run); run += c; }
    __syncthreads();
    // 3. fill: bump the 16-bit cursor of list s (slot order is arbitrary here ...)
    for (int p = threadIdx.x; p < hw; p += kRwThreads) {
        const int s = static_cast<int>(map[p]) - s_lo;
        if (s >= 0 && s < ns) {
            const uint32_t old = atomicAdd(&off32[s >> 1], 1u << (16 * (s & 1)));
            lst[(old >> (16 * (s & 1))) & 0xffffu] = code(p);
        }
    }
    __syncthreads();
    // 4. ... so sort every list (code() is monotone in p); lists hold ~1/scale^2 entries
    for (int s = threadIdx.x; s < ns; s += kRwThreads) {
        const int st = s == 0 ? 0 : off[s - 1], en = off[s];
        for (int q = st + 1; q < en; ++q) {
            const uint16_t v = lst[q];
            int k = q - 1;
            while (k >= st && lst[k] > v) { lst[k + 1] = lst[k]; --k; }
            lst[k + 1] = v;
        }
    }
    __syncthreads();
}

// composed map (source pixel of every output pixel, 0xffff = none) of pixels [p0, p1)
__device__ __forceinline__ void build_map_slice(const float* __restrict__ theta, const RewarpArgs& a, uint16_t* map,
                                                int p0, int p1) {
    StageRegs R;
    load_stages(R, theta, a.stages);
    const int dj = kRwThreads / a.W, di = kRwThreads - dj * a.W;
    int p = p0 + threadIdx.x;
    int j0 = p / a.W, i0 = p - j0 * a.W;
    for (; p < p1; p += kRwThreads) {
        int i = i0, j = j0;
        map[p] = composed_source_ij(i, j, R, a) ? static_cast<uint16_t>(j * a.W + i) : 0xffffu;
        i0 += di; j0 += dj;
        if (i0 >= a.W) { i0 -= a.W; ++j0; }
    }
}

// The composed map of one sample built and inverted by the whole cluster: every CTA computes its slice
// of the map, pulls the peers' slices (DSMEM), inverts the map for ITS slice of the source pixels,
// and the shares are then gathered so that every CTA ends up with the complete result:
//   off[s]  = end of the list of source pixel s (start = off[s-1]) in lst[]
//   lst[]   = the output pixels of every list, ascending, as padded byte offsets (row stride `stride`
//             words, ES-byte elements).  lst aliases map.
// Preconditions: off[] zeroed over (hw + 8) / 2 words, theta in shared memory, a CTA barrier passed.
template <size_t ES>
__device__ __forceinline__ void cluster_invert(cg::cluster_group& cluster, const float* __restrict__ s_theta,
                                               const RewarpArgs& a, int stride, uint16_t* off, uint16_t* lst_loc,
                                               uint16_t* map, uint32_t* s_scan) {
    __shared__ uint32_t s_total, s_base[9], s_vbase[9];
    const int nrank = static_cast<int>(cluster.num_blocks()), rank = static_cast<int>(cluster.block_rank());
    const int hw = a.H * a.W;
    uint16_t* lst = map;
    // 1. the composed map: own slice, then the peers' through DSMEM
    const int slice_log2 = ceil_log2((hw + nrank - 1) / nrank);
    const int lo = min(hw, rank << slice_log2), hi = min(hw, (rank + 1) << slice_log2);
    build_map_slice(s_theta, a, map, lo, hi);
    cluster.sync();
    pull_slices(cluster, map, hw, slice_log2, rank);
    __syncthreads();
    // 2. every CTA inverts the map for ITS slice of the source pixels only (list ends relative to the share)
    {
        const int W = a.W;
        invert_map(map, hw, lo, hi, off + lo, lst_loc, s_scan, [=](int p) {
            const int row = p / W;
            return static_cast<uint16_t>(row * stride * 4 + (p - row * W) * static_cast<int>(ES));
        });
    }
    if (threadIdx.x == 0) s_total = hi > lo ? off[hi - 1] : 0u;
    cluster.sync();   // every share is complete (and nobody reads a peer's map any more: it becomes `lst`)
    // 3. gather the shares: the lists of rank r follow those of the ranks below it
    if (threadIdx.x == 0) {
        uint32_t run = 0, vrun = 0;
        for (int r = 0; r < 8; ++r) {
            const uint32_t tot = r < nrank ? *cluster.map_shared_rank(&s_total, r) : 0u;
            s_base[r] = run; s_vbase[r] = vrun;
            run += tot; vrun += (tot + 7) / 8;
        }
        s_base[8] = run; s_vbase[8] = vrun;
    }
    pull_slices(cluster, off, hw, slice_log2, rank);
    __syncthreads();
    const int nvecs = static_cast<int>(s_vbase[8]);   // 8-entry vectors of every share, appended at the share's base
    for (int g = threadIdx.x; g < nvecs; g += kRwThreads) {
        int r = 0;
#pragma unroll
        for (int t = 1; t < 8; ++t) r += static_cast<uint32_t>(g) >= s_vbase[t] ? 1 : 0;
        const uint32_t bs = s_base[r], tot = s_base[r + 1] - bs;
        const int q0 = (g - static_cast<int>(s_vbase[r])) * 8;
        const uint4 v4 = *reinterpret_cast<const uint4*>(cluster.map_shared_rank(lst_loc + q0, r));
        const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (static_cast<uint32_t>(q0 + e) < tot) lst[bs + q0 + e] = static_cast<uint16_t>(w[e >> 1] >> (16 * (e & 1)));
    }
    cluster.sync();   // the peers have pulled this CTA's list ends and lists: both may change / go away now
    for (int s = threadIdx.x; s < hw; s += kRwThreads)   // list ends become absolute
        off[s] = static_cast<uint16_t>(off[s] + s_base[s >> slice_log2]);
    __syncthreads();
}

// the first four contributors of source pixel s as padded byte offsets (missing ones -> the zero word)
__device__ __forceinline__ uint2 list_slots(const uint16_t* off, const uint16_t* lst, int s, uint32_t zero_byte, bool& overflow) {
    const int st = s == 0 ? 0 : off[s - 1], en = off[s];
    uint32_t o[4] = {zero_byte, zero_byte, zero_byte, zero_byte};
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (st + q < en) o[q] = lst[st + q];
    overflow = en - st > 4;
    return make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
}

// Source pixels with more than four contributors (zoom factors above ~1.7; a few per cent of a plane at
// most): the thread redoes exactly those of its pixels (bit k of `mask`), all entries in ascending
// order, and overwrites the element it has just stored.
template <typename T>
__device__ __forceinline__ void redo_long_lists(uint32_t mask, const uint16_t* __restrict__ off, const uint16_t* __restrict__ lst,
                                                const uint8_t* bytes, T* __restrict__ o) {
    constexpr int EPW = 4 / static_cast<int>(sizeof(T));
    while (mask) {
        const int k = __ffs(mask) - 1;
        mask &= mask - 1;
        const int s = ((k / EPW) * kRwThreads + threadIdx.x) * EPW + (k % EPW);
        const int st = s == 0 ? 0 : off[s - 1], en = off[s];
        float sum = 0.0f;
        for (int q = st; q < en; ++q) sum += to_f32<T>(*reinterpret_cast<const T*>(bytes + lst[q]));
        o[s] = from_f32<T>(sum);
    }
}

// heatmap route (planes up to 4096 px): gradient planes staged through padded shared memory; the
// CTAs of a sample form a cluster and build the composed map together (see rewarp_smem_kernel).
// dynamic smem: kRwRing plane buffers | uint16 off[hw + 8] | lst_loc[hw] | map[hw] (-> lst)
template <typename T>
__global__ void __launch_bounds__(kRwThreads, 2)
rewarp_bwd_smem_kernel(const RewarpArgs a, const T* __restrict__ gout, T* __restrict__ gin, int buf_words) {
    constexpr int EPW = 4 / static_cast<int>(sizeof(T));
    constexpr int SLOTS = kRwPix / EPW;
    extern __shared__ __align__(16) uint32_t rw_smem[];
    __shared__ float s_theta[kRwMaxStages * 6];
    __shared__ uint32_t s_scan[kRwThreads];
    cg::cluster_group cluster = cg::this_cluster();
    const int nrank = static_cast<int>(cluster.num_blocks()), rank = static_cast<int>(cluster.block_rank());
    const int hw = a.H * a.W, nwords = hw / EPW;
    // off      : list ends, indexed by source pixel; own slice inverted here, the peers' slices pulled
    // lst_loc  : this CTA's share of the lists (read by the peers)
    // map->lst : the composed map, later the complete list array
    uint16_t* off = reinterpret_cast<uint16_t*>(rw_smem + kRwRing * buf_words);
    uint16_t* lst_loc = off + ((hw + 8 + 7) & ~7);
    uint16_t* map = lst_loc + ((hw + 7) & ~7);
    uint16_t* lst = map;
    const int b = blockIdx.x / nrank;
    const int c0 = min(a.C, rank * a.cpc), c1 = min(a.C, c0 + a.cpc);
    if (threadIdx.x < a.stages * 6)
        s_theta[threadIdx.x] = a.view[0].theta[static_cast<int64_t>(b) * a.stages * 6 + threadIdx.x];
    uint32_t* off32 = reinterpret_cast<uint32_t*>(off);
    for (int s = threadIdx.x; s < (hw + 8) / 2; s += kRwThreads) off32[s] = 0u;
    if (threadIdx.x < kRwRing) rw_smem[threadIdx.x * buf_words + buf_words - 4] = 0u;   // the zero word of every buffer
    __syncthreads();
    const int row_words = a.W / EPW, vpr = row_words / 4, nvec = nwords / 4;
    float J[4];
    composed_jacobian(s_theta, a, J);
    // consecutive source pixels pull from output pixels one step of the INVERSE map apart:
    // J^-1 (1,0) = (J[3], -J[2]) / det
    const float det = J[0] * J[3] - J[1] * J[2];
    const float inv = det != 0.0f ? 1.0f / det : 0.0f;
    const int stride = pick_stride(row_words, J[3] * inv, -J[2] * inv, EPW - 1);
    int so[kRwMaxVec];
    stage_offsets(so, nvec, vpr, stride);
    auto issue = [&](int it) {
        stage_issue<T>(smem_u32(rw_smem + (it % kRwRing) * buf_words), so, gout + (static_cast<int64_t>(b) * a.C + c0 + it) * hw);
    };
    const int nitems = c1 - c0;
#pragma unroll
    for (int it = 0; it < kRwRing - 1; ++it) {   // in flight during the inversion
        if (it < nitems) issue(it);
        cp_async_commit();
    }
    // the composed map built and inverted by the cluster together: off[] = absolute list ends, lst[] = lists
    cluster_invert<sizeof(T)>(cluster, s_theta, a, stride, off, lst_loc, map, s_scan);
    if (nitems > 0) {
        // The first four contributors of each source pixel this thread owns, as padded byte offsets in
        // registers (missing ones point at the zero word): the per-plane sum is four independent LDS
        // and three adds per pixel, no data-dependent loop.  Lists longer than four entries (zoom
        // factors above ~1.7) take the loop for the rest; the order stays ascending p either way.
        const uint32_t zero_byte = static_cast<uint32_t>(buf_words - 4) * 4u;
        uint2 slot[kRwPix];
        uint32_t long_mask = 0;
#pragma unroll
        for (int k = 0; k < kRwPix; ++k) {
            const int word = (k / EPW) * kRwThreads + threadIdx.x;
            bool overflow = false;
            slot[k] = word < nwords ? list_slots(off, lst, word * EPW + (k % EPW), zero_byte, overflow)
                                    : make_uint2(zero_byte | (zero_byte << 16), zero_byte | (zero_byte << 16));
            if (overflow) long_mask |= 1u << k;
        }
        for (int it = 0; it < nitems; ++it) {
            cp_async_wait<kRwRing - 2>();
            __syncthreads();
            if (it + kRwRing - 1 < nitems) issue(it + kRwRing - 1);
            cp_async_commit();
            const uint8_t* bytes = reinterpret_cast<const uint8_t*>(rw_smem + (it % kRwRing) * buf_words);
            T* o = gin + (static_cast<int64_t>(b) * a.C + c0 + it) * hw;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl) {
                float f[EPW];
#pragma unroll
                for (int e = 0; e < EPW; ++e) {
                    const uint2 sq = slot[sl * EPW + e];
                    const float v0 = to_f32<T>(*reinterpret_cast<const T*>(bytes + (sq.x & 0xffffu)));
                    const float v1 = to_f32<T>(*reinterpret_cast<const T*>(bytes + (sq.x >> 16)));
                    const float v2 = to_f32<T>(*reinterpret_cast<const T*>(bytes + (sq.y & 0xffffu)));
                    const float v3 = to_f32<T>(*reinterpret_cast<const T*>(bytes + (sq.y >> 16)));
                    f[e] = ((v0 + v1) + v2) + v3;   // ascending p, fp32, one rounding to T
                }
                const int word = sl * kRwThreads + threadIdx.x;
                if (word < nwords) store_word<T>(o, word, f);
            }
            redo_long_lists<T>(long_mask, off, lst, bytes, o);
        }
    }
    cp_async_wait<0>();
}

// ---- wide forward route: one CTA of kWideThreads per sample ------------------------------------------------
// With 256 threads a sample's CTA spends three quarters of its instructions on the composed map (16 pixels per
// thread, ~250 issue slots each on a half grid: four fp16 roundings, two int->float and two float->int
// conversions per stage and pixel, all quarter-rate) and only then starts moving data; 256 such CTAs leave the
// SMs at 8 warps each (IPC 1.3, 42-47 % of the HBM roofline for fp16 planes, profiles/r02f_rewarp_*).  Here a
// sample gets 512 threads (8 pixels each): no cluster, no DSMEM exchange, two CTAs share an SM (one in its
// arithmetic phase while the other streams planes), and 256 samples fit one wave.  Measured (r02k, C5 / C2):
// fp16 26.3 / 12.2 us against 28.4 / 14.0 us for the 256-thread kernel, fp32 33.7 us (80 % of the HBM roofline)
// against 35.0.  The same idea for the backward (map + inversion + ordered sums in one 512-thread CTA, no plan)
// was slower than the plan route (C5 91-119 us against 66 us; C2 56-69 us against 15 us): the inversion is a
// chain of dependent shared-memory phases per sample and gains nothing from more threads — dropped.
constexpr int kWideThreads = 512;
constexpr int kWidePix = 8;                // pixels per thread: planes up to 512 * 8 = 4096 px

// The composed map of one sample into shared memory, one pixel per loop trip and NOT unrolled: inlining the
// per-pixel arithmetic (four stage bodies x four variants) once per pixel of a thread made the wide kernels 17 000
// SASS instructions long, and they stalled on instruction fetch (ncu "no_instruction", profiles/r02j_*).
// CODE(i, j) -> uint16: what is stored for a pixel whose source is (i, j); `none` for pixels that leave the image.
template <int HM, int GD, typename Code>
__device__ __noinline__ void build_map_variant(const float* __restrict__ s_theta, int W, int H, int stages, int half_mask,
                                               int grid_dtype, uint16_t* __restrict__ map, uint16_t none, Code code) {
    const int hw = H * W;
    const int nt = blockDim.x;
    const int dj = nt / W, di = nt - dj * W;
    const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
    const float cx = 0.5f - 0.5f * Wf, cy = 0.5f - 0.5f * Hf;
    const bool xy_exact = GD == UDAPE_F16 ? true : grid_xy_exact(W, H, grid_dtype);
    int p = threadIdx.x;
    int j0 = p / W, i0 = p - j0 * W;
#pragma unroll 1
    for (; p < hw; p += nt) {
        float fi = static_cast<float>(i0), fj = static_cast<float>(j0);
        bool ok = true;
#pragma unroll 1
        for (int st = 0; st < stages && ok; ++st)
            ok = stage_source_f<HM, GD>(fi, fj, s_theta + 6 * st, Wf, Hf, cx, cy, (half_mask >> st) & 1, grid_dtype, xy_exact);
        map[p] = ok ? code(static_cast<int>(fi), static_cast<int>(fj)) : none;
        i0 += di; j0 += dj;
        if (i0 >= W) { i0 -= W; ++j0; }
    }
}
template <typename Code>
__device__ __forceinline__ void build_map_compact(const float* __restrict__ s_theta, const RewarpArgs& a, uint16_t* __restrict__ map,
                                                  uint16_t none, Code code) {
    if (a.half_mask == 0) build_map_variant<0, -1>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
    else if (a.half_mask == (1 << a.stages) - 1) {
        if (a.grid_dtype == UDAPE_F16) build_map_variant<1, UDAPE_F16>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
        else build_map_variant<1, UDAPE_BF16>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
    } else build_map_variant<2, -1>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
}

// the same builder with the stage table in registers (no LDS.64 of theta per stage and pixel)
template <int HM, int GD, typename Code>
__device__ __noinline__ void build_map_regs(const float* __restrict__ s_theta, int W, int H, int stages, int half_mask,
                                            int grid_dtype, uint16_t* __restrict__ map, uint16_t none, Code code) {
    const int hw = H * W;
    const int nt = blockDim.x;
    const int dj = nt / W, di = nt - dj * W;
    const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
    const float cx = 0.5f - 0.5f * Wf, cy = 0.5f - 0.5f * Hf;
    const bool xy_exact = GD == UDAPE_F16 ? true : grid_xy_exact(W, H, grid_dtype);
    float r[kRwMaxStages][6];
#pragma unroll
    for (int st = 0; st < kRwMaxStages; ++st)
#pragma unroll
        for (int k = 0; k < 6; ++k) r[st][k] = st < stages ? s_theta[6 * st + k] : 0.0f;
    int p = threadIdx.x;
    int j0 = p / W, i0 = p - j0 * W;
#pragma unroll 1
    for (; p < hw; p += nt) {
        float fi = static_cast<float>(i0), fj = static_cast<float>(j0);
        bool ok = true;
#pragma unroll
        for (int st = 0; st < kRwMaxStages; ++st)
            if (st < stages && ok)
                ok = stage_source_f<HM, GD>(fi, fj, r[st], Wf, Hf, cx, cy, (half_mask >> st) & 1, grid_dtype, xy_exact);
        map[p] = ok ? code(static_cast<int>(fi), static_cast<int>(fj)) : none;
        i0 += di; j0 += dj;
        if (i0 >= W) { i0 -= W; ++j0; }
    }
}
template <typename Code>
__device__ __forceinline__ void build_map_fast(const float* __restrict__ s_theta, const RewarpArgs& a, uint16_t* __restrict__ map,
                                               uint16_t none, Code code) {
    if (a.half_mask == 0) build_map_regs<0, -1>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
    else if (a.half_mask == (1 << a.stages) - 1) {
        if (a.grid_dtype == UDAPE_F16) build_map_regs<1, UDAPE_F16>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
        else build_map_regs<1, UDAPE_BF16>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
    } else build_map_regs<2, -1>(s_theta, a.W, a.H, a.stages, a.half_mask, a.grid_dtype, map, none, code);
}

// ---- the push plan: what the backward needs, computed once per batch and kept in global memory ----------------
// The backward is a SCATTER  grad_in[src(p)] += grad_out[p].  It is made deterministic — and free of atomics,
// sorting and per-pixel branches — by giving every output pixel p its own SLOT in shared memory:
//   * the first contributor of a source pixel (rank 0, the smallest p) is stored straight into the padded fp32
//     accumulator plane at the source pixel;
//   * every later contributor (rank >= 1) is stored into a TAIL array ordered by (source pixel, rank);
//   * pixels without a source carry a marker offset and are not stored at all.
// One barrier later the thread that owns a GROUP (a source pixel with two or more contributors) folds the group's
// tail entries — contiguous, ascending p — into the accumulator:  ((v0 + v1) + v2) + ...  the same fp32 order, bit
// for bit, as the list-based kernels below and as a sequential CPU scatter.  Source pixels without contributors
// keep the zero the previous read-out left.
//   per sample, uint16 units:  hdr[8] = {row stride of the accumulator plane in words, groups, tail entries,
//                                        groups with a single tail entry, 0 ...}
//                              | uint16 code[hw8]        byte offset of pixel p's slot in a plane's shared memory
//                              | uint2  group[hw8/2]     {source-pixel byte offset | first tail slot << 16, tail entries};
//                                                        the groups with ONE tail entry first (their warps run no
//                                                        inner loop), then the longer ones, each part in source order
// The ranks come from max-key rounds in shared memory (integer atomicMax, scheduling-independent): in round r every
// pending pixel offers key = (r+1) << 16 | (0xffff - p) to owner[src]; the largest key of the round is the smallest
// pending p, it takes rank r and retires.  Keys of earlier rounds are smaller than any key of round r, so owner[]
// is never cleared, and after the last round owner[src] >> 16 is the contributor count of src.
// (Measured on B200 with the reference's augmentation range — rotation 180, scale .6-1.3, shear 30: a sample has
// ~1400 tail entries and lists of up to 16; rank-by-rank scatter rounds straight from registers, the first
// version of this kernel, spent 20 instructions per pixel on divergent per-rank branches: 93-135 us at C5.)
__host__ __device__ inline int64_t plan_elems_for(int64_t hw) { const int64_t hw8 = (hw + 7) & ~7ll; return 8 + 3 * hw8; }
// shared memory of one plane in the backward: accumulator plane | tail (one slot per pixel) | dummy word (16 bytes)
__host__ __device__ inline int push_pitch_bytes(int acc_words, int hw) { return acc_words * 4 + ((hw + 7) & ~7) * 4 + 16; }

constexpr int kPlanThreads = 512;
constexpr int kPlanPix = 8;   // pixels per thread: planes up to 4096 px

// dynamic smem: uint32 owner[acc_words] | uint16 map[hw8]
template <int ES>
__global__ void __launch_bounds__(kPlanThreads, 2)
rewarp_push_plan_kernel(const RewarpArgs a, uint16_t* __restrict__ plan, int acc_words) {
    constexpr int EPW = 4 / ES;
    extern __shared__ __align__(16) uint32_t rw_smem[];
    __shared__ float s_theta[kRwMaxStages * 6];
    __shared__ uint32_t s_warp[kPlanThreads / 32], s_warp_t[kPlanThreads / 32];
    const int hw = a.H * a.W, hw8 = (hw + 7) & ~7;
    const int b = blockIdx.x;
    uint32_t* owner = rw_smem;
    uint16_t* map = reinterpret_cast<uint16_t*>(rw_smem + acc_words);
    if (threadIdx.x < a.stages * 6)
        s_theta[threadIdx.x] = a.view[0].theta[static_cast<int64_t>(b) * a.stages * 6 + threadIdx.x];
    for (int i = threadIdx.x; i < acc_words / 4; i += kPlanThreads) reinterpret_cast<uint4*>(owner)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    // the backward's lanes own consecutive 32-bit words of the gradient plane (EPW pixels): one lane step moves
    // the source by EPW columns of the composed Jacobian
    float J[4];
    composed_jacobian(s_theta, a, J);
    const int stride = pick_stride(a.W, static_cast<float>(EPW) * J[0], static_cast<float>(EPW) * J[2], 0);
    build_map_fast(s_theta, a, map, static_cast<uint16_t>(0xffffu), [=](int i, int j) {
        return static_cast<uint16_t>((j * stride + i) * 4);
    });
    __syncthreads();
    // 1. ranks
    uint32_t code[kPlanPix];   // source byte offset | rank << 16 ; 0xffffffff: no source
    uint32_t pending = 0;
#pragma unroll
    for (int k = 0; k < kPlanPix; ++k) {
        const int p = k * kPlanThreads + threadIdx.x;
        code[k] = p < hw ? map[p] : 0xffffu;
        if (code[k] != 0xffffu) pending |= 1u << k;
        else code[k] = 0xffffffffu;
    }
    for (uint32_t r = 0;; ++r) {   // (after two or three rounds most threads have nothing pending and only meet the barriers)
        const uint32_t key_hi = (r + 1) << 16;
        if (pending) {
#pragma unroll
            for (int k = 0; k < kPlanPix; ++k)
                if (pending & (1u << k)) atomicMax(&owner[code[k] >> 2], key_hi | (0xffffu - static_cast<uint32_t>(k * kPlanThreads + threadIdx.x)));
        }
        __syncthreads();
        if (pending) {
#pragma unroll
            for (int k = 0; k < kPlanPix; ++k)
                if ((pending & (1u << k)) && owner[code[k] >> 2] == (key_hi | (0xffffu - static_cast<uint32_t>(k * kPlanThreads + threadIdx.x)))) {
                    code[k] |= r << 16;
                    pending &= ~(1u << k);
                }
        }
        if (!__syncthreads_or(pending != 0)) break;   // also: this round's reads are over before the next round's atomics
    }
    // 2. groups (source pixels with >= 2 contributors): exclusive scans of (single-tail groups | longer groups << 16)
    //    and of the tail entries over the accumulator words, a contiguous run per thread
    const int per = (acc_words + kPlanThreads - 1) / kPlanThreads;
    const int lo = min(acc_words, static_cast<int>(threadIdx.x) * per), hi = min(acc_words, lo + per);
    uint32_t run_g = 0, run_t = 0;
    for (int w = lo; w < hi; ++w) {
        const uint32_t cnt = owner[w] >> 16;
        if (cnt >= 2u) { run_g += cnt == 2u ? 1u : (1u << 16); run_t += cnt - 1u; }
    }
    uint32_t incl_g = run_g, incl_t = run_t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t tg = __shfl_up_sync(0xffffffffu, incl_g, o), tt = __shfl_up_sync(0xffffffffu, incl_t, o);
        if (static_cast<int>(threadIdx.x & 31) >= o) { incl_g += tg; incl_t += tt; }
    }
    if ((threadIdx.x & 31) == 31) { s_warp[threadIdx.x >> 5] = incl_g; s_warp_t[threadIdx.x >> 5] = incl_t; }
    __syncthreads();
    uint32_t base_g = 0, base_t = 0, total_g = 0, total_t = 0;
#pragma unroll
    for (int q = 0; q < kPlanThreads / 32; ++q) {
        const uint32_t tg = s_warp[q], tt = s_warp_t[q];
        if (q < static_cast<int>(threadIdx.x >> 5)) { base_g += tg; base_t += tt; }
        total_g += tg; total_t += tt;
    }
    const uint32_t nsingle = total_g & 0xffffu, ngroups = nsingle + (total_g >> 16);
    uint16_t* P = plan + static_cast<int64_t>(b) * plan_elems_for(hw);
    uint16_t* g_code = P + 8;
    uint2* g_group = reinterpret_cast<uint2*>(P + 8 + hw8);
    run_g = base_g + incl_g - run_g;   // exclusive
    run_t = base_t + incl_t - run_t;
    for (int w = lo; w < hi; ++w) {
        const uint32_t cnt = owner[w] >> 16;
        if (cnt >= 2u) {
            const uint32_t gi = cnt == 2u ? (run_g & 0xffffu) : nsingle + (run_g >> 16);
            g_group[gi] = make_uint2(static_cast<uint32_t>(w * 4) | (run_t << 16), cnt - 1u);
            owner[w] = run_t;        // read back by the group's contributors below
            run_g += cnt == 2u ? 1u : (1u << 16);
            run_t += cnt - 1u;
        }
    }
    __syncthreads();
    // 3. slots
    const uint32_t tail0 = static_cast<uint32_t>(acc_words) * 4u, dummy = tail0 + static_cast<uint32_t>(hw8) * 4u;
#pragma unroll
    for (int k = 0; k < kPlanPix; ++k) {
        const int p = k * kPlanThreads + threadIdx.x;
        uint32_t slot = dummy;
        if (code[k] != 0xffffffffu) {
            const uint32_t src = code[k] & 0xffffu, rk = code[k] >> 16;
            slot = rk == 0u ? src : tail0 + (owner[src >> 2] + rk - 1u) * 4u;
        }
        if (p < hw8) g_code[p] = static_cast<uint16_t>(slot);
    }
    if (threadIdx.x < 8) {
        const uint32_t t = threadIdx.x;
        P[t] = static_cast<uint16_t>(t == 0 ? static_cast<uint32_t>(stride) : (t == 1 ? ngroups : (t == 2 ? total_t : (t == 3 ? nsingle : 0u))));
    }
}

template <typename T> __device__ __forceinline__ float word_elem(uint32_t w, int e);
template <> __device__ __forceinline__ float word_elem<float>(uint32_t w, int) { return __uint_as_float(w); }
template <> __device__ __forceinline__ float word_elem<__half>(uint32_t w, int e) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(e ? (w >> 16) : (w & 0xffffu))));
}
template <> __device__ __forceinline__ float word_elem<__nv_bfloat16>(uint32_t w, int e) {
    return __uint_as_float(e ? (w & 0xffff0000u) : (w << 16));
}
// four accumulators -> four elements of T, stored at vector index vi of the plane; "+ 0" turns a lone -0.0 into
// +0.0 like the zero-initialised scatter target of the reference (done on the packed pairs for the 2-byte types)
template <typename T> __device__ __forceinline__ void store_quad(T* __restrict__ plane, int vi, float4 x);
template <> __device__ __forceinline__ void store_quad<float>(float* __restrict__ plane, int vi, float4 x) {
    x.x = __fadd_rn(x.x, 0.0f); x.y = __fadd_rn(x.y, 0.0f); x.z = __fadd_rn(x.z, 0.0f); x.w = __fadd_rn(x.w, 0.0f);
    reinterpret_cast<float4*>(plane)[vi] = x;
}
template <> __device__ __forceinline__ void store_quad<__half>(__half* __restrict__ plane, int vi, float4 x) {
    const __half2 z = __floats2half2_rn(0.0f, 0.0f);
    const __half2 lo = __hadd2(__floats2half2_rn(x.x, x.y), z), hi = __hadd2(__floats2half2_rn(x.z, x.w), z);
    reinterpret_cast<uint2*>(plane)[vi] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}
template <> __device__ __forceinline__ void store_quad<__nv_bfloat16>(__nv_bfloat16* __restrict__ plane, int vi, float4 x) {
    const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
    const __nv_bfloat162 lo = __hadd2(__floats2bfloat162_rn(x.x, x.y), z), hi = __hadd2(__floats2bfloat162_rn(x.z, x.w), z);
    reinterpret_cast<uint2*>(plane)[vi] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

// The passes of a launch, sample-major (ppc passes of NP planes per sample, the last one possibly shorter); a CTA
// takes a contiguous run of equal length.  (Measured and dropped, profiles/r02ak: every sample's full passes first and
// the short ones after, in runs of equal cost (fixed + planes) — a pass costs nearly the same whatever its plane
// count: 25.4 - 46.8 us against 24.7 at C5.  Requesting the next sample's plan words one pass ahead changed
// nothing either (24.5 us): what a pass costs is its ~3 k cycles of shared-memory wavefronts — scattered pushes,
// 128-bit read-out — and its three barriers, not the plan's latency.)
struct PassOrder {
    int ppc, np;
    long long n_all;
    __device__ __forceinline__ PassOrder(int B, int C, int NP) : ppc((C + NP - 1) / NP), np(NP), n_all(static_cast<long long>(B) * ((C + NP - 1) / NP)) {}
    __device__ __forceinline__ int first(unsigned k, unsigned grid) const { return static_cast<int>(n_all * k / grid); }
    __device__ __forceinline__ void at(int v, int& b, int& c) const { b = v / ppc; c = (v - b * ppc) * np; }
};

// backward from the push plan.  The work is cut into PASSES of NP gradient planes of one sample; the grid is one
// wave of CTAs and CTA i owns a contiguous, equal share of the passes (a sample's passes may be split between
// CTAs; a CTA reloads its per-sample state — 8 registers of slots — when its run crosses into the next sample).
// Per pass: the planes come straight from global memory into registers (coalesced 32-bit words; the next pass's
// loads are issued as soon as the registers are free), every pixel is pushed into its slot, the groups are
// folded, and the accumulators are read out with 128-bit loads (re-zeroed in the same instruction pair) and
// stored coalesced.  dynamic smem: NP x (float acc[acc_words] | float tail[hw8] | 16 bytes)
template <typename T, int NP>
__global__ void __launch_bounds__(kRwThreads, 3)
rewarp_bwd_push_kernel(const RewarpArgs a, const T* __restrict__ gout, T* __restrict__ gin, int acc_words,
                       const uint16_t* __restrict__ plan) {
    constexpr int EPW = 4 / static_cast<int>(sizeof(T));
    constexpr int SLOTS = kRwPix / EPW;
    constexpr int QUADS = kRwPix / 4;
    extern __shared__ __align__(16) uint32_t rw_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(rw_smem);
    const int hw = a.H * a.W, hw8 = (hw + 7) & ~7, nwords = hw / EPW, nquads = hw / 4;
    const PassOrder order(a.B, a.C, NP);
    const int pass0 = order.first(blockIdx.x, gridDim.x), pass1 = order.first(blockIdx.x + 1, gridDim.x);
    if (pass0 >= pass1) return;
    const int pitch = push_pitch_bytes(acc_words, hw);
    const uint32_t tail0 = static_cast<uint32_t>(acc_words) * 4u, dummy = tail0 + static_cast<uint32_t>(hw8) * 4u;
    // the accumulator quads this thread reads out: row and 4 * column of quad t + 256 * v
    int ro_row[QUADS], ro_col[QUADS];
#pragma unroll
    for (int v = 0; v < QUADS; ++v) {
        const int vi = v * kRwThreads + threadIdx.x;
        ro_row[v] = vi < nquads ? (vi * 4) / a.W : -1;
        ro_col[v] = (vi * 4 - ro_row[v] * a.W) * 4;
    }
#pragma unroll
    for (int n = 0; n < NP; ++n)
        for (int i = threadIdx.x; i < acc_words / 4; i += kRwThreads) reinterpret_cast<uint4*>(smem + n * pitch)[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t* g32 = reinterpret_cast<const uint32_t*>(gout);
    uint32_t g[NP][SLOTS];
    auto load = [&](int pass) {
        int b, c;
        order.at(pass, b, c);
#pragma unroll
        for (int n = 0; n < NP; ++n) {
            const uint32_t* src = g32 + (static_cast<int64_t>(b) * a.C + c + n) * nwords + threadIdx.x;
#pragma unroll
            for (int sl = 0; sl < SLOTS; ++sl)
                g[n][sl] = (c + n < a.C && sl * kRwThreads + static_cast<int>(threadIdx.x) < nwords) ? ldg_stream_u32(src + sl * kRwThreads) : 0u;
        }
    };
    load(pass0);
    uint32_t cw[kRwPix / 2];   // this thread's pixels: word t + 256*slot of every plane, EPW pixels per word; two 16-bit slots per register
    int cur_b = -1, stride4 = 0, ngroups = 0;
    const uint2* g_group = nullptr;
    __syncthreads();
    for (int pass = pass0; pass < pass1; ++pass) {
        int b, c;
        order.at(pass, b, c);
        const int np = min(NP, a.C - c);
        if (b != cur_b) {   // CTA-uniform: the per-sample state
            cur_b = b;
            const uint16_t* P = plan + static_cast<int64_t>(b) * plan_elems_for(hw);
            const uint16_t* g_code = P + 8;
            g_group = reinterpret_cast<const uint2*>(P + 8 + hw8);
            stride4 = 4 * static_cast<int>(P[0]);
            ngroups = P[1];
#pragma unroll
            for (int m = 0; m < kRwPix / 2; ++m) {
                if constexpr (EPW == 2) {
                    const int word = m * kRwThreads + threadIdx.x;
                    cw[m] = word < nwords ? __ldg(reinterpret_cast<const uint32_t*>(g_code) + word) : (dummy | (dummy << 16));
                } else {
                    const int w0 = (2 * m) * kRwThreads + threadIdx.x, w1 = (2 * m + 1) * kRwThreads + threadIdx.x;
                    const uint32_t lo = w0 < nwords ? __ldg(g_code + w0) : dummy, hi = w1 < nwords ? __ldg(g_code + w1) : dummy;
                    cw[m] = lo | (hi << 16);
                }
            }
        }
        // push: every pixel into its own slot
#pragma unroll
        for (int k = 0; k < kRwPix; ++k) {
            const uint32_t off = (cw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
            if (off != dummy) {   // pixels without a source push nothing (their slot offset is only a marker)
#pragma unroll
                for (int n = 0; n < NP; ++n) *reinterpret_cast<float*>(smem + n * pitch + off) = word_elem<T>(g[n][k / EPW], k % EPW);
            }
        }
        __syncthreads();
        if (pass + 1 < pass1) load(pass + 1);   // in flight while the groups are folded and the sums leave
        // groups: tails folded in ascending p
        for (int i = threadIdx.x; i < ngroups; i += kRwThreads) {
            const uint2 e = __ldg(g_group + i);
            const uint32_t q = e.x & 0xffffu;
            uint32_t t = tail0 + (e.x >> 16) * 4u;
            float acc[NP];
#pragma unroll
            for (int n = 0; n < NP; ++n) acc[n] = *reinterpret_cast<const float*>(smem + n * pitch + q);
#pragma unroll 1
            for (uint32_t j = e.y; j != 0u; --j, t += 4u) {
#pragma unroll
                for (int n = 0; n < NP; ++n) acc[n] += *reinterpret_cast<const float*>(smem + n * pitch + t);
            }
#pragma unroll
            for (int n = 0; n < NP; ++n) *reinterpret_cast<float*>(smem + n * pitch + q) = acc[n];
        }
        __syncthreads();
        // read-out: the sums leave as T, the accumulators return to zero for the next pass
#pragma unroll
        for (int v = 0; v < QUADS; ++v) {
            if (ro_row[v] >= 0) {
                const int ro = ro_row[v] * stride4 + ro_col[v];
#pragma unroll
                for (int n = 0; n < NP; ++n) {
                    if (n < np) {
                        float4* sp = reinterpret_cast<float4*>(smem + n * pitch + ro);
                        const float4 x = *sp;
                        *sp = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        store_quad<T>(gin + (static_cast<int64_t>(b) * a.C + c + n) * hw, v * kRwThreads + threadIdx.x, x);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// The same backward for the 2-byte types, with TWO planes per 32-bit slot.  A slot only has to hold a pixel's
// ORIGINAL value: the sums are formed in fp32 registers by the thread that folds a group (ascending p, exactly as
// above) and rounded to T once — so nothing is lost by keeping the slots in T, and the planes (2m, 2m+1) of a pass
// share one word per slot: one PRMT + one STS pushes a pixel of both planes, shared-memory traffic per pixel halves
// (the scatter's bank conflicts made shared-memory wavefronts the floor of the fp32-slot version: 18 us of 38 at
// C5), and a pass covers 2 * NPAIR planes in the footprint of NPAIR.
// dynamic smem: NPAIR x (uint32 acc[acc_words] | uint32 tail[hw8] | 16 bytes)
template <typename T> struct Pair2;
template <> struct Pair2<__half> {
    static __device__ __forceinline__ float lo(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w & 0xffffu))); }
    static __device__ __forceinline__ float hi(uint32_t w) { return __half2float(__ushort_as_half(static_cast<unsigned short>(w >> 16))); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) { const __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
    static __device__ __forceinline__ uint32_t plus_zero(uint32_t w) {   // -0.0 -> +0.0 in both halves
        const __half2 z = __floats2half2_rn(0.0f, 0.0f);
        const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&w), z);
        return *reinterpret_cast<const uint32_t*>(&r);
    }
};
template <> struct Pair2<__nv_bfloat16> {
    static __device__ __forceinline__ float lo(uint32_t w) { return __uint_as_float(w << 16); }
    static __device__ __forceinline__ float hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
    static __device__ __forceinline__ uint32_t pack(float a, float b) { const __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<const uint32_t*>(&h); }
    static __device__ __forceinline__ uint32_t plus_zero(uint32_t w) {
        const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
        const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&w), z);
        return *reinterpret_cast<const uint32_t*>(&r);
    }
};

// The planes of a pass are contiguous in global memory: ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier) brings the
// next pass into a staging buffer while the current one is folded and read out.  (Prefetching into registers — 32
// outstanding LDG per thread — did not overlap anything: the loads share the warp's few scoreboards with the fold's
// own LDS / LDG, so the fold waited for HBM at its first dependent instruction; ncu: 44 % of the warp time in the
// fold phase on long_scoreboard, and the same ~29 us at C5 for 256, 512 and 1024 threads per CTA.)
// FULL: planes of exactly 4096 pixels (the 64 x 64 heatmaps) — every thread slot is live and the output addresses of
// a pass with all of its planes are one base pointer plus compile-time offsets.
// dynamic smem: NPAIR x (uint32 acc[acc_words] | uint32 tail[hw8] | 16 bytes) | T stage[2 * NPAIR][hw8] | mbarrier
template <typename T, int NPAIR, bool FULL, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
rewarp_bwd_push2_kernel(const RewarpArgs a, const T* __restrict__ gout, T* __restrict__ gin, int acc_words,
                        const uint16_t* __restrict__ plan) {
    static_assert(sizeof(T) == 2, "packed pairs of 2-byte elements");
    constexpr int NP = 2 * NPAIR;
    constexpr int PIX = kRwThreads * kRwPix / NT;   // pixels per thread and plane: 16 / 8 for 256 / 512 threads
    constexpr int SLOTS = PIX / 2;
    constexpr int QUADS = PIX / 4;
    extern __shared__ __align__(16) uint32_t rw_smem[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(rw_smem);
    const int hw = FULL ? kRwThreads * kRwPix : a.H * a.W;
    const int hw8 = (hw + 7) & ~7, nwords = hw / 2, nquads = hw / 4;
    const PassOrder order(a.B, a.C, NP);
    const int pass0 = order.first(blockIdx.x, gridDim.x), pass1 = order.first(blockIdx.x + 1, gridDim.x);
    if (pass0 >= pass1) return;
    const int pitch = push_pitch_bytes(acc_words, hw);
    const uint32_t tail0 = static_cast<uint32_t>(acc_words) * 4u, dummy = tail0 + static_cast<uint32_t>(hw8) * 4u;
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem + NPAIR * pitch);                 // [NP][nwords]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NPAIR * pitch + NP * hw8 * 2);
    auto fetch = [&](int pass) {   // one thread: the planes of `pass` into the staging buffer
        int b, c;
        order.at(pass, b, c);
        const uint32_t bytes = static_cast<uint32_t>(min(NP, a.C - c)) * static_cast<uint32_t>(hw) * 2u;
        mbar_expect_tx(bar, bytes);
        bulk_load(stage, gout + (static_cast<int64_t>(b) * a.C + c) * hw, bytes, bar);
    };
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
        fetch(pass0);
    }
    // the accumulator quads this thread reads out: row and 4 * column of quad t + NT * v (one division per thread)
    int ro_row[QUADS], ro_col[QUADS];
    {
        const int r0 = (4 * static_cast<int>(threadIdx.x)) / a.W, c0 = 4 * static_cast<int>(threadIdx.x) - r0 * a.W;
        const bool even = (4 * NT) % a.W == 0;        // a quad step of NT is a whole number of rows
#pragma unroll
        for (int v = 0; v < QUADS; ++v) {
            const int vi = v * NT + threadIdx.x;
            const int row = even ? r0 + v * ((4 * NT) / a.W) : (vi * 4) / a.W;
            ro_row[v] = (FULL || vi < nquads) ? row : -1;
            ro_col[v] = (even ? c0 : vi * 4 - row * a.W) * 4;
        }
    }
#pragma unroll
    for (int m = 0; m < NPAIR; ++m)
        for (int i = threadIdx.x; i < acc_words / 4; i += NT) reinterpret_cast<uint4*>(smem + m * pitch)[i] = make_uint4(0u, 0u, 0u, 0u);
    uint32_t cw[SLOTS];   // the slots of this thread's pixels (word t + NT*slot: pixels 2w, 2w+1), 16 bits each
    int cur_b = -1, stride4 = 0, ngroups = 0;
    const uint2* g_group = nullptr;
    uint2 e_first = make_uint2(0u, 1u);   // this thread's first group of the sample (kept across the sample's passes)
    uint32_t parity = 0;
    __syncthreads();
    for (int pass = pass0; pass < pass1; ++pass) {
        int b, c;
        order.at(pass, b, c);
        if (b != cur_b) {   // CTA-uniform: the per-sample state
            cur_b = b;
            const uint16_t* P = plan + static_cast<int64_t>(b) * plan_elems_for(hw);
            g_group = reinterpret_cast<const uint2*>(P + 8 + hw8);
            stride4 = 4 * static_cast<int>(P[0]);
            ngroups = P[1];
#pragma unroll
            for (int m = 0; m < SLOTS; ++m) {
                const int word = m * NT + threadIdx.x;
                cw[m] = (FULL || word < nwords) ? __ldg(reinterpret_cast<const uint32_t*>(P + 8) + word) : (dummy | (dummy << 16));
            }
            if (static_cast<int>(threadIdx.x) < ngroups) e_first = __ldg(g_group + threadIdx.x);
        }
        const int live = (min(NP, a.C - c) + 1) / 2;   // plane pairs of this pass (the last pass of a sample may hold fewer)
        mbar_wait(bar, parity);
        parity ^= 1u;
        // push: pixel (2w + e) of planes (2m, 2m + 1) as one word into its slot.  (Planes beyond C hold whatever
        // the buffer held before: they only ever reach the unused half of a pair.)
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {
            const int word = sl * NT + threadIdx.x;
            if (FULL || word < nwords) {
                const uint32_t o0 = cw[sl] & 0xffffu, o1 = cw[sl] >> 16;
#pragma unroll
                for (int m = 0; m < NPAIR; ++m) {
                    if (m < live) {
                        const uint32_t ga = stage[(2 * m) * nwords + word], gb = stage[(2 * m + 1) * nwords + word];
                        // (pixels without a source push nothing: their slot offset is only a marker)
                        if (o0 != dummy) *reinterpret_cast<uint32_t*>(smem + m * pitch + o0) = __byte_perm(ga, gb, 0x5410);
                        if (o1 != dummy) *reinterpret_cast<uint32_t*>(smem + m * pitch + o1) = __byte_perm(ga, gb, 0x7632);
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && pass + 1 < pass1) fetch(pass + 1);   // in flight while the groups are folded and the sums leave
        // groups: fp32 sums in ascending p, one rounding, back into the source pixel's slot.  Every group has at
        // least one tail entry; the groups with exactly one come first, so most warps never enter the loop.
        // (A group entry is 8 bytes from L1 / L2: the next one is requested before the current one is folded.)
        uint2 e = e_first;
        for (int i = threadIdx.x; i < ngroups; i += NT) {
            const uint2 nxt = i + NT < ngroups ? __ldg(g_group + i + NT) : e;
            const uint32_t q = e.x & 0xffffu;
            uint32_t t = tail0 + (e.x >> 16) * 4u;
            float s0[NPAIR], s1[NPAIR];
#pragma unroll
            for (int m = 0; m < NPAIR; ++m) {
                if (m < live) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(smem + m * pitch + q);
                    const uint32_t x = *reinterpret_cast<const uint32_t*>(smem + m * pitch + t);
                    s0[m] = Pair2<T>::lo(w) + Pair2<T>::lo(x); s1[m] = Pair2<T>::hi(w) + Pair2<T>::hi(x);
                }
            }
#pragma unroll 1
            for (uint32_t j = e.y - 1u; j != 0u; --j) {
                t += 4u;
#pragma unroll
                for (int m = 0; m < NPAIR; ++m) {
                    if (m < live) {
                        const uint32_t x = *reinterpret_cast<const uint32_t*>(smem + m * pitch + t);
                        s0[m] += Pair2<T>::lo(x); s1[m] += Pair2<T>::hi(x);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < NPAIR; ++m)
                if (m < live) *reinterpret_cast<uint32_t*>(smem + m * pitch + q) = Pair2<T>::pack(s0[m], s1[m]);
            e = nxt;
        }
        __syncthreads();
        // read-out: four source pixels of a plane pair per 128-bit load, planes separated by PRMT, slots re-zeroed
        uint2* o2 = reinterpret_cast<uint2*>(gin + (static_cast<int64_t>(b) * a.C + c) * hw) + threadIdx.x;
        if (FULL && c + NP <= a.C) {
#pragma unroll
            for (int v = 0; v < QUADS; ++v) {
                const int ro = ro_row[v] * stride4 + ro_col[v];
#pragma unroll
                for (int m = 0; m < NPAIR; ++m) {
                    uint4* sp = reinterpret_cast<uint4*>(smem + m * pitch + ro);
                    uint4 w = *sp;
                    *sp = make_uint4(0u, 0u, 0u, 0u);
                    w.x = Pair2<T>::plus_zero(w.x); w.y = Pair2<T>::plus_zero(w.y); w.z = Pair2<T>::plus_zero(w.z); w.w = Pair2<T>::plus_zero(w.w);
                    o2[(2 * m) * (NT * QUADS) + v * NT] = make_uint2(__byte_perm(w.x, w.y, 0x5410), __byte_perm(w.z, w.w, 0x5410));
                    o2[(2 * m + 1) * (NT * QUADS) + v * NT] = make_uint2(__byte_perm(w.x, w.y, 0x7632), __byte_perm(w.z, w.w, 0x7632));
                }
            }
        } else {
#pragma unroll
            for (int v = 0; v < QUADS; ++v) {
                if (ro_row[v] >= 0) {
                    const int ro = ro_row[v] * stride4 + ro_col[v];
#pragma unroll
                    for (int m = 0; m < NPAIR; ++m) {
                        if (c + 2 * m < a.C) {
                            uint4* sp = reinterpret_cast<uint4*>(smem + m * pitch + ro);
                            uint4 w = *sp;
                            *sp = make_uint4(0u, 0u, 0u, 0u);
                            w.x = Pair2<T>::plus_zero(w.x); w.y = Pair2<T>::plus_zero(w.y); w.z = Pair2<T>::plus_zero(w.z); w.w = Pair2<T>::plus_zero(w.w);
                            o2[static_cast<int64_t>(2 * m) * nquads + v * NT] = make_uint2(__byte_perm(w.x, w.y, 0x5410), __byte_perm(w.z, w.w, 0x5410));
                            if (c + 2 * m + 1 < a.C)
                                o2[static_cast<int64_t>(2 * m + 1) * nquads + v * NT] = make_uint2(__byte_perm(w.x, w.y, 0x7632), __byte_perm(w.z, w.w, 0x7632));
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

// dynamic smem: kRwRing padded plane buffers | uint16 map[hw8].  Single view, no paste / pass-through (the heatmap re-warps).
// RING - 1 planes are in flight per CTA: with two CTAs per SM a ring of 3 keeps 32 KB of 8 KB fp16 planes in
// flight per SM — 4.7 MB on the chip, which at ~1.5 us of loaded latency is 3.2 TB/s, exactly where the fp16
// gather sat (51 % of the roofline) while the fp32 gather (16 KB planes) reached 80 %: fp16 / bf16 take a ring of 6.
template <typename T, int RING>
__global__ void __launch_bounds__(kWideThreads, 2)
rewarp_wide_kernel(const RewarpArgs a, T* __restrict__ out, int buf_words) {
    constexpr int EPW = 4 / static_cast<int>(sizeof(T));
    constexpr int SLOTS = kWidePix / EPW;                  // words per thread and plane
    constexpr int VEC = (kWidePix * static_cast<int>(sizeof(T)) * kWideThreads / 16 + kWideThreads - 1) / kWideThreads;  // 16-byte copies per thread and plane
    extern __shared__ __align__(16) uint32_t rw_smem[];
    __shared__ float s_theta[kRwMaxStages * 6];
    const int hw = a.H * a.W, nwords = hw / EPW;
    const int b = blockIdx.x;
    if (threadIdx.x < a.stages * 6)
        s_theta[threadIdx.x] = a.view[0].theta[static_cast<int64_t>(b) * a.stages * 6 + threadIdx.x];
    const uint32_t zero_byte = static_cast<uint32_t>(buf_words - 4) * 4u;
    if (threadIdx.x < RING) rw_smem[threadIdx.x * buf_words + buf_words - 4] = 0u;   // out-of-bounds pixels gather from it
    __syncthreads();
    const int row_words = a.W / EPW, vpr = row_words / 4, nvec = nwords / 4;
    float J[4];
    composed_jacobian(s_theta, a, J);
    const int stride = pick_stride(row_words, J[0], J[2], EPW - 1);
    int so[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        const int v = q * kWideThreads + threadIdx.x;
        const int row = v / vpr;
        so[q] = v < nvec ? (row * stride + 4 * (v - row * vpr)) * 4 : -1;
    }
    auto issue = [&](int it) {
        const uint4* src = reinterpret_cast<const uint4*>(static_cast<const T*>(a.view[0].in) + (static_cast<int64_t>(b) * a.C + it) * hw) + threadIdx.x;
        const uint32_t dst = smem_u32(rw_smem + (it % RING) * buf_words);
#pragma unroll
        for (int q = 0; q < VEC; ++q)
            if (so[q] >= 0) cp_async16(dst + so[q], src + q * kWideThreads);
    };
    const int nitems = a.C;
#pragma unroll
    for (int it = 0; it < RING - 1; ++it) {   // in flight while the map is computed
        if (it < nitems) issue(it);
        cp_async_commit();
    }
    // the composed source of every pixel as a padded byte offset, then this thread's 8 (words t + NT*slot) in registers
    uint16_t* map = reinterpret_cast<uint16_t*>(rw_smem + RING * buf_words);
    build_map_compact(s_theta, a, map, static_cast<uint16_t>(zero_byte), [=](int i, int j) {
        return static_cast<uint16_t>(j * stride * 4 + i * static_cast<int>(sizeof(T)));
    });
    __syncthreads();
    uint32_t idx[kWidePix / 2];
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
        const int word = sl * kWideThreads + threadIdx.x;
#pragma unroll
        for (int e = 0; e < EPW; ++e) {
            const int k = sl * EPW + e;
            const uint32_t o = word < nwords ? map[word * EPW + e] : zero_byte;
            if (k & 1) idx[k >> 1] |= o << 16;
            else idx[k >> 1] = o;
        }
    }
    for (int it = 0; it < nitems; ++it) {
        uint32_t* o32 = reinterpret_cast<uint32_t*>(out + (static_cast<int64_t>(b) * a.C + it) * hw) + threadIdx.x;
        cp_async_wait<RING - 2>();   // this thread's copies of plane `it` have landed ...
        __syncthreads();                // ... everybody's have, and everybody is done with plane it-1
        if (it + RING - 1 < nitems) issue(it + RING - 1);
        cp_async_commit();
        const uint8_t* bytes = reinterpret_cast<const uint8_t*>(rw_smem + (it % RING) * buf_words);
#pragma unroll
        for (int sl = 0; sl < SLOTS; ++sl) {   // the values are moved, never converted
            uint32_t w32;
            if constexpr (EPW == 1) {
                const uint32_t o = (idx[sl >> 1] >> (16 * (sl & 1))) & 0xffffu;
                w32 = *reinterpret_cast<const uint32_t*>(bytes + o);
            } else {
                const uint32_t lo = *reinterpret_cast<const uint16_t*>(bytes + (idx[sl] & 0xffffu));
                const uint32_t hi = *reinterpret_cast<const uint16_t*>(bytes + (idx[sl] >> 16));
                w32 = lo | (hi << 16);
            }
            if (sl * kWideThreads + static_cast<int>(threadIdx.x) < nwords) o32[sl * kWideThreads] = w32;
        }
    }
    cp_async_wait<0>();
}

// ---- second-generation wide forward (full 4096-pixel planes, single view) -------------------------------------
// ncu on the first wide kernel at C5 (profiles/r02q): the fp16 gather was bound by shared-memory WAVEFRONTS, not
// by HBM — 5.65 M per launch = 20 us of its 25: LDS.U16 gathers at 2.25 x their ideal (bank conflicts of the padded
// rows), LDGSTS staging writes at 2.6 x, and three LDS.64 of theta per stage and pixel in the map builder; and it
// paid one CTA barrier, one cp.async wait and a set of predicates / 64-bit address computations per PLANE.  Here:
//   * rows are not padded but XOR-swizzled at 16-byte granularity (chunk ^ row): a plane is exactly H*W elements
//     (+ one zero chunk that out-of-image pixels read).  (A word-granular skew staged with 4-byte cp.async was
//     measured too: fp32 41.3 us, fp16 26.5 us at C5 against 31.5 / 24.1 — four times the copy instructions.)
//   * an item of the ring is G planes (4 x 8 KB or 2 x 16 KB): one barrier and one wait per 32 KB;
//   * the stage table lives in registers while the map is built;
//   * every slot of every thread is live (H*W = 512 threads x 8 pixels): no predicates, addresses are one base
//     pointer per plane plus compile-time offsets.
constexpr int kW2Threads = 512;
constexpr int kW2Pixels = 4096;

// dynamic smem: RING stages x G planes x (H*W*sizeof(T) + 16 zero bytes) | uint16 map[4096]
// DECODE: the gathered planes are not stored but arg-maxed where they are gathered (udape_rewarp_decode_select: the
// teacher chain of the step only ever decodes its re-warped map, train_human.py:359-383, :427-430).  Every thread
// keeps, per plane of an item, the best (ordered key, ~output pixel) of its eight pixels; warps leave their best in
// a [planes][16] table and after the last item thread c folds row c and writes plane c's decode outputs; the CTA
// that takes the last ticket runs the k-th value select over the batch.  Same keys, same tie rule (first output
// pixel) and same zero for pixels that left the image as decoding the stored map: bit-identical outputs.
struct RewarpDecodeOut {
    int32_t* idx;            // [B*C] or NULL
    float* preds;            // [B*C, 2] or NULL
    float* maxvals_f32;      // [B*C]
    int64_t* position;       // [B*C, 2] or NULL
    uint8_t* conf_table;     // [B*C] or NULL
    float occlude_thresh;
    SelectArgs sel;          // ticket == NULL: no select
};
constexpr int kW2DecodeMaxC = 64;   // planes per sample of the decode route: its [C][16] table of warp maxima reuses the map's 8 KB

template <typename T> __device__ __forceinline__ float word_value(uint32_t w32, int e);
template <> __device__ __forceinline__ float word_value<float>(uint32_t w32, int) { return __uint_as_float(w32); }
template <> __device__ __forceinline__ float word_value<__half>(uint32_t w32, int e) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(e ? (w32 >> 16) : (w32 & 0xffffu))));
}
template <> __device__ __forceinline__ float word_value<__nv_bfloat16>(uint32_t w32, int e) {
    return __uint_as_float(e ? (w32 & 0xffff0000u) : (w32 << 16));
}

template <typename T, int G, int RING, bool DECODE = false>
__global__ void __launch_bounds__(kW2Threads, 2)
rewarp_wide2_kernel(const RewarpArgs a, T* __restrict__ out, const RewarpDecodeOut dec = RewarpDecodeOut{}) {
    constexpr int ES = static_cast<int>(sizeof(T)), EPW = 4 / ES;
    constexpr int PIX = kW2Pixels / kW2Threads;         // 8 pixels per thread and plane
    constexpr int SLOTS = PIX / EPW;                     // words per thread and plane
    constexpr int WORDS = kW2Pixels / EPW;               // 32-bit words of a plane
    constexpr int CHUNKS = kW2Pixels * ES / 16;          // 16-byte chunks of a plane
    constexpr int VEC = CHUNKS / kW2Threads;             // staging copies per thread and plane
    constexpr int PITCH = kW2Pixels * ES + 16;           // a staged plane and its zero chunk
    constexpr int STAGE = G * PITCH;
    extern __shared__ __align__(16) uint32_t rw_smem[];
    __shared__ float s_theta[kRwMaxStages * 6];
    uint8_t* smem = reinterpret_cast<uint8_t*>(rw_smem);
    uint16_t* map = reinterpret_cast<uint16_t*>(smem + RING * STAGE);
    const int b = blockIdx.x;
    if (threadIdx.x < a.stages * 6)
        s_theta[threadIdx.x] = a.view[0].theta[static_cast<int64_t>(b) * a.stages * 6 + threadIdx.x];
    for (int i = threadIdx.x; i < RING * G * 4; i += kW2Threads)       // the zero chunk behind every staged plane
        *reinterpret_cast<uint32_t*>(smem + (i >> 2) * PITCH + kW2Pixels * ES + (i & 3) * 4) = 0u;
    // chunk c of a plane = row (c / cpr), chunk-in-row (c % cpr); staged at chunk-in-row ^ (row % 8)
    const int cpr = a.W * ES / 16, cpr_log2 = 31 - __clz(cpr), swz = min(cpr, 8) - 1;
    uint32_t so[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        const int c = q * kW2Threads + threadIdx.x, row = c >> cpr_log2, cc = c & (cpr - 1);
        so[q] = static_cast<uint32_t>(((row << cpr_log2) + (cc ^ (row & swz))) * 16);
    }
    const int nitems = (a.C + G - 1) / G;
    const uint32_t smem0 = smem_u32(smem);
    const uint4* in0 = reinterpret_cast<const uint4*>(static_cast<const T*>(a.view[0].in) + static_cast<int64_t>(b) * a.C * kW2Pixels) + threadIdx.x;
    auto issue = [&](int it) {
        const uint32_t dst = smem0 + (it % RING) * STAGE;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (it * G + g < a.C) {
#pragma unroll
                for (int q = 0; q < VEC; ++q) cp_async16(dst + g * PITCH + so[q], in0 + static_cast<int64_t>(it * G + g) * CHUNKS + q * kW2Threads);
            }
        }
    };
#pragma unroll
    for (int it = 0; it < RING - 1; ++it) {   // in flight while the map is computed
        if (it < nitems) issue(it);
        cp_async_commit();
    }
    __syncthreads();   // theta
    {
        const int W = a.W;
        build_map_fast(s_theta, a, map, static_cast<uint16_t>(kW2Pixels * ES), [=](int i, int j) {
            const int byte = i * ES, cc = byte >> 4;
            return static_cast<uint16_t>((((j * W * ES) >> 4) + (cc ^ (j & swz))) * 16 + (byte & 15));
        });
    }
    __syncthreads();
    // Which output words a thread owns: word tw + 512 * slot.  tw = threadIdx.x makes the 32 lanes of a gather walk
    // one output row, i.e. a LINE of the source plane: whatever the row layout, some rotation makes that line collide
    // in the banks (ncu, C5 fp16: LDS.U16 at 2.4-2.8 wavefronts each, the kernel bound by them).  With tile_log2 = t
    // the lanes cover a compact tile of 2^t words x 2^(5-t) rows instead, which maps to a compact patch of the source
    // for every rotation: the XOR swizzle spreads the patch's rows over the chunk columns.  Measured at C5, fp16
    // (profiles/r02ai): row walk 24.0 us, 16 x 4 px 21.0 us (t = 3, the default), 32 x 2 px 21.3, 8 x 8 px 26.4 (its
    // stores fill half a 32-byte sector per row); other row keys of the swizzle (bit-reversed row) change nothing;
    // fp32 planes keep the row walk (31.8 us against 33.1 / 47.2 tiled).
    int tw = threadIdx.x;
    if (a.tile_log2 > 0) {
        const int tl = a.tile_log2, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int row_words = a.W / EPW, tpr_log2 = (31 - __clz(row_words)) - tl;    // tiles per row
        const int th_log2 = 5 - tl;                                                  // rows of a tile
        tw = ((((w >> tpr_log2) << th_log2) + (lane >> tl)) * row_words) + ((w & ((1 << tpr_log2) - 1)) << tl) + (lane & ((1 << tl) - 1));
    }
    uint32_t idx[PIX / 2];   // the staged byte offsets of this thread's pixels (words tw + 512*slot), two per register
#pragma unroll
    for (int k = 0; k < PIX; k += 2) {
        const int p0 = ((k / EPW) * kW2Threads + tw) * EPW + (k % EPW);
        const int p1 = (((k + 1) / EPW) * kW2Threads + tw) * EPW + ((k + 1) % EPW);
        idx[k >> 1] = static_cast<uint32_t>(map[p0]) | (static_cast<uint32_t>(map[p1]) << 16);
    }
    uint32_t* o32 = DECODE ? nullptr : reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(b) * a.C * kW2Pixels) + tw;
    // (the [C][16] table of warp maxima takes the map's place: the map lives on in `idx` once the item loop starts, and
    // the loop's first barrier separates its last read from the first write here; 64 planes x 16 warps x 8 bytes = 8 KB)
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(map);
    for (int it = 0; it < nitems; ++it) {
        cp_async_wait<RING - 2>();   // this thread's copies of item `it` have landed ...
        __syncthreads();                // ... everybody's have, and everybody is done with item it-1
        if (it + RING - 1 < nitems) issue(it + RING - 1);
        cp_async_commit();
        const uint8_t* base = smem + (it % RING) * STAGE;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (it * G + g < a.C) {
                const uint8_t* bytes = base + g * PITCH;
                uint32_t* o = DECODE ? nullptr : o32 + static_cast<int64_t>(it * G + g) * WORDS;
                float vals[DECODE ? PIX : 1];   // DECODE: this thread's eight gathered values, in ascending output pixel
#pragma unroll
                for (int sl = 0; sl < SLOTS; ++sl) {   // the values are moved, never converted
                    uint32_t w32;
                    if constexpr (EPW == 1) {
                        w32 = *reinterpret_cast<const uint32_t*>(bytes + ((idx[sl >> 1] >> (16 * (sl & 1))) & 0xffffu));
                    } else {
                        const uint32_t lo = *reinterpret_cast<const uint16_t*>(bytes + (idx[sl] & 0xffffu));
                        const uint32_t hi = *reinterpret_cast<const uint16_t*>(bytes + (idx[sl] >> 16));
                        w32 = lo | (hi << 16);
                    }
                    if constexpr (DECODE) {
#pragma unroll
                        for (int e = 0; e < EPW; ++e) vals[sl * EPW + e] = word_value<T>(w32, e);
                    } else {
                        o[sl * kW2Threads] = w32;
                    }
                }
                if constexpr (DECODE) {
                    // one NaN-propagating max per pixel, then the FIRST pixel that equals it (floating == ties -0.0
                    // with +0.0; a NaN maximum matches the first NaN) — the per-pixel 64-bit keys of the first version
                    // doubled the kernel's instruction count (42.7 us at C5 against 31.8 for the plain gather)
                    float m = vals[0];
#pragma unroll
                    for (int q = 1; q < PIX; ++q) m = fmax_nan(m, vals[q]);
                    const bool m_nan = m != m;
                    int first = PIX - 1;
#pragma unroll
                    for (int q = PIX - 2; q >= 0; --q) first = (vals[q] == m || (m_nan && vals[q] != vals[q])) ? q : first;
                    const uint32_t p = static_cast<uint32_t>(((first / EPW) * kW2Threads + tw) * EPW + (first % EPW));   // its output pixel
                    unsigned long long best = warp_max_u64(pack_arg(order_key(m), p));
                    if ((threadIdx.x & 31) == 0) s_best[(it * G + g) * (kW2Threads / 32) + (threadIdx.x >> 5)] = best;
                }
            }
        }
    }
    cp_async_wait<0>();
    if constexpr (DECODE) {
        __syncthreads();
        if (static_cast<int>(threadIdx.x) < a.C) {
            unsigned long long best = 0ull;
#pragma unroll
            for (int w = 0; w < kW2Threads / 32; ++w) {
                const unsigned long long v = s_best[threadIdx.x * (kW2Threads / 32) + w];
                best = v > best ? v : best;
            }
            // the outputs of decode_kernel (decode.cu), for plane b * C + c
            const int64_t plane = static_cast<int64_t>(b) * a.C + threadIdx.x;
            const uint32_t pidx = arg_idx(best);
            const float mv = key_value(arg_key(best));
            const int ix = static_cast<int>(pidx % static_cast<uint32_t>(a.W)), iy = static_cast<int>(pidx / static_cast<uint32_t>(a.W));
            const bool positive = mv > 0.0f;   // false for NaN, like np.greater / torch.gt
            if (dec.idx) dec.idx[plane] = static_cast<int32_t>(pidx);
            if (dec.preds) {
                dec.preds[2 * plane] = positive ? static_cast<float>(ix) : 0.0f;
                dec.preds[2 * plane + 1] = positive ? static_cast<float>(iy) : 0.0f;
            }
            dec.maxvals_f32[plane] = mv;
            if (dec.position) { dec.position[2 * plane] = ix; dec.position[2 * plane + 1] = iy; }
            if (dec.conf_table) dec.conf_table[plane] = (mv >= dec.occlude_thresh) ? 1 : 0;
        }
        __syncthreads();   // every writer is done before thread 0 fences and takes the ticket (fence cumulativity, as in pck_tma_kernel)
        // train_human.py:427-430 in the same launch: the CTA that finishes last selects the k-th activation; the
        // staging ring is idle by now and caches the batch's values
        if (dec.sel.ticket != nullptr && last_block_done(dec.sel.ticket, gridDim.x))
            select_body(dec.maxvals_f32, a.B * a.C, dec.sel, reinterpret_cast<float*>(smem), RING * STAGE / 4);
    }
}

// general route (any plane up to 25600 px): gradient gathered from global memory.
// dynamic smem: uint16 off[hw + 2] | uint16 lst[hw] | uint16 map[hw]
template <typename T>
__global__ void __launch_bounds__(kRwThreads)
rewarp_bwd_kernel(const RewarpArgs a, const T* __restrict__ gout, T* __restrict__ gin) {
    extern __shared__ __align__(16) uint32_t rw_smem[];
    __shared__ float s_theta[kRwMaxStages * 6];
    __shared__ uint32_t s_scan[kRwThreads];
    const int hw = a.H * a.W;
    uint16_t* off = reinterpret_cast<uint16_t*>(rw_smem);
    uint16_t* lst = off + ((hw + 2 + 7) & ~7);
    uint16_t* map = lst + ((hw + 7) & ~7);
    const int cgroups = (a.C + a.cpc - 1) / a.cpc;
    const int cg = blockIdx.x % cgroups, b = blockIdx.x / cgroups;
    const int c0 = cg * a.cpc, c1 = min(a.C, c0 + a.cpc);
    if (threadIdx.x < a.stages * 6)
        s_theta[threadIdx.x] = a.view[0].theta[static_cast<int64_t>(b) * a.stages * 6 + threadIdx.x];
    uint32_t* off32 = reinterpret_cast<uint32_t*>(off);
    for (int s = threadIdx.x; s < (hw + 2) / 2; s += kRwThreads) off32[s] = 0u;
    __syncthreads();
    build_map_slice(s_theta, a, map, 0, hw);
    __syncthreads();
    invert_map(map, hw, 0, hw, off, lst, s_scan, [](int p) { return static_cast<uint16_t>(p); });
    for (int c = c0; c < c1; ++c) {
        const int64_t base = (static_cast<int64_t>(b) * a.C + c) * hw;
        const T* go = gout + base;
        for (int s = threadIdx.x; s < hw; s += kRwThreads) {
            const int st = s == 0 ? 0 : off[s - 1], en = off[s];
            float acc = 0.0f;
            for (int q = st; q < en; ++q) acc += to_f32<T>(go[lst[q]]);
            gin[base + s] = from_f32<T>(acc);
        }
    }
}

// channels per CTA: amortise the per-sample index work over channels while keeping ~`target` CTAs
static int channels_per_cta(int64_t units, int64_t c, int64_t target) {
    int64_t cpc = (units * c + target - 1) / target;
    if (cpc < 1) cpc = 1;
    if (cpc > c) cpc = c;
    const int64_t groups = (c + cpc - 1) / cpc;
    return static_cast<int>((c + groups - 1) / groups);  // even split
}

static int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;
    return n;
}

// the shared-memory route needs 16-byte rows, planes of at most 4096 pixels and a padded plane
// that fits the per-buffer budget; returns the buffer size in words, 0 if the route does not apply
static int smem_route_words(int64_t H, int64_t W, int elem_bytes) {
    if ((W * elem_bytes) % 16 != 0 || H * W > kRwThreads * kRwPix || (H * W) % 8 != 0) return 0;
    const int row_words = static_cast<int>(W * elem_bytes / 4);
    const int sa = padded_stride(row_words, 4), sb = padded_stride(row_words, 12);
    const int64_t words = H * (sa > sb ? sa : sb) + 4;   // either padding per sample, + the zero word (16-byte tail)
    return words * 4 <= kRwBufBytes ? static_cast<int>(words) : 0;
}

// The push plan / scatter backward: 16-byte rows, planes of at most 4096 pixels, and a padded fp32 accumulator
// plane small enough that the slot offsets (accumulators | tail | dummy) fit 16 bits and two planes fit a CTA.
// Returns the accumulator plane in words (a multiple of 4), 0 if the route does not apply.
constexpr int kRankAccBytes = 40 * 1024;
static int rank_route_words(int64_t H, int64_t W, int elem_bytes) {
    if ((W * elem_bytes) % 16 != 0 || H * W > kRwThreads * kRwPix || (H * W) % 8 != 0) return 0;
    const int sa = padded_stride(static_cast<int>(W), 4), sb = padded_stride(static_cast<int>(W), 12);
    const int64_t words = (H * (sa > sb ? sa : sb) + 3) & ~3ll;
    return words * 4 <= kRankAccBytes ? static_cast<int>(words) : 0;
}

static bool route_disabled() {
    const char* e = std::getenv("UDAPE_REWARP_GLOBAL");  // tests compare both routes within one process
    return e && e[0] == '1';
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
    a.paste = nullptr; a.active = nullptr; a.paste_after = 0; a.tile_log2 = 0;
    return UDAPE_OK;
}

// CTAs per sample (= cluster size): the smallest power of two that gives at least 32 CTAs, at most the
// portable cluster limit of 8 and the channel count
static int cluster_size_for(int64_t B, int64_t C, int64_t hw) {
    if (const char* e = std::getenv("UDAPE_REWARP_CLUSTER")) {   // tuning / tests: force 1, 2, 4 or 8
        const int n = std::atoi(e);
        if (n == 1 || ((n == 2 || n == 4 || n == 8) && n <= C && hw / n >= 8)) return n;
    }
    // Measured on B200 (64x64 planes, batch 32, tools/step_probe.py): timed ALONE the gather is fastest with
    // 1.5-2 CTAs per SM (clusters of 8: 7.7 us against ~12 us for one CTA per sample), but inside the step —
    // beside the AdaIN / EMA streams and the other heatmap chains — every doubling of the cluster costs
    // ~6 us of step time (184 / 190 / 197 / 214 us for clusters of 1 / 2 / 4 / 8): a cluster needs all of its
    // CTAs placed in one GPC at once, which stalls the block scheduler of a busy GPU, and its CTAs hold SM
    // slots while moving little data.  One CTA per sample from 32 samples up; smaller batches still split.
    const int64_t want = 32;
    int n = 1;
    while (n < 8 && 2 * n <= C && hw / (2 * n) >= 8 && B * n < want) n *= 2;   // slices of the map hold >= 8 entries
    return n;
}

template <typename... KArgs, typename... Args>
static int launch_cluster(void (*kernel)(KArgs...), unsigned grid, unsigned cluster, size_t smem, cudaStream_t st,
                          const char* name, Args... args) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: cannot reserve %zu bytes of shared memory", name, smem);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(kRwThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s (%s)", name, cudaGetErrorName(e), cudaGetErrorString(e));
    return UDAPE_OK;
}

template <typename K>
static int reserve_smem(K kernel, size_t bytes, const char* name) {
    if (bytes <= 48 * 1024) return UDAPE_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: cannot reserve %zu bytes of shared memory", name, bytes);
    return UDAPE_OK;
}

// the wide route takes planes of exactly kWideThreads * kWidePix pixels or fewer in whole 8-pixel groups
static bool wide_route(int64_t B, int64_t hw) {
    const char* e = std::getenv("UDAPE_REWARP_WIDE");   // tests / tuning: "0" keeps the 256-thread kernels
    if (e && e[0] == '0') return false;
    return hw <= kWideThreads * kWidePix && (hw % 8) == 0 && B >= 1;
}

// the second-generation wide route: planes of exactly 4096 pixels whose rows are a power-of-two number of 16-byte chunks
static bool wide2_route(int64_t H, int64_t W, int es) {
    const char* e = std::getenv("UDAPE_REWARP_WIDE2");   // tests / tuning: "0" keeps the first wide kernel
    if (e && e[0] == '0') return false;
    const int64_t cpr = W * es / 16;   // 16-byte chunks per row: a power of two
    return H * W == kW2Pixels && (W * es) % 16 == 0 && cpr >= 1 && (cpr & (cpr - 1)) == 0;
}

// the tile a warp's gather covers (see the kernel): 32 words must split into whole tile rows, and 16 warps x 32
// words = 512 words must be whole rows of tiles
static int wide2_tile_log2(int64_t W, int es) {
    int tl = es == 2 ? 3 : 0;
    if (const char* e = std::getenv("UDAPE_RW_TILE")) tl = std::atoi(e);   // tuning: 0 = row walk, 2 = 8x8 px, 3 = 16x4 px
    const int64_t row_words = W * es / 4;
    if (tl < 1 || tl > 4 || (row_words & (row_words - 1)) != 0 || (1 << tl) > row_words) return 0;
    const int64_t tiles_per_row = row_words >> tl, rows_per_tile = 32 >> tl;
    // a slot of 512 words = 512 / row_words rows must hold whole tile rows
    if ((512 / row_words) % rows_per_tile != 0 || 512 % row_words != 0 || tiles_per_row > 16) return 0;
    return tl;
}

constexpr int kRwDeepRing = 6;   // single view, a CTA that owns >= 6 planes

template <typename T>
static int launch_smem_fwd(const RewarpArgs& a, int views, int ring, unsigned grid, unsigned cluster, size_t smem, cudaStream_t st,
                           T* out, int buf_words) {
    const char* name = "udape_rewarp_fwd";
    switch (views) {
        case 1:
            if (ring == kRwDeepRing)
                return launch_cluster(rewarp_smem_kernel<T, 1, kRwDeepRing>, grid, cluster, smem, st, name, a, out, buf_words);
            return launch_cluster(rewarp_smem_kernel<T, 1, kRwRing>, grid, cluster, smem, st, name, a, out, buf_words);
        case 2: return launch_cluster(rewarp_smem_kernel<T, 2, kRwRing>, grid, cluster, smem, st, name, a, out, buf_words);
        case 3: return launch_cluster(rewarp_smem_kernel<T, 3, kRwRing>, grid, cluster, smem, st, name, a, out, buf_words);
        default: return launch_cluster(rewarp_smem_kernel<T, 4, kRwRing>, grid, cluster, smem, st, name, a, out, buf_words);
    }
}

}  // namespace udape

using namespace udape;

extern "C" int udape_rewarp_fwd(const void* const* in, const float* const* theta, int views, int stages,
                                int half_mask, int grid_dtype, const int32_t* paste, int paste_after,
                                const uint8_t* active, int64_t B, int64_t C, int64_t H, int64_t W, int dtype,
                                void* out, uint16_t* inverse_plan, void* stream) {
    RewarpArgs a = {};
    // out == NULL with an inverse plan: only the plan is built (it depends on theta alone, so a caller can
    // run it on another stream, off the chain  gather -> loss -> backward)
    UDAPE_REQUIRE(in && theta && (out || inverse_plan), UDAPE_ERR_NULL, "udape_rewarp_fwd: NULL pointer");
    const int rc = fill_common(a, "udape_rewarp_fwd", views, stages, half_mask, grid_dtype, B, C, H, W, dtype);
    if (rc) return rc;
    const int es = dtype_size(dtype);
    bool all16 = aligned16(out);
    for (int v = 0; v < views; ++v) {
        UDAPE_REQUIRE((in[v] || !out) && theta[v], UDAPE_ERR_NULL, "udape_rewarp_fwd: NULL view %d", v);
        UDAPE_REQUIRE(aligned_to(in[v], es) && aligned_to(theta[v], 4), UDAPE_ERR_ALIGN, "udape_rewarp_fwd: misaligned view %d", v);
        UDAPE_REQUIRE(!out || in[v] != out, UDAPE_ERR_ARG, "udape_rewarp_fwd: the gather cannot run in place");
        a.view[v].in = in[v];
        a.view[v].theta = theta[v];
        all16 = all16 && aligned16(in[v]);
    }
    UDAPE_REQUIRE(aligned_to(out, es) && (!paste || aligned_to(paste, 4)), UDAPE_ERR_ALIGN, "udape_rewarp_fwd: misaligned pointer");
    UDAPE_REQUIRE(!paste || (paste_after >= 0 && paste_after < stages), UDAPE_ERR_ARG,
                  "udape_rewarp_fwd: paste_after must name an evaluated stage");
    UDAPE_REQUIRE(!active || views == 1, UDAPE_ERR_ARG, "udape_rewarp_fwd: pass-through samples need a single view");
    a.paste = paste; a.paste_after = paste_after; a.active = active;
    const int64_t hw = H * W;
    cudaStream_t st = as_stream(stream);
    const int buf_words = (all16 && !paste && !active && !route_disabled()) ? smem_route_words(H, W, es) : 0;
    if (inverse_plan) {
        const int acc_words = rank_route_words(H, W, es);
        UDAPE_REQUIRE(views == 1 && !paste && !active && acc_words != 0 && aligned16(inverse_plan),
                      UDAPE_ERR_ARG, "udape_rewarp_fwd: an inverse plan needs a single view, no paste / pass-through and a "
                      "plane udape_rewarp_plan_elems() accepts");
        // what the backward needs (a shared-memory slot for every output pixel + the groups to fold), built once per batch
        const size_t smem = sizeof(uint32_t) * static_cast<size_t>(acc_words) + sizeof(uint16_t) * ((hw + 7) & ~7ll);
        if (es == 4) {
            const int r2 = reserve_smem(rewarp_push_plan_kernel<4>, smem, "udape_rewarp_fwd(plan)");
            if (r2) return r2;
            rewarp_push_plan_kernel<4><<<static_cast<unsigned>(B), kPlanThreads, smem, st>>>(a, inverse_plan, acc_words);
        } else {
            const int r2 = reserve_smem(rewarp_push_plan_kernel<2>, smem, "udape_rewarp_fwd(plan)");
            if (r2) return r2;
            rewarp_push_plan_kernel<2><<<static_cast<unsigned>(B), kPlanThreads, smem, st>>>(a, inverse_plan, acc_words);
        }
        const int r3 = check_launch("udape_rewarp_fwd(plan)");
        if (r3 || !out) return r3;
    }
    // (batches that do not fill the GPU are latency-bound per CTA and stay with the per-plane ring: C2, 32 samples,
    // fp16 10.7 us against 11.9 us; C5: 25.9 -> 24.1 us fp16, 33.5 -> 31.5 us fp32)
    if (buf_words && views == 1 && wide_route(B, hw) && wide2_route(H, W, es) && B * C >= 2048) {
        // second-generation wide route: full 4096-pixel planes, swizzled staging, G planes per barrier
        UDAPE_DISPATCH_FLOAT(dtype, T, {
            constexpr int G = sizeof(T) == 2 ? 4 : 2;
            constexpr int RING = 3;
            const size_t smem = static_cast<size_t>(RING) * G * (kW2Pixels * sizeof(T) + 16) + sizeof(uint16_t) * kW2Pixels;
            const int r2 = reserve_smem(rewarp_wide2_kernel<T, G, RING>, smem, "udape_rewarp_fwd");
            if (r2) return r2;
            a.tile_log2 = wide2_tile_log2(W, es);
            rewarp_wide2_kernel<T, G, RING><<<static_cast<unsigned>(B), kW2Threads, smem, st>>>(a, static_cast<T*>(out));
        });
        return check_launch("udape_rewarp_fwd");
    }
    if (buf_words && views == 1 && wide_route(B, hw)) {
        // wide route: one CTA of 512 threads per sample
        UDAPE_DISPATCH_FLOAT(dtype, T, {
            constexpr int RING = sizeof(T) == 2 ? kWideRingHalf : kRwRing;
            const size_t smem = RING * sizeof(uint32_t) * static_cast<size_t>(buf_words) + sizeof(uint16_t) * ((hw + 7) & ~7ll);
            const int r2 = reserve_smem(rewarp_wide_kernel<T, RING>, smem, "udape_rewarp_fwd");
            if (r2) return r2;
            rewarp_wide_kernel<T, RING><<<static_cast<unsigned>(B), kWideThreads, smem, st>>>(a, static_cast<T*>(out), buf_words);
        });
        return check_launch("udape_rewarp_fwd");
    }
    if (buf_words) {
        // one cluster of `n` CTAs per sample: the channels are split n ways, the map is built once
        const int n = cluster_size_for(B, C, hw);
        a.cpc = static_cast<int>((C + n - 1) / n);
        const int64_t grid = B * n;
        const char* e_ring = std::getenv("UDAPE_REWARP_RING");   // tests: force the shallow ring ("3")
        // deep ring only when the launch has at most one CTA per SM (few fat CTAs must keep the bytes in flight
        // themselves); with many CTAs the shallow ring fits three of them per SM and is faster (C5 fp32:
        // 35.7 us shallow, 39.8 us deep)
        const int ring = (views == 1 && a.cpc >= kRwDeepRing && B * n <= sm_count() && !(e_ring && e_ring[0] == '3'))
                             ? kRwDeepRing : kRwRing;
        const size_t smem = ring * sizeof(uint32_t) * static_cast<size_t>(buf_words) + sizeof(uint16_t) * views * ((hw + 7) & ~7ll);
        UDAPE_DISPATCH_FLOAT(dtype, T, {
            const int r2 = launch_smem_fwd<T>(a, views, ring, static_cast<unsigned>(grid), static_cast<unsigned>(n), smem, st,
                                              static_cast<T*>(out), buf_words);
            if (r2) return r2;
        });
        return check_launch("udape_rewarp_fwd");
    }
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        constexpr int EPV = Vec16<T>::EPV;
        const bool vec = (hw % EPV) == 0 && aligned16(out);
        const int px = vec ? EPV : 1;
        const int64_t bands = (hw + kRwThreads * px - 1) / (kRwThreads * px);
        a.cpc = channels_per_cta(B * bands, C, 16 * static_cast<int64_t>(sm_count()));
        const int64_t grid = B * bands * ((C + a.cpc - 1) / a.cpc);
        UDAPE_REQUIRE(grid < (1ll << 31), UDAPE_ERR_SHAPE, "udape_rewarp_fwd: grid too large");
        if (vec) rewarp_fwd_kernel<T, EPV><<<static_cast<unsigned>(grid), kRwThreads, 0, st>>>(a, static_cast<T*>(out));
        else rewarp_fwd_kernel<T, 1><<<static_cast<unsigned>(grid), kRwThreads, 0, st>>>(a, static_cast<T*>(out));
    });
    return check_launch("udape_rewarp_fwd");
}

extern "C" int64_t udape_rewarp_plan_elems(int64_t H, int64_t W, int elem_bytes) {
    if (H <= 0 || W <= 0 || (elem_bytes != 2 && elem_bytes != 4)) return 0;
    return rank_route_words(H, W, elem_bytes) ? plan_elems_for(H * W) : 0;
}

extern "C" int udape_rewarp_decode_select(const void* in, const float* theta, int stages, int half_mask, int grid_dtype,
                                          int64_t B, int64_t C, int64_t H, int64_t W, int dtype, int32_t* idx, float* preds,
                                          float* maxvals_f32, int64_t* position, float occlude_thresh, uint8_t* conf_table,
                                          int64_t kth, const float* tea_mask_in, float* thresh_out, uint8_t* tea_mask_out,
                                          uint32_t* ticket, void* stream) {
    RewarpArgs a = {};
    UDAPE_REQUIRE(in && theta && maxvals_f32, UDAPE_ERR_NULL, "udape_rewarp_decode_select: NULL pointer (in, theta and maxvals_f32 are required)");
    const int rc = fill_common(a, "udape_rewarp_decode_select", 1, stages, half_mask, grid_dtype, B, C, H, W, dtype);
    if (rc) return rc;
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(aligned16(in) && aligned_to(theta, 4) && aligned_to(maxvals_f32, 4) && (!idx || aligned_to(idx, 4)) &&
                      (!preds || aligned_to(preds, 4)) && (!position || aligned_to(position, 8)) &&
                      (!ticket || aligned_to(ticket, 4)) && (!thresh_out || aligned_to(thresh_out, 4)) &&
                      (!tea_mask_in || aligned_to(tea_mask_in, 4)),
                  UDAPE_ERR_ALIGN, "udape_rewarp_decode_select: misaligned pointer (the planes need 16-byte alignment)");
    // the fused route is the second-generation wide kernel: planes of exactly 4096 pixels, rows a power-of-two number of
    // 16-byte chunks, at most kW2DecodeMaxC planes per sample.  Anything else: udape_rewarp_fwd + udape_decode_select.
    const int64_t cpr = W * es / 16;
    UDAPE_REQUIRE(H * W == kW2Pixels && (W * es) % 16 == 0 && cpr >= 1 && (cpr & (cpr - 1)) == 0 && C <= kW2DecodeMaxC,
                  UDAPE_ERR_SHAPE, "udape_rewarp_decode_select: needs planes of 4096 pixels with power-of-two rows and C <= %d "
                  "(got C=%lld H=%lld W=%lld): call udape_rewarp_fwd and udape_decode_select", kW2DecodeMaxC, (long long)C,
                  (long long)H, (long long)W);
    UDAPE_REQUIRE(kth >= 0 && kth <= B * C, UDAPE_ERR_ARG, "udape_rewarp_decode_select: kth=%lld outside [0,%lld] (0 = no select)",
                  (long long)kth, (long long)(B * C));
    UDAPE_REQUIRE(kth == 0 || ticket, UDAPE_ERR_NULL, "udape_rewarp_decode_select: the select needs a ticket");
    a.view[0].in = in;
    a.view[0].theta = theta;
    RewarpDecodeOut dec = {};
    dec.idx = idx; dec.preds = preds; dec.maxvals_f32 = maxvals_f32; dec.position = position; dec.conf_table = conf_table;
    dec.occlude_thresh = occlude_thresh;
    dec.sel = SelectArgs{static_cast<int>(kth), tea_mask_in, thresh_out, tea_mask_out, kth > 0 ? ticket : nullptr, 0};
    cudaStream_t st = as_stream(stream);
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        constexpr int G = sizeof(T) == 2 ? 4 : 2;
        constexpr int RING = 3;
        const size_t smem = static_cast<size_t>(RING) * G * (kW2Pixels * sizeof(T) + 16) + sizeof(uint16_t) * kW2Pixels;
        const int r2 = reserve_smem(rewarp_wide2_kernel<T, G, RING, true>, smem, "udape_rewarp_decode_select");
        if (r2) return r2;
        a.tile_log2 = wide2_tile_log2(W, es);
        rewarp_wide2_kernel<T, G, RING, true><<<static_cast<unsigned>(B), kW2Threads, smem, st>>>(a, nullptr, dec);
    });
    return check_launch("udape_rewarp_decode_select");
}

extern "C" int udape_rewarp_bwd(const void* grad_out, const float* theta, int stages, int half_mask, int grid_dtype,
                                int64_t B, int64_t C, int64_t H, int64_t W, int dtype, void* grad_in,
                                const uint16_t* inverse_plan, void* stream) {
    RewarpArgs a = {};
    UDAPE_REQUIRE(grad_out && theta && grad_in, UDAPE_ERR_NULL, "udape_rewarp_bwd: NULL pointer");
    const int rc = fill_common(a, "udape_rewarp_bwd", 1, stages, half_mask, grid_dtype, B, C, H, W, dtype);
    if (rc) return rc;
    const int es = dtype_size(dtype);
    UDAPE_REQUIRE(aligned_to(grad_out, es) && aligned_to(grad_in, es) && aligned_to(theta, 4), UDAPE_ERR_ALIGN,
                  "udape_rewarp_bwd: misaligned pointer");
    UDAPE_REQUIRE(grad_out != grad_in, UDAPE_ERR_ARG, "udape_rewarp_bwd: cannot run in place");
    const int64_t hw = H * W;
    // the inverted map lives in shared memory: 6 bytes per pixel, 16-bit pixel indices
    UDAPE_REQUIRE(hw <= 25600, UDAPE_ERR_SHAPE, "udape_rewarp_bwd: planes above 25600 pixels are not supported (H*W=%lld)",
                  (long long)hw);
    a.view[0].in = grad_out;
    a.view[0].theta = theta;
    cudaStream_t st = as_stream(stream);
    // 16-bit arrays: smem route off[hw + 8] | lst_loc[hw] | map[hw]; general route off[hw + 2] | lst[hw] | map[hw]
    const size_t list_bytes = sizeof(uint16_t) * (((hw + 8 + 7) & ~7ll) + 2 * ((hw + 7) & ~7ll));
    const int buf_words = (aligned16(grad_out) && aligned16(grad_in) && !route_disabled()) ? smem_route_words(H, W, es) : 0;
    if (inverse_plan) {
        const int acc_words = rank_route_words(H, W, es);
        UDAPE_REQUIRE(acc_words != 0 && aligned16(grad_out) && aligned16(grad_in) && aligned16(inverse_plan),
                      UDAPE_ERR_ARG, "udape_rewarp_bwd: the inverse plan does not apply to this plane / alignment");
        // a pass = 4 planes as two packed pairs (2-byte types) or 2 planes (float32): 2 x (17 KB accumulators + 16 KB
        // tail) of shared memory per CTA (+ 32 KB of staged planes for the 2-byte types); one wave of CTAs, every CTA
        // an equal share of the passes
        const int planes_per_pass = es == 2 ? 4 : 2;
        int nt = 512;
        if (const char* e = std::getenv("UDAPE_REWARP_BWD_NT")) { const int v = std::atoi(e); if (v == 256 || v == 512) nt = v; }
        if (es != 2 || hw != kRwThreads * kRwPix) nt = 256;
        int per_sm = es == 2 ? 2 : 3;
        if (const char* e = std::getenv("UDAPE_REWARP_BWD_PER_SM")) { const int v = std::atoi(e); if (v >= 1 && v <= 3) per_sm = v; }
        const int64_t npass = B * ((C + planes_per_pass - 1) / planes_per_pass);
        int64_t grid = per_sm * static_cast<int64_t>(sm_count());
        if (grid > npass) grid = npass;
        UDAPE_REQUIRE(npass < (1ll << 31), UDAPE_ERR_SHAPE, "udape_rewarp_bwd: too many planes");
        const int64_t hw8 = (hw + 7) & ~7ll;
        const size_t smem = 2 * static_cast<size_t>(push_pitch_bytes(acc_words, static_cast<int>(hw))) + (es == 2 ? 4 * hw8 * 2 + 16 : 0);
        UDAPE_DISPATCH_FLOAT(dtype, T, {
            auto go = [&](auto kernel, int threads) -> int {
                const int r2 = reserve_smem(kernel, smem, "udape_rewarp_bwd");
                if (r2) return r2;
                kernel<<<static_cast<unsigned>(grid), threads, smem, st>>>(a, static_cast<const T*>(grad_out), static_cast<T*>(grad_in),
                                                                           acc_words, inverse_plan);
                return UDAPE_OK;
            };
            int r2;
            if constexpr (sizeof(T) == 2) {
                if (hw != kRwThreads * kRwPix) r2 = go(rewarp_bwd_push2_kernel<T, 2, false, 256, 2>, 256);
                else if (nt == 512) r2 = go(rewarp_bwd_push2_kernel<T, 2, true, 512, 2>, 512);
                else r2 = go(rewarp_bwd_push2_kernel<T, 2, true, 256, 2>, 256);
            } else {
                r2 = go(rewarp_bwd_push_kernel<T, 2>, 256);
            }
            if (r2) return r2;
        });
        return check_launch("udape_rewarp_bwd");
    }
    UDAPE_DISPATCH_FLOAT(dtype, T, {
        if (buf_words) {
            const int n = cluster_size_for(B, C, hw);
            a.cpc = static_cast<int>((C + n - 1) / n);
            const size_t smem = kRwRing * sizeof(uint32_t) * static_cast<size_t>(buf_words) + list_bytes;
            const int r2 = launch_cluster(rewarp_bwd_smem_kernel<T>, static_cast<unsigned>(B * n), static_cast<unsigned>(n), smem, st,
                                          "udape_rewarp_bwd", a, static_cast<const T*>(grad_out), static_cast<T*>(grad_in), buf_words);
            if (r2) return r2;
        } else {
            a.cpc = channels_per_cta(B, C, 2 * static_cast<int64_t>(sm_count()));
            const int64_t grid = B * ((C + a.cpc - 1) / a.cpc);
            const size_t smem = list_bytes;
            const int r2 = reserve_smem(rewarp_bwd_kernel<T>, smem, "udape_rewarp_bwd");
            if (r2) return r2;
            rewarp_bwd_kernel<T><<<static_cast<unsigned>(grid), kRwThreads, smem, st>>>(
                a, static_cast<const T*>(grad_out), static_cast<T*>(grad_in));
        }
    });
    return check_launch("udape_rewarp_bwd");
}
