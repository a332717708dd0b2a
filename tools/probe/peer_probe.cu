// peer_probe.cu — measurement tool (not part of the library): how fast can SMs pull another GPU's memory over
// NVLink with (0) 128-bit loads, (1) cp.async.bulk (1-D TMA) into shared memory?  Both write what they pulled
// to local memory.  Built by tools/probe/build.sh into build/peer_probe.so; driven by tools/probe/peer_probe.py.
#include <cuda_runtime.h>
#include <cstdint>

#include "../../uda_poseestimation_b200/csrc/pipeline.cuh"

using namespace udape;

__device__ __forceinline__ uint4 ld_weak(const void* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

template <int UN>
__global__ void __launch_bounds__(256) pull_ldg(const uint4* __restrict__ src, uint4* __restrict__ dst, long long nvec) {
    const long long per_cta = 256ll * UN;
    for (long long base = blockIdx.x * per_cta; base < nvec; base += gridDim.x * per_cta) {
        uint4 x[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < nvec) x[u] = ld_weak(src + i);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < nvec) dst[i] = x[u];
        }
    }
}

// one producer thread per CTA issues bulk loads of `chunk` bytes into a ring of STAGES; all threads copy a landed
// chunk to local memory with 128-bit stores
template <int STAGES>
__global__ void __launch_bounds__(256) pull_tma(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long bytes, int chunk) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[STAGES];
    const long long nchunks = bytes / chunk;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        mbar_init_fence();
    }
    __syncthreads();
    // chunks of this CTA: blockIdx.x, + gridDim.x, ...
    const long long mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto issue = [&](long long k) {
        const int s = static_cast<int>(k % STAGES);
        mbar_expect_tx(&full[s], static_cast<uint32_t>(chunk));
        bulk_load(smem + static_cast<size_t>(s) * chunk, src + (blockIdx.x + k * gridDim.x) * static_cast<long long>(chunk), static_cast<uint32_t>(chunk), &full[s]);
    };
    if (threadIdx.x == 0)
        for (long long k = 0; k < STAGES - 1 && k < mine; ++k) issue(k);
    for (long long k = 0; k < mine; ++k) {
        const int s = static_cast<int>(k % STAGES);
        __syncthreads();   // everybody is done with the stage that chunk k + STAGES - 1 will overwrite
        if (threadIdx.x == 0 && k + STAGES - 1 < mine) issue(k + STAGES - 1);
        mbar_wait(&full[s], static_cast<uint32_t>((k / STAGES) & 1));
        const uint4* in = reinterpret_cast<const uint4*>(smem + static_cast<size_t>(s) * chunk);
        uint4* out = reinterpret_cast<uint4*>(dst + (blockIdx.x + k * gridDim.x) * static_cast<long long>(chunk));
        for (int i = threadIdx.x; i < chunk / 16; i += 256) out[i] = in[i];
    }
}

// push: read local memory, store to the peer (128-bit stores), or stage through shared memory and bulk-store
template <int UN>
__global__ void __launch_bounds__(256) push_stg(const uint4* __restrict__ src, uint4* __restrict__ dst, long long nvec) {
    const long long per_cta = 256ll * UN;
    for (long long base = blockIdx.x * per_cta; base < nvec; base += gridDim.x * per_cta) {
        uint4 x[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < nvec) x[u] = __ldcs(src + i);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const long long i = base + u * 256 + threadIdx.x;
            if (i < nvec) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i), "r"(x[u].x), "r"(x[u].y), "r"(x[u].z), "r"(x[u].w) : "memory");
        }
    }
}

__global__ void __launch_bounds__(256) push_tma(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long bytes, int chunk) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[2];
    const long long nchunks = bytes / chunk;
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init_fence();
    }
    __syncthreads();
    const long long mine = (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x;
    if (threadIdx.x == 0) {   // one thread drives both directions: load chunk k into stage k % 2, store it out, keep 1 store group in flight
        for (long long k = 0; k < mine; ++k) {
            const int s = static_cast<int>(k & 1);
            if (k >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that read stage s has finished reading it
            const long long off = (blockIdx.x + k * gridDim.x) * static_cast<long long>(chunk);
            mbar_expect_tx(&full[s], static_cast<uint32_t>(chunk));
            bulk_load(smem + static_cast<size_t>(s) * chunk, src + off, static_cast<uint32_t>(chunk), &full[s]);
            mbar_wait(&full[s], static_cast<uint32_t>((k >> 1) & 1));
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(smem_addr(smem + static_cast<size_t>(s) * chunk)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

extern "C" int probe_push(const void* src, void* dst, long long bytes, int mode, int grid, int param, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == 0) {
        if (param == 8) push_stg<8><<<grid, 256, 0, st>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), bytes / 16);
        else push_stg<4><<<grid, 256, 0, st>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), bytes / 16);
    } else {
        const size_t smem = 2 * static_cast<size_t>(param);
        cudaFuncSetAttribute(push_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        push_tma<<<grid, 256, smem, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), bytes, param);
    }
    return static_cast<int>(cudaGetLastError());
}

extern "C" int probe_pull(const void* src, void* dst, long long bytes, int mode, int grid, int param, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == 0) {
        if (param == 8) pull_ldg<8><<<grid, 256, 0, st>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), bytes / 16);
        else if (param == 4) pull_ldg<4><<<grid, 256, 0, st>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), bytes / 16);
        else pull_ldg<2><<<grid, 256, 0, st>>>(static_cast<const uint4*>(src), static_cast<uint4*>(dst), bytes / 16);
    } else {
        const int chunk = param;   // bytes per bulk copy
        const size_t smem = 4 * static_cast<size_t>(chunk);
        cudaFuncSetAttribute(pull_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        pull_tma<4><<<grid, 256, smem, st>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), bytes, chunk);
    }
    return static_cast<int>(cudaGetLastError());
}
