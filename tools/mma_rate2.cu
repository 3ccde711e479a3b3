// Bring-up microbenchmark #2: (a) cta_group::2 MMA issue rate, (b) multicast bulk-copy delivery rate per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../neuro__b200/csrc/sm100_ptx.cuh"
using namespace nb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int N>
__global__ void pair_rate_kernel(int reps, long long* out)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 96 * 1024);
    uint32_t* slot = (uint32_t*)(bar + 1);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = ptx::cluster_ctarank();
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc_2sm(slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before_sync(); ptx::cluster_sync(); ptx::tc_fence_after_sync();
    const uint32_t tm = *slot;
    long long dt = 0;
    if (warp == 0)
    {
        const long long t0 = clock64();
        if (rank == 0)
        {
            const uint32_t idesc = ptx::idesc_tf32(256, N, 0, 0);
            const uint64_t db = ptx::smem_desc_sw128(ptx::smem_u32(smem + 32 * 1024), 16, 1024);
            if (ptx::elect_one())
            {
                for (int r = 0; r < reps; ++r)
                {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::mma_tf32_ts_2sm(tm, tm + 256 + kk * 8, db + kk * 2, idesc, 1);
                }
                ptx::mma_commit_2sm(bar, 3);
            }
            __syncwarp();
        }
        ptx::mbar_wait(bar, 0);
        dt = clock64() - t0;
    }
    ptx::tc_fence_before_sync(); ptx::cluster_sync();
    if (threadIdx.x == 0) out[blockIdx.x] = dt;
    if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc_2sm(tm, 512); }
}

// multicast delivery: every CTA of a cluster of CS issues 1/CS of the chunks, each multicast to all CS CTAs.
template <int CS>
__global__ void mcast_kernel(int chunks, const float* gsrc, long long* out)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 64 * 1024); // [4] ring of 16 KB slots
    const uint32_t rank = ptx::cluster_ctarank();
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(&bar[i], 1); ptx::fence_mbar_init(); }
    ptx::cluster_sync();
    long long dt = 0;
    if (threadIdx.x == 0)
    {
        const long long t0 = clock64();
        uint32_t ph[4] = {0, 0, 0, 0};
        // chunk j lands in slot j % 4 of EVERY CTA; CTA (j % CS) issues it. Each CTA expects 16 KB per chunk on its own barrier.
        // (No consumer release protocol: slots are simply overwritten; we only measure delivery rate.)
        const float* src = gsrc + (size_t)(blockIdx.x / CS) * 65536;
        for (int j = 0; j < chunks + 4; ++j)
        {
            const int s = j & 3;
            if (j >= 4) { ptx::mbar_wait(&bar[s], ph[s]); ph[s] ^= 1; }
            if (j < chunks)
            {
                ptx::mbar_arrive_expect_tx(&bar[s], 16384);
                if ((j % CS) == (int)rank)
                {
                    if (CS == 1)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(ptx::smem_u32(smem + s * 16384)), "l"(src + (j & 3) * 4096), "r"(16384), "r"(ptx::smem_u32(&bar[s])) : "memory");
                    else
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                                     ::"r"(ptx::smem_u32(smem + s * 16384)), "l"(src + (j & 3) * 4096), "r"(16384), "r"(ptx::smem_u32(&bar[s])), "h"((uint16_t)((1 << CS) - 1)) : "memory");
                }
            }
        }
        dt = clock64() - t0;
    }
    ptx::cluster_sync();
    if (threadIdx.x == 0) out[blockIdx.x] = dt;
}

template <int N> int run_pair()
{
    const int reps = 2000, grid = 148, smemBytes = 100 * 1024;
    long long* d; CK(cudaMalloc(&d, grid * sizeof(long long)));
    CK(cudaFuncSetAttribute(pair_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smemBytes;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = 2; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    for (int it = 0; it < 2; ++it) { CK(cudaLaunchKernelEx(&cfg, pair_rate_kernel<N>, reps, d)); CK(cudaDeviceSynchronize()); }
    long long h[148]; CK(cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    printf("2CTA TS M=256 N=%3d: %.1f cycles per MMA -> %.0f MAC/cycle/SM\n", N, avg / (reps * 4.0), 128.0 * N * 8 / (avg / (reps * 4.0)));
    cudaFree(d); return 0;
}

template <int CS> int run_mcast(const float* gsrc)
{
    const int chunks = 4000, grid = 148 / CS * CS, smemBytes = 70 * 1024;
    long long* d; CK(cudaMalloc(&d, 148 * sizeof(long long)));
    CK(cudaFuncSetAttribute(mcast_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smemBytes;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = CS; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    for (int it = 0; it < 2; ++it) { CK(cudaLaunchKernelEx(&cfg, mcast_kernel<CS>, chunks, gsrc, d)); CK(cudaDeviceSynchronize()); }
    long long h[148]; CK(cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    printf("bulk copy cluster=%d grid=%d: %.1f B/cycle DELIVERED per SM (%.0f cycles per 16 KB chunk)\n", CS, grid, chunks * 16384.0 / avg, avg / chunks);
    cudaFree(d); return 0;
}

int main()
{
    run_pair<64>(); run_pair<128>(); run_pair<256>();
    float* gsrc; CK(cudaMalloc(&gsrc, 148 * 65536 * 4)); CK(cudaMemset(gsrc, 0, 148 * 65536 * 4));
    run_mcast<1>(gsrc); run_mcast<2>(gsrc); run_mcast<4>(gsrc);
    return 0;
}
