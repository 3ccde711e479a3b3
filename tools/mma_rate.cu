// Bring-up microbenchmark: back-to-back tcgen05.mma issue rate per instruction, TS (A in TMEM) vs SS (A in smem),
// for several N. One CTA per SM (grid 148) so the power/clock state is realistic. Not product, not a test.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../neuro__b200/csrc/sm100_ptx.cuh"
using namespace nb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int N, int TS>
__global__ void rate_kernel(int reps, long long* out, int mode, const float* gsrc, volatile int* sink)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 96 * 1024);
    uint32_t* slot = (uint32_t*)(bar + 1);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar + 2, 1); ptx::mbar_init(bar + 3, 1); ptx::fence_mbar_init(); *(volatile int*)(slot + 1) = 0; }
    if (warp == 0) ptx::tmem_alloc(slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before_sync(); __syncthreads(); ptx::tc_fence_after_sync();
    const uint32_t tm = *slot;
    long long dt = 0;
    if (warp == 0)
    {
        const uint32_t idesc = ptx::idesc_tf32(128, N, 0, 0);
        const uint64_t da = ptx::smem_desc_sw128(ptx::smem_u32(smem), 16, 1024);
        const uint64_t db = ptx::smem_desc_sw128(ptx::smem_u32(smem + 32 * 1024), 16, 1024);
        const long long t0 = clock64();
        if (ptx::elect_one())
        {
            for (int r = 0; r < reps; ++r)
            {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                {
                    if (TS) ptx::mma_tf32_ts(tm, tm + 256 + kk * 8, db + kk * 2, idesc, 1);
                    else ptx::mma_tf32_ss(tm, da + kk * 2, db + kk * 2, idesc, 1);
                }
                if (mode & 8) ptx::mma_commit(bar + 2);
                if (mode & 16) ptx::mma_commit(bar + 3);
            }
            ptx::mma_commit(bar);
        }
        __syncwarp();
        ptx::mbar_wait(bar, 0);
        dt = clock64() - t0;
    }
    else if (warp >= 4 && warp < 11 && (mode & 1))
    {
        // background: conflict-free LDS stream (like the converters' halo reads) until the MMA warp is done
        volatile int* flag = (volatile int*)(slot + 1);
        const float* src = (const float*)(smem + 64 * 1024) + (threadIdx.x & 31);
        float acc = 0.f;
        while (*flag == 0)
        {
#pragma unroll
            for (int c = 0; c < 32; ++c) acc += src[c * 240];
        }
        if (acc == 123.456f) *sink = 1;
    }
    else if (warp >= 1 && warp < 4 && (mode & 2))
    {
        // background: TMEM stores into unused columns (like the converters' A tiles)
        volatile int* flag = (volatile int*)(slot + 1);
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = c;
        while (*flag == 0)
        {
            ptx::tmem_st_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + 320, v);
            ptx::tmem_st_wait();
        }
    }
    else if (warp == 4 + 7 && (mode & 4))
    {
        // background: bulk-copy 32 KB chunks global -> smem (like the filter TMA ring), 2 in flight
        volatile int* flag = (volatile int*)(slot + 1);
        uint64_t* tb = bar + 8;
        if ((threadIdx.x & 31) == 0)
        {
            ptx::mbar_init(&tb[0], 1); ptx::mbar_init(&tb[1], 1); ptx::fence_mbar_init();
            uint32_t ph[2] = {0, 0};
            int i = 0;
            for (int k = 0; k < 2; ++k)
            {
                ptx::mbar_arrive_expect_tx(&tb[k], 16384);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ptx::smem_u32(smem + 64 * 1024 + k * 16384)), "l"(gsrc + (size_t)blockIdx.x * 65536 + k * 4096), "r"(16384), "r"(ptx::smem_u32(&tb[k])) : "memory");
            }
            while (*flag == 0)
            {
                ptx::mbar_wait(&tb[i], ph[i]); ph[i] ^= 1;
                ptx::mbar_arrive_expect_tx(&tb[i], 16384);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(ptx::smem_u32(smem + 64 * 1024 + i * 16384)), "l"(gsrc + (size_t)blockIdx.x * 65536 + i * 4096), "r"(16384), "r"(ptx::smem_u32(&tb[i])) : "memory");
                i ^= 1;
                atomicAdd((int*)sink + 1, 1);
            }
            ptx::mbar_wait(&tb[0], ph[0]); ptx::mbar_wait(&tb[1], ph[1]);
        }
    }
    if (warp == 0) { *(volatile int*)(slot + 1) = 1; }
    ptx::tc_fence_before_sync(); __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = dt;
    if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tm, 512); }
}

template <int N, int TS>
int run(const char* name, int grid, int mode)
{
    const int reps = 2000, smemBytes = 100 * 1024;
    long long* d; CK(cudaMalloc(&d, grid * sizeof(long long)));
    int* sinkd; CK(cudaMalloc(&sinkd, 8));
    static float* gsrc = nullptr; if (!gsrc) { CK(cudaMalloc(&gsrc, 148 * 65536 * 4)); CK(cudaMemset(gsrc, 0, 148 * 65536 * 4)); }
    CK(cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
    CK(cudaMemset(sinkd, 0, 8));
    rate_kernel<N, TS><<<grid, 384, smemBytes>>>(reps, d, mode, gsrc, sinkd);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(sinkd, 0, 8));
    rate_kernel<N, TS><<<grid, 384, smemBytes>>>(reps, d, mode, gsrc, sinkd);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
    const double perMma = avg / (reps * 4.0);
    printf("%-3s bg=%d N=%3d grid=%3d: %.1f cycles per MMA (M128 x N x K8) -> %.0f MAC/cycle/SM (peak model 2048)\n", name, mode, N, grid, perMma, 128.0 * N * 8 / perMma);
    int hs[2]; cudaMemcpy(hs, sinkd, 8, cudaMemcpyDeviceToHost); if (mode & 4) printf("      background bulk copies per CTA: %.0f (x16 KB) over %.0f cycles -> %.1f B/cycle/SM\n", hs[1] / (double)grid, avg, hs[1] / (double)grid * 16384 / avg);
    cudaFree(d);
    return 0;
}

int main()
{
    for (int mode : {0, 8, 24})
    {
        run<128, 1>("TS", 148, mode); run<256, 1>("TS", 148, mode); run<256, 0>("SS", 148, mode);
    }
    return 0;
}
