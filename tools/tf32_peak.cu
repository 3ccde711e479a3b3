// Measures the chip's dense TF32 tensor-core peak the way MEASURED_PEAKS.json measures bf16 (which has no TF32 entry):
// back-to-back tcgen05.mma kind::tf32 (M128 x N256 x K8, SS form, operands resident in shared memory, pseudo-random
// data so the datapath toggles), one CTA per SM, timed with CUDA events.
//   burst     = best of 10 launches of ~2 ms separated by idle gaps (the power controller has not reacted yet)
//   sustained = all launches of a 4 s back-to-back loop (clock settles where the power cap puts it)
// Prints ONE JSON line; bench.py runs this binary on rank 0 before its own warm-up and uses the result as the roofline
// denominator (burst or sustained according to the SM clock it samples during its own timed region).
// Not product, not a test: build with  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tf32_peak tools/tf32_peak.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../neuro__b200/csrc/sm100_ptx.cuh"
using namespace nb200;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int kN = 256;

__global__ void __launch_bounds__(128, 1) peak_kernel(int reps)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bar = (uint64_t*)(smem + 96 * 1024);
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5;
    // operands: A 128 x 32 (16 KB) at 0, B 256 x 32 (32 KB) at 32 KB; values in (-1, 1) from a hash, TF32-representable
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x)
    {
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const float v = ((int)(h & 0xFFFF) - 32768) * (1.0f / 32768.0f);
        ((float*)smem)[i] = v;
    }
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before_sync(); __syncthreads(); ptx::tc_fence_after_sync();
    const uint32_t tm = *slot;
    if (warp == 0)
    {
        const uint32_t idesc = ptx::idesc_tf32(128, kN, 0, 0);
        const uint64_t da = ptx::smem_desc_sw128(ptx::smem_u32(smem), 16, 1024);
        const uint64_t db = ptx::smem_desc_sw128(ptx::smem_u32(smem + 32 * 1024), 16, 1024);
        if (ptx::elect_one())
        {
            for (int r = 0; r < reps; ++r)
            {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)   // two accumulators alternate so consecutive MMAs are independent
                    ptx::mma_tf32_ss(tm + (r & 1) * 256, da + kk * 2, db + kk * 2, idesc, (r > 1 || kk > 0) ? 1 : 0);
            }
            ptx::mma_commit(bar);
        }
        __syncwarp();
        ptx::mbar_wait(bar, 0);
    }
    ptx::tc_fence_before_sync(); __syncthreads();
    if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tm, 512); }
}

int main(int argc, char** argv)
{
    const double sustainSeconds = argc > 1 ? atof(argv[1]) : 4.0;
    int dev = 0, sms = 0, clk = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    const int smemBytes = 100 * 1024;
    CK(cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
    const int reps = 16384;                                              // 65536 MMAs of 128 cycles: ~4 ms at 1.9 GHz
    const double flopsPerLaunch = 2.0 * 128 * kN * 8 * 4.0 * reps * sms;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    peak_kernel<<<sms, 128, smemBytes>>>(reps); CK(cudaDeviceSynchronize());    // warm-up
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
    double burst = 0;
    for (int i = 0; i < 10; ++i)
    {
        CK(cudaEventRecord(e0)); peak_kernel<<<sms, 128, smemBytes>>>(reps / 2); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = flopsPerLaunch / 2 / (ms * 1e-3) / 1e12;
        if (tf > burst) burst = tf;
        std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
    // sustained: back-to-back launches until the wall clock says enough; the LAST second is what the clock has settled to
    const auto t0 = std::chrono::steady_clock::now();
    double total = 0, last = 0; int launches = 0;
    CK(cudaEventRecord(e0));
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < sustainSeconds)
    {
        for (int i = 0; i < 16; ++i) peak_kernel<<<sms, 128, smemBytes>>>(reps);
        launches += 16;
        CK(cudaStreamSynchronize(0));
    }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float msAll = 0; CK(cudaEventElapsedTime(&msAll, e0, e1));
    total = flopsPerLaunch * launches / (msAll * 1e-3) / 1e12;
    {
        const int n = 64;
        CK(cudaEventRecord(e0));
        for (int i = 0; i < n; ++i) peak_kernel<<<sms, 128, smemBytes>>>(reps);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        last = flopsPerLaunch * n / (ms * 1e-3) / 1e12;
    }
    printf("{\"tf32_tflops_burst\": %.1f, \"tf32_tflops_sustained\": %.1f, \"tf32_tflops_settled\": %.1f, \"sms\": %d, \"sm_max_khz\": %d, "
           "\"nominal_at_max_clock\": %.1f, \"how\": \"tcgen05.mma kind::tf32 M128xN256xK8 SS form, 1 CTA/SM, pseudo-random operands; burst = best of 10 x ~2 ms with idle gaps, "
           "sustained = %.1f s back to back, settled = the 64 launches after that\"}\n",
           burst, total, last, sms, clk, 2.0 * 2048 * sms * (double)clk * 1e3 / 1e12, sustainSeconds);
    return 0;
}
