// Bring-up tool (not product, not a test): one CTA does TMA(A,B) -> dump smem -> 4x tcgen05.mma -> dump TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/tc_debug tools/tc_debug.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../neuro__b200/csrc/sm100_ptx.cuh"
using namespace nb200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int BN = 64;
constexpr uint32_t kABytes = 4 * 32 * 32 * 4, kBBytes = BN * 32 * 4;

__global__ void dbg_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, int cw, int ch,
                           int doMma, int aMajorMN, uint32_t lboA, uint32_t sboA, int mode, float* dumpA, float* dumpB, float* dumpD)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a = smem; uint8_t* b = smem + kABytes;
    uint64_t* bar = (uint64_t*)(smem + kABytes + kBBytes);
    uint64_t* bar2 = bar + 1;
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar2, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(slot, 2 * BN);
    ptx::tc_fence_before_sync(); __syncthreads(); ptx::tc_fence_after_sync();
    const uint32_t tm = *slot;
    if (threadIdx.x == 0)
    {
        ptx::mbar_arrive_expect_tx(bar, kABytes + kBBytes);
        ptx::tma_load_4d(a, &mapX, bar, cw, 0, ch, 0);
        ptx::tma_load_3d(b, &mapW, bar, 0, 0, 0);
    }
    ptx::mbar_wait(bar, 0);
    __syncthreads();
    for (int i = threadIdx.x; i < kABytes / 4; i += blockDim.x) dumpA[i] = ((float*)a)[i];
    for (int i = threadIdx.x; i < kBBytes / 4; i += blockDim.x) dumpB[i] = ((float*)b)[i];
    __syncthreads();
    if (doMma)
    {
        if (mode == 2)
        {
            // TS: thread = pixel row m; gather its 32 channels from smem [h][c][w] (128B-swizzled) and store to TMEM cols [BN, BN+32)
            const int m = threadIdx.x, h = m >> 5, wv = m & 31;
            uint32_t v[32];
            for (int c = 0; c < 32; ++c)
            {
                uint32_t off = h * 4096 + c * 128 + wv * 4;
                off ^= ((off >> 7) & 7) << 4;
                v[c] = *(const uint32_t*)(a + off);
            }
            ptx::tmem_st_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + BN, v);
            ptx::tmem_st_wait();
            ptx::tc_fence_before_sync();
            __syncthreads();
        }
        if (threadIdx.x == 0)
        {
            ptx::tc_fence_after_sync();
            const uint32_t idesc = ptx::idesc_tf32(128, BN, mode == 2 ? 0 : aMajorMN, 0);
            for (int kk = 0; kk < 4; ++kk)
            {
                uint64_t da = ptx::smem_desc_sw128(ptx::smem_u32(a) + kk * 1024, lboA, sboA);
                if (mode == 1) da = (da & ~((uint64_t)7 << 61)) | ((uint64_t)1 << 61);
                const uint64_t db = ptx::smem_desc_sw128(ptx::smem_u32(b) + kk * 32, 16, 1024);
                if (mode == 2) ptx::mma_tf32_ts(tm, tm + BN + kk * 8, db, idesc, kk != 0);
                else ptx::mma_tf32_ss(tm, da, db, idesc, kk != 0);
            }
            ptx::mma_commit(bar2);
        }
        ptx::mbar_wait(bar2, 0);
        ptx::tc_fence_after_sync();
        if (warp < 4)
        {
            for (int c0 = 0; c0 < BN; c0 += 32)
            {
                uint32_t v[32];
                ptx::tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
                ptx::tmem_ld_wait();
                for (int j = 0; j < 32; ++j) dumpD[(warp * 32 + lane) * BN + c0 + j] = __uint_as_float(v[j]);
            }
        }
    }
    ptx::tc_fence_before_sync(); __syncthreads();
    if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tm, 2 * BN); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7) << 4); }
static float tf32r(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x1000; u &= ~0x1FFFu; memcpy(&f, &u, 4); return f; }

struct Exp { int cw, ch, mma, mn; uint32_t lbo, sbo; int mode; };
int main(int argc, char** argv)
{
    int C = 32, H = 8, W = 32, K = 64;
    Exp one = {0, 0, 0, 1, 4096, 1024, 0};
    if (argc > 6) { one.cw = atoi(argv[1]); one.ch = atoi(argv[2]); one.mma = atoi(argv[3]); one.mn = atoi(argv[4]); one.lbo = atoi(argv[5]); one.sbo = atoi(argv[6]); }
    if (argc > 7) one.mode = atoi(argv[7]);
    if (argc > 10) { C = atoi(argv[8]); H = atoi(argv[9]); W = atoi(argv[10]); }
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    std::vector<float> x((size_t)C * H * W), w((size_t)K * 32, 0.f);
    for (size_t i = 0; i < x.size(); ++i) x[i] = tf32r((float)((i * 2654435761u) % 1000) / 500.f - 1.f);
    for (int k = 0; k < K; ++k) for (int c = 0; c < 32; ++c) w[k * 32 + c] = c < C ? tf32r((float)(((k * 32 + c) * 40503u) % 1000) / 500.f - 1.f) : 0.f;
    float *dx, *dw, *dA, *dB, *dD;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dw, w.size() * 4));
    CK(cudaMalloc(&dA, kABytes)); CK(cudaMalloc(&dB, kBBytes)); CK(cudaMalloc(&dD, 128 * BN * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    CUtensorMap mx, mw; cuuint32_t es[4] = {1, 1, 1, 1};
    { cuuint64_t d[4] = {(cuuint64_t)W, (cuuint64_t)C, (cuuint64_t)H, 1}; cuuint64_t s[3] = {(cuuint64_t)H * W * 4, (cuuint64_t)W * 4, (cuuint64_t)C * H * W * 4}; cuuint32_t b[4] = {32, 32, 4, 1};
      CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, one.mode == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); printf("encode X -> %d\n", (int)r); if (r) return 1; }
    { cuuint64_t d[3] = {32, (cuuint64_t)K, 1}; cuuint64_t s[2] = {32 * 4, (cuuint64_t)32 * K * 4}; cuuint32_t b[3] = {32, BN, 1};
      CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dw, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); printf("encode W -> %d\n", (int)r); if (r) return 1; }
    const int smemBytes = kABytes + kBBytes + 1024 + 64;
    CK(cudaFuncSetAttribute(dbg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
    std::vector<float> hA(kABytes / 4), hB(kBBytes / 4), hD(128 * BN);
    Exp exps[] = { one };
    for (const Exp& e : exps)
    {
        CK(cudaMemset(dA, 0xFF, kABytes)); CK(cudaMemset(dB, 0xFF, kBBytes)); CK(cudaMemset(dD, 0xFF, 128 * BN * 4));
        dbg_kernel<<<1, 128, smemBytes>>>(mx, mw, e.cw, e.ch, e.mma, e.mn, e.lbo, e.sbo, e.mode, dA, dB, dD);
        cudaError_t err = cudaDeviceSynchronize();
        printf("exp cw=%d ch=%d mma=%d mn=%d lbo=%u sbo=%u -> %s\n", e.cw, e.ch, e.mma, e.mn, e.lbo, e.sbo, cudaGetErrorString(err));
        if (err != cudaSuccess) return 1;
        CK(cudaMemcpy(hA.data(), dA, kABytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hB.data(), dB, kBBytes, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hD.data(), dD, 128 * BN * 4, cudaMemcpyDeviceToHost));
        // check A: smem[h][c][w] swizzled
        int badA = 0, badB = 0; 
        std::vector<float> Am(128 * 32), Bm(BN * 32);
        for (int h = 0; h < 4; ++h) for (int c = 0; c < 32; ++c) for (int ww = 0; ww < 32; ++ww)
        {
            int gh = e.ch + h, gw = e.cw + ww;
            float expect = (c < C && gh >= 0 && gh < H && gw >= 0 && gw < W) ? x[((size_t)c * H + gh) * W + gw] : 0.f;
            uint32_t o = h * 4096 + c * 128 + ww * 4; float got = hA[(e.mode == 1 ? (o ^ (((o >> 7) & 3) << 5)) : swz(o)) / 4];
            if (got != expect) { if (badA < 3) printf("  A mismatch h%d c%d w%d got %f expect %f\n", h, c, ww, got, expect); ++badA; }
            Am[(h * 32 + ww) * 32 + c] = expect;
        }
        for (int k = 0; k < BN; ++k) for (int c = 0; c < 32; ++c)
        {
            float expect = w[k * 32 + c]; float got = hB[swz(k * 128 + c * 4) / 4];
            if (got != expect) { if (badB < 3) printf("  B mismatch k%d c%d got %f expect %f\n", k, c, got, expect); ++badB; }
            Bm[k * 32 + c] = expect;
        }
        printf("  smem A mismatches %d, B mismatches %d\n", badA, badB);
        if (e.mma)
        {
            double maxerr = 0, maxref = 0; int nan = 0;
            for (int m = 0; m < 128; ++m) for (int n = 0; n < BN; ++n)
            {
                double ref = 0; for (int c = 0; c < 32; ++c) ref += (double)Am[m * 32 + c] * Bm[n * 32 + c];
                float got = hD[m * BN + n]; if (got != got) ++nan;
                maxerr = fmax(maxerr, fabs(got - ref)); maxref = fmax(maxref, fabs(ref));
            }
            printf("  D max err %.4g (max ref %.4g) nan %d   D[0][0..3] = %f %f %f %f\n", maxerr, maxref, nan, hD[0], hD[1], hD[2], hD[3]);
        }
    }
    return 0;
}
