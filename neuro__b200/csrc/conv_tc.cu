// tcgen05 / TMA implicit-GEMM convolution kernels for NCHW fp32 tensors (TF32 tensor-core math).
//
// Forward (and the stride-1 input gradient, which is a forward conv of dy with flipped, transposed filters):
//
//   GEMM view      D[M = pixels][N = filters] = sum over (tap, channel) A[pixel][channel@tap] * B[filter][channel@tap]
//   CTA tile       128 pixels (4 output rows x 32 output columns of one image) x BN filters
//   A operand      the NCHW activation tile itself: for one filter tap the 32 pixels of a row are CONTIGUOUS in W and the
//                  channels are strided, i.e. the tile is "MN-major". One 4-D TMA box {32 w, 32 c, 4 h, 1 n} per
//                  (tap, channel block) lands in shared memory as [h][c][w] with the 128-byte swizzle, which is exactly
//                  the canonical MN-major SWIZZLE_128B UMMA layout (row atoms at LBO = 4 KB, 8-channel groups at 1 KB).
//                  No im2col buffer exists anywhere: the tap shift is the TMA box origin, and zero padding is the
//                  TMA out-of-bounds fill (negative or past-the-end coordinates read as 0).
//   B operand      filters repacked once per call to [tap][filter][channel] (channel contiguous, rounded to TF32):
//                  K-major, one 3-D TMA box {32 c, BN k, 1 tap} per stage.
//   accumulator    128 lanes x BN columns of TMEM, fp32. One thread issues tcgen05.mma (M128 x BN x K8, 4 per stage).
//   epilogue       4 warps read TMEM with tcgen05.ld (warp q owns output row q of the tile: lane = output column, so each
//                  store instruction writes one full 128-byte line of y), fused bias + activation.
//   pipeline       warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue; smem ring of kStages
//                  {A 16 KB, B BN*128 B} guarded by full/empty mbarriers; tcgen05.commit releases ring slots and
//                  signals the epilogue. Two CTAs are co-resident per SM so one tile's epilogue overlaps the other's MMAs.
#include <cuda.h>
#include <mutex>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace nb200
{
    namespace
    {
        constexpr int kTileW = 32;   // pixels per swizzle atom row (128 B of fp32)
        constexpr int kTileH = 4;    // rows per CTA tile -> M = 128
        constexpr int kBlockC = 32;  // reduction channels per pipeline stage (4 MMAs of K = 8)
        constexpr int kThreads = 192;
        constexpr uint32_t kABytes = kTileH * kBlockC * kTileW * 4; // 16 KB

        template <int BN>
        struct FpropCfg
        {
            static constexpr uint32_t bBytes = BN * kBlockC * 4;
            static constexpr uint32_t stageBytes = kABytes + bBytes;
            // two CTAs per SM: stay under ~110 KB each
            static constexpr int stages = (BN <= 64) ? 4 : 3;
            static constexpr uint32_t smemBytes = stages * stageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
        };

        struct FpropParams
        {
            int Cblocks;      // ceil(C / 32)
            int R, S;
            int padX, padY;
            int Ho, Wo, K;    // output extent and filter count
            int tilesW, tilesH, tilesK;
            int act;
            float alpha;
            long long yStrideN, yStrideK; // elements
        };

        // ---------------------------------------------------------------- filter repack
        // out[tap][k][c] (c padded to Cp with zeros), TF32-rounded (round to nearest, ties away: cvt.rna).
        // mode 0 (forward):           tap = r*S+s,                 k = filter, c = channel   <- w[k][c][r][s]
        // mode 1 (input gradient):    tap = (R-1-r)*S + (S-1-s),   "k" = channel, "c" = filter <- w[f][ch][r][s]
        __global__ void repack_filters_kernel(const float* __restrict__ w, float* __restrict__ out, int K, int C, int R, int S,
                                              int outRows, int outCp, int mode)
        {
            const long long total = (long long)R * S * outRows * outCp;
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
            {
                const int col = (int)(i % outCp);
                const int row = (int)((i / outCp) % outRows);
                const int tap = (int)(i / ((long long)outCp * outRows));
                float v = 0.f;
                if (mode == 0)
                {
                    if (col < C)
                        v = w[((long long)row * C + col) * R * S + tap];
                }
                else
                {
                    const int r = R - 1 - tap / S, s = S - 1 - tap % S;
                    if (col < K)
                        v = w[(((long long)col * C + row) * R + r) * S + s];
                }
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
                out[i] = __uint_as_float(t);
            }
        }

        // ---------------------------------------------------------------- forward kernel
        template <int BN>
        __global__ void __launch_bounds__(kThreads, 2)
        tc_fprop_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, FpropParams p,
                        const float* __restrict__ bias, float* __restrict__ y)
        {
            using Cfg = FpropCfg<BN>;
            constexpr int kStages = Cfg::stages;

            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            uint8_t* ring = smem;
            uint64_t* fullBar = (uint64_t*)(smem + kStages * Cfg::stageBytes);
            uint64_t* emptyBar = fullBar + kStages;
            uint64_t* accBar = emptyBar + kStages;
            uint32_t* tmemSlot = (uint32_t*)(accBar + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;

            // tile coordinates: filter tile fastest so CTAs sharing the same activation tile run together (L2 reuse)
            int t = blockIdx.x;
            const int kt = t % p.tilesK; t /= p.tilesK;
            const int tw = t % p.tilesW; t /= p.tilesW;
            const int th = t % p.tilesH; t /= p.tilesH;
            const int n = t;
            const int ow0 = tw * kTileW, oh0 = th * kTileH, k0 = kt * BN;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapW);
                for (int s = 0; s < kStages; ++s)
                {
                    ptx::mbar_init(&fullBar[s], 1);
                    ptx::mbar_init(&emptyBar[s], 1);
                }
                ptx::mbar_init(accBar, 1);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, BN);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            const uint32_t tmemBase = *tmemSlot;

            const int taps = p.R * p.S;
            const int iters = taps * p.Cblocks;

            if (warp == 0)
            {
                if (lane == 0)
                {
                    // ===== TMA producer =====
                    int stage = 0;
                    uint32_t phase = 0;
                    for (int it = 0; it < iters; ++it)
                    {
                        const int tap = it % taps, cb = it / taps;
                        const int r = tap / p.S, s = tap % p.S;
                        ptx::mbar_wait(&emptyBar[stage], phase ^ 1);
                        uint8_t* a = ring + stage * Cfg::stageBytes;
                        uint8_t* b = a + kABytes;
                        ptx::mbar_arrive_expect_tx(&fullBar[stage], Cfg::stageBytes);
                        // x viewed as (W, C, H, N): box {32, 32, 4, 1} at the tap-shifted origin; OOB -> 0 (= zero padding)
                        ptx::tma_load_4d(a, &mapX, &fullBar[stage], ow0 - p.padX + s, cb * kBlockC, oh0 - p.padY + r, n);
                        // repacked filters (Cp, K, taps): box {32, BN, 1}
                        ptx::tma_load_3d(b, &mapW, &fullBar[stage], cb * kBlockC, k0, tap);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            else if (warp == 1)
            {
                if (lane == 0)
                {
                    // ===== MMA issuer =====
                    constexpr uint32_t idesc = ptx::idesc_tf32(128, BN, /*A MN-major*/ 1, /*B K-major*/ 0);
                    int stage = 0;
                    uint32_t phase = 0;
                    for (int it = 0; it < iters; ++it)
                    {
                        ptx::mbar_wait(&fullBar[stage], phase);
                        ptx::tc_fence_after_sync();
                        const uint32_t a = ptx::smem_u32(ring + stage * Cfg::stageBytes);
                        const uint32_t b = a + kABytes;
#pragma unroll
                        for (int kk = 0; kk < kBlockC / 8; ++kk)
                        {
                            // A: MN-major SW128. Row atoms (h) 4 KB apart (LBO); this K=8 slice is the kk-th 1 KB channel group.
                            const uint64_t da = ptx::smem_desc_sw128(a + kk * 1024, /*LBO*/ kBlockC * 128, /*SBO*/ 1024);
                            // B: K-major SW128. 8-filter groups 1 KB apart (SBO); this K=8 slice starts 32 B into the 128 B row.
                            const uint64_t db = ptx::smem_desc_sw128(b + kk * 32, /*LBO*/ 16, /*SBO*/ 1024);
                            ptx::mma_tf32_ss(tmemBase, da, db, idesc, (it | kk) != 0);
                        }
                        ptx::mma_commit(&emptyBar[stage]); // slot reusable once these MMAs have read it
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    ptx::mma_commit(accBar); // accumulator complete
                }
            }
            else
            {
                // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 = output row of the tile =====
                const int q = warp & 3;
                const int oh = oh0 + q, ow = ow0 + lane;
                ptx::mbar_wait(accBar, 0);
                ptx::tc_fence_after_sync();
                const bool pixelOk = oh < p.Ho && ow < p.Wo;
                float* yp = y + n * p.yStrideN + (long long)oh * p.Wo + ow;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32)
                {
                    if (k0 + c0 >= p.K)
                        break; // warp-uniform
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(tmemBase + ((uint32_t)(q * 32) << 16) + c0, v);
                    ptx::tmem_ld_wait();
                    if (pixelOk)
                    {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                        {
                            const int k = k0 + c0 + j;
                            if (k < p.K)
                            {
                                float f = __uint_as_float(v[j]);
                                if (bias)
                                    f += __ldg(bias + k);
                                yp[k * p.yStrideK] = apply_activation(p.act, p.alpha, f);
                            }
                        }
                    }
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemBase, BN);
            }
        }

        // ---------------------------------------------------------------- host side
        typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

        EncodeTiledFn encode_fn()
        {
            static EncodeTiledFn fn = nullptr;
            static std::once_flag once;
            std::call_once(once, [] {
                void* p = nullptr;
                cudaDriverEntryPointQueryResult q;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
                    fn = (EncodeTiledFn)p;
            });
            return fn;
        }

        int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* stridesBytes, const cuuint32_t* box)
        {
            EncodeTiledFn fn = encode_fn();
            if (!fn)
                return fail(NB200_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, stridesBytes, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS)
                return fail(NB200_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return NB200_OK;
        }

        inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

        int pick_bn(int K)
        {
            return K <= 64 ? 64 : 128;
        }

        // Forward-shaped problem: act tensor `in` (N, Cin, Hin, Win) -> out (N, Kout, Hout, Wout), filters repacked by `mode`.
        struct FwdShape
        {
            int N, Cin, Hin, Win, Kout, Hout, Wout, R, S, padX, padY;
        };

        bool shape_ok(const FwdShape& f)
        {
            return f.Win % 4 == 0 && f.Win >= kTileW && f.Hin >= 1 && f.Cin >= 8 && f.Kout >= 8 && f.padX >= 0 && f.padY >= 0 &&
                   f.R * f.S <= 64 && f.N >= 1;
        }

        size_t repack_bytes(const FwdShape& f)
        {
            return (size_t)f.R * f.S * f.Kout * round_up(f.Cin, kBlockC) * sizeof(float);
        }

        template <int BN>
        int launch_fprop(const FwdShape& f, const CUtensorMap& mapX, const CUtensorMap& mapW, const FpropParams& p, const float* bias,
                         float* out, cudaStream_t st)
        {
            using Cfg = FpropCfg<BN>;
            static bool attrSet = false;
            if (!attrSet)
            {
                NB200_CUDA_TRY(cudaFuncSetAttribute(tc_fprop_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smemBytes));
                attrSet = true;
            }
            const long long tiles = (long long)p.tilesK * p.tilesW * p.tilesH * f.N;
            if (tiles > 0x7FFFFFFFll)
                return fail(NB200_E_UNSUPPORTED, "too many tiles");
            tc_fprop_kernel<BN><<<(unsigned)tiles, kThreads, Cfg::smemBytes, st>>>(mapX, mapW, p, bias, out);
            NB200_CUDA_TRY(cudaGetLastError());
            return NB200_OK;
        }

        int run_fwd_shaped(const FwdShape& f, int repackMode, int wK, int wC, const float* in, const float* w, const float* bias, int act,
                           float alpha, float* out, void* ws, size_t wsBytes, cudaStream_t st)
        {
            const size_t need = repack_bytes(f);
            if (wsBytes < need || !ws)
                return fail(NB200_E_WORKSPACE, "tcgen05 conv needs %zu workspace bytes, got %zu", need, wsBytes);
            if (((uintptr_t)in & 15) || ((uintptr_t)ws & 15))
                return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");

            const int Cp = round_up(f.Cin, kBlockC);
            float* wr = (float*)ws;
            {
                const long long total = (long long)f.R * f.S * f.Kout * Cp;
                const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
                repack_filters_kernel<<<blocks, 256, 0, st>>>(w, wr, wK, wC, f.R, f.S, f.Kout, Cp, repackMode);
                NB200_CUDA_TRY(cudaGetLastError());
            }

            CUtensorMap mapX, mapW;
            {
                // activation viewed as (W, C, H, N) so that one box lands in smem as [h][c][w]
                cuuint64_t dims[4] = {(cuuint64_t)f.Win, (cuuint64_t)f.Cin, (cuuint64_t)f.Hin, (cuuint64_t)f.N};
                cuuint64_t strides[3] = {(cuuint64_t)f.Hin * f.Win * 4, (cuuint64_t)f.Win * 4, (cuuint64_t)f.Cin * f.Hin * f.Win * 4};
                cuuint32_t box[4] = {kTileW, kBlockC, kTileH, 1};
                int rc = make_map(&mapX, in, 4, dims, strides, box);
                if (rc) return rc;
            }
            const int BN = pick_bn(f.Kout);
            {
                cuuint64_t dims[3] = {(cuuint64_t)Cp, (cuuint64_t)f.Kout, (cuuint64_t)(f.R * f.S)};
                cuuint64_t strides[2] = {(cuuint64_t)Cp * 4, (cuuint64_t)Cp * f.Kout * 4};
                cuuint32_t box[3] = {kBlockC, (cuuint32_t)BN, 1};
                int rc = make_map(&mapW, wr, 3, dims, strides, box);
                if (rc) return rc;
            }

            FpropParams p;
            p.Cblocks = Cp / kBlockC;
            p.R = f.R; p.S = f.S; p.padX = f.padX; p.padY = f.padY;
            p.Ho = f.Hout; p.Wo = f.Wout; p.K = f.Kout;
            p.tilesW = ceil_div(f.Wout, kTileW); p.tilesH = ceil_div(f.Hout, kTileH); p.tilesK = ceil_div(f.Kout, BN);
            p.act = act; p.alpha = alpha;
            p.yStrideK = (long long)f.Hout * f.Wout;
            p.yStrideN = p.yStrideK * f.Kout;
            return BN == 64 ? launch_fprop<64>(f, mapX, mapW, p, bias, out, st) : launch_fprop<128>(f, mapX, mapW, p, bias, out, st);
        }

        FwdShape fwd_shape(const nb200_conv_desc& d)
        {
            return FwdShape{d.N, d.C, d.H, d.W, d.K, d.Ho, d.Wo, d.R, d.S, d.padX, d.padY};
        }

        // stride-1 input gradient as a forward conv of dy: pad' = F-1-pad, flipped taps, filter/channel roles swapped
        FwdShape dgrad_shape(const nb200_conv_desc& d)
        {
            return FwdShape{d.N, d.K, d.Ho, d.Wo, d.C, d.H, d.W, d.R, d.S, d.S - 1 - d.padX, d.R - 1 - d.padY};
        }
    }

    bool tc_forward_supported(const nb200_conv_desc& d)
    {
        return d.fmt == NB200_NCHW && d.math == NB200_MATH_TF32 && d.stride == 1 && shape_ok(fwd_shape(d));
    }

    bool tc_input_gradient_supported(const nb200_conv_desc& d)
    {
        if (d.fmt != NB200_NCHW || d.math != NB200_MATH_TF32 || d.stride != 1)
            return false;
        // the gather needs dx extent == what a forward conv of dy with pad' = F-1-pad produces
        if (d.H != d.Ho + d.R - 1 - 2 * d.padY || d.W != d.Wo + d.S - 1 - 2 * d.padX)
            return false;
        return shape_ok(dgrad_shape(d));
    }

    bool tc_kernels_gradient_supported(const nb200_conv_desc&) { return false; }

    size_t tc_workspace_bytes(int op, const nb200_conv_desc& d)
    {
        switch (op)
        {
        case NB200_OP_FORWARD: return repack_bytes(fwd_shape(d));
        case NB200_OP_INPUT_GRADIENT: return repack_bytes(dgrad_shape(d));
        default: return 0;
        }
    }

    int tc_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha, float* y, void* ws,
                   size_t wsBytes, cudaStream_t st)
    {
        return run_fwd_shaped(fwd_shape(d), 0, d.K, d.C, x, w, bias, act, alpha, y, ws, wsBytes, st);
    }

    int tc_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, void* ws, size_t wsBytes, cudaStream_t st)
    {
        return run_fwd_shaped(dgrad_shape(d), 1, d.K, d.C, dy, w, nullptr, NB200_ACT_IDENTITY, 0.f, dx, ws, wsBytes, st);
    }

    int tc_kernels_gradient(const nb200_conv_desc&, const float*, const float*, float*, void*, size_t, cudaStream_t)
    {
        return fail(NB200_E_UNSUPPORTED, "tcgen05 kernel gradient not available");
    }
}
