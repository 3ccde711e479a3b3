// tcgen05 / TMA implicit-GEMM convolution kernels for NCHW fp32 tensors (TF32 tensor-core math).
//
// Forward (and the stride-1 input gradient, which is a forward conv of dy with flipped, transposed filters):
//
//   GEMM view      D[M = pixels][N = filters] = sum over (tap, channel) A[pixel][channel@tap] * B[filter][channel@tap]
//   CTA tile       128 pixels (4 output rows x 32 output columns of one image) x BN filters, full reduction.
//   halo tile      For each block of 32 input channels ONE 4-D TMA box {WB w, 4+R-1 h, 32 c, 1 n} brings the input
//                  halo tile into shared memory (zero padding = TMA out-of-bounds fill). It is read once from L2 and
//                  serves all R*S filter taps -- no im2col buffer, and no per-tap re-fetch.
//   A operand      lives in TENSOR MEMORY. TMA box origins (and UMMA shared-memory descriptors) are 16-byte granular,
//                  so a one-pixel tap shift along W cannot be expressed by either; instead 4 "converter" warps read the
//                  tap-shifted window out of the halo tile (thread = pixel, conflict-free LDS), round to TF32
//                  (cvt.rna) and tcgen05.st it as a 128-lane x 32-column K-major A tile. The MMA then runs in
//                  TS form (A from TMEM, B from shared memory), which also halves shared-memory operand traffic.
//   B operand      filters repacked once per call to [tap][filter][channel] (channel contiguous, TF32-rounded):
//                  K-major SWIZZLE_128B, one 3-D TMA box {32 c, BN k, 1 tap} per (channel block, tap).
//   accumulator    128 lanes x BN columns of TMEM, fp32; one thread issues tcgen05.mma (M128 x BN x K8, 4 per tap).
//   epilogue       the converter warps read TMEM with tcgen05.ld (warp q owns output row q of the tile, lane = output
//                  column, so every store instruction writes one full 128-byte line of y), fused bias + activation.
//   pipeline       warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = converters/epilogue. Rings guarded by
//                  mbarriers: halo (TMA -> converters), B (TMA -> MMA), A-in-TMEM (converters -> MMA);
//                  tcgen05.commit releases B and A slots and signals the epilogue. Two CTAs are co-resident per SM
//                  so one tile's epilogue and prologue overlap the other's MMAs.
#include <cuda.h>
#include <mutex>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace nb200
{
    thread_local int g_tcFilterMode = kFiltersRepack; // set by the *_prepared / prepare_filters entry points (api.cu)

    namespace
    {
        constexpr int kTileW = 32;   // output columns per tile (= lanes of a converter warp)
        constexpr int kTileH = 4;    // output rows per tile (= converter warps) -> M = 128
        constexpr int kBlockC = 32;  // reduction channels per step (4 MMAs of K = 8)
        constexpr int kThreads = 320;        // kernel gradient: producer + MMA + 2 groups of 4 converter warps
        constexpr int kFpropThreads = 352;   // forward: filter producer + MMA + halo producer + 2 groups of 4 converter warps
        constexpr int kFirstConvWarp = 3;
        constexpr int kConvGroups = 2;
        constexpr int kAStagesMax = 8;
        // A tiles in TMEM (32 columns each): BN = 256 owns all 512 columns (256 accumulator + 8 A tiles); BN <= 128 shares the
        // SM with a second CTA (256 columns: accumulator + 4 A tiles).
        __host__ __device__ constexpr int a_stages(int BN) { return BN > 128 ? 8 : 4; }
        constexpr int kSmemBudget2 = 112 * 1024; // per CTA when two CTAs share an SM (BN <= 128)
        constexpr int kSmemBudget1 = 220 * 1024; // one CTA per SM (BN = 256)

        // Optional wait-time instrumentation (NB200_DEBUG_WAITS=1): each role accumulates the cycles it spends blocked on
        // each barrier class into dbg[cta][slot]. Null pointer = disabled (one uniform branch per wait).
        enum { kDbgBFull = 0, kDbgAFull, kDbgBEmpty, kDbgXEmpty, kDbgXFull, kDbgAEmpty, kDbgAcc, kDbgTotal, kDbgSlots = 8 };

        struct WaitAcc
        {
            long long v[kDbgSlots];
            long long* sink; // global row to flush into, or nullptr (instrumentation off)
            __device__ __forceinline__ explicit WaitAcc(long long* s) : sink(s)
            {
#pragma unroll
                for (int i = 0; i < kDbgSlots; ++i) v[i] = 0;
            }
            __device__ __forceinline__ void flush()
            {
                if (sink && (threadIdx.x & 31) == 0) // every lane of the role's warp keeps the (identical) timers
                {
#pragma unroll
                    for (int i = 0; i < kDbgSlots; ++i)
                        if (v[i]) sink[i] = v[i];
                }
            }
        };

        __device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, WaitAcc& acc, int slot)
        {
            if (acc.sink == nullptr)
            {
                ptx::mbar_wait(bar, parity);
                return;
            }
            const long long t0 = clock64();
            ptx::mbar_wait(bar, parity);
            acc.v[slot] += clock64() - t0;
        }

        __device__ __forceinline__ void timed_wait(uint32_t bar, uint32_t parity, WaitAcc& acc, int slot)
        {
            if (acc.sink == nullptr)
            {
                ptx::mbar_wait(bar, parity);
                return;
            }
            const long long t0 = clock64();
            ptx::mbar_wait(bar, parity);
            acc.v[slot] += clock64() - t0;
        }

        struct FpropParams
        {
            int Cblocks;      // ceil(C / 32)
            int R, S;
            int padX, padY;
            int wOff;         // halo tile starts at ow0 - wOff (multiple of 4 so the TMA origin is 16-byte aligned)
            int WB, HR;       // halo tile extent
            int xStages, bStages;
            int Ho, Wo, K;    // output extent and filter count
            int tilesW, tilesH, tilesK;
            int act;
            float alpha;
            long long yStrideN, yStrideK; // elements
            long long* dbg;               // wait-time instrumentation sink or nullptr
            // channel split (tc_fprop_kernel only): CTA sp of a tile reduces channel blocks [sp * cbPer, ...) and writes a raw
            // partial to partial + sp * partialStride; fprop_split_reduce_kernel adds them, then bias and activation
            int splits, cbPer;
            float* partial;
            long long partialStride;
        };

        // ---------------------------------------------------------------- filter repack
        // out[tap][row][col] (col padded to outCp with zeros), TF32-rounded (cvt.rna).
        // mode 0 (forward):           tap = r*S+s,                 row = filter,  col = channel <- w[row][col][r][s]
        // mode 1 (input gradient):    tap = (R-1-r)*S + (S-1-s),   row = channel, col = filter  <- w[col][row][r][s]
        // x3 != 0: out holds two such tensors back to back, hi = tf32(v) then lo = tf32(v - hi) (3xTF32 operand split).
        __global__ void repack_filters_kernel(const float* __restrict__ w, float* __restrict__ out, int K, int C, int R, int S,
                                              int outRows, int outCp, int mode, int x3)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
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
                else if (mode == 1)
                {
                    const int r = R - 1 - tap / S, s = S - 1 - tap % S;
                    if (col < K)
                        v = w[(((long long)col * C + row) * R + r) * S + s];
                }
                else // mode 2 (gathered input gradient): tap = r*S+s unflipped, row = channel, col = filter
                {
                    if (col < K)
                        v = w[((long long)col * C + row) * R * S + tap];
                }
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
                out[i] = __uint_as_float(t);
                if (x3)
                {
                    uint32_t tl;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tl) : "f"(v - __uint_as_float(t)));
                    out[total + i] = __uint_as_float(tl);
                }
            }
        }

        // ---------------------------------------------------------------- forward kernel
        // X3 = 3xTF32: every operand is split into hi = tf32(v) and lo = tf32(v - hi); the product is accumulated as
        // hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-22 relative), which recovers fp32-class accuracy
        // (<= 1e-5 max-normalised against the reference) at a third of the TF32 rate. A tiles carry [hi | lo] (64 columns),
        // filter stages carry a hi and a lo tile (the repack writes both).
        // Tiled variants of the filter repack (taps <= 32): the element-wise kernel above reads w with a stride of R*S floats
        // per thread (9x read amplification for 3x3) or writes 4-byte pieces; at 24 launches per VGG16 step that was 2.7 % of
        // the step. These stage a tile in shared memory so that both the reads and the writes are contiguous runs.
        __device__ __forceinline__ void repack_store(float* __restrict__ out, long long i, long long total, float v, int x3)
        {
            uint32_t t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
            out[i] = __uint_as_float(t);
            if (x3)
            {
                uint32_t tl;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tl) : "f"(v - __uint_as_float(t)));
                out[total + i] = __uint_as_float(tl);
            }
        }

        // mode 0: out[tap][k][c] = w[k][c][tap]. Block = (32-channel tile, filter k): reads 32*taps contiguous floats.
        __global__ void repack_fwd_tiled_kernel(const float* __restrict__ w, float* __restrict__ out, int K, int C, int taps, int outCp, int x3)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            extern __shared__ float tile[]; // [32 channels][taps]
            const int k = blockIdx.y, c0 = blockIdx.x * 32;
            const int n = 32 * taps;
            const long long base = ((long long)k * C + c0) * taps;
            const int valid = (C - c0 < 32 ? (C - c0 > 0 ? C - c0 : 0) : 32) * taps;
            for (int j = threadIdx.x; j < n; j += blockDim.x)
                tile[j] = j < valid ? w[base + j] : 0.f;
            __syncthreads();
            const long long total = (long long)taps * K * outCp;
            for (int j = threadIdx.x; j < n; j += blockDim.x)
            {
                const int tap = j >> 5, cl = j & 31;
                repack_store(out, ((long long)tap * K + k) * outCp + c0 + cl, total, tile[cl * taps + tap], x3);
            }
        }

        // modes 1 / 2: out[tap'][c][k] = w[k][c][tap] (mode 1: tap' = the flipped tap). Block = (8-channel tile, 32-filter
        // tile): reads 32 runs of 8*taps floats, writes 8*taps runs of 32 floats.
        __global__ void repack_dgrad_tiled_kernel(const float* __restrict__ w, float* __restrict__ out, int K, int C, int R, int S, int outRows,
                                                  int outCp, int flip, int x3)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            extern __shared__ float tile[]; // [32 filters][8 * taps + 1]
            const int taps = R * S, run = 8 * taps, pitch = run + 1;
            const int c0 = blockIdx.x * 8, k0 = blockIdx.y * 32;
            const int validRun = (C - c0 < 8 ? C - c0 : 8) * taps;
            for (int j = threadIdx.x; j < 32 * run; j += blockDim.x)
            {
                const int kl = j / run, i = j - kl * run;
                tile[kl * pitch + i] = (k0 + kl < K && i < validRun) ? w[((long long)(k0 + kl) * C + c0) * taps + i] : 0.f;
            }
            __syncthreads();
            const long long total = (long long)taps * outRows * outCp;
            for (int j = threadIdx.x; j < 32 * run; j += blockDim.x)
            {
                const int kl = j & 31, i = j >> 5;          // i = cl * taps + tap
                const int cl = i / taps, tap = i - cl * taps;
                if (c0 + cl >= outRows)
                    continue;
                const int tapOut = flip ? (R - 1 - tap / S) * S + (S - 1 - tap % S) : tap;
                repack_store(out, ((long long)tapOut * outRows + c0 + cl) * outCp + k0 + kl, total, tile[kl * pitch + i], x3);
            }
        }

        // dw[k][c][tap] = sum over splits of partial[split][tap][k][c]; block = (32-channel tile, filter k): coalesced reads
        // along c, one contiguous run of 32*taps floats written (the element-wise reduce wrote with a stride of `taps` floats).
        __global__ void wgrad_reduce_tiled_kernel(const float* __restrict__ ws, float* __restrict__ dw, int K, int C, int taps, int splits)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            extern __shared__ float tile[]; // [32 channels][taps]
            const int k = blockIdx.y, c0 = blockIdx.x * 32;
            const int n = 32 * taps;
            const long long total = (long long)K * C * taps;
            for (int j = threadIdx.x; j < n; j += blockDim.x)
            {
                const int tap = j >> 5, cl = j & 31;
                float acc = 0.f;
                if (c0 + cl < C)
                {
                    const float* src = ws + ((long long)tap * K + k) * C + c0 + cl;
                    for (int sp = 0; sp < splits; ++sp)
                        acc += src[(long long)sp * total];
                }
                tile[cl * taps + tap] = acc;
            }
            __syncthreads();
            const int valid = (C - c0 < 32 ? C - c0 : 32) * taps;
            float* dst = dw + ((long long)k * C + c0) * taps;
            for (int j = threadIdx.x; j < valid; j += blockDim.x)
                dst[j] = tile[j];
        }

        // Bias + activation + store of one 32-filter chunk of this thread's pixel (lanes = 32 consecutive output columns, so
        // every store instruction writes one full 128-byte line). The activation is a template parameter: the per-element
        // switch and the exp paths stay out of the identity / ReLU / leaky-ReLU instantiations (the epilogue used to cost
        // as many issue slots as the whole main loop of a 64-filter tile). Lane j carries bias[kBase + j]; shuffles broadcast it.
        template <int ACT>
        __device__ __forceinline__ void store_chunk_t(float* yp, long long strideK, int kBase, int K, float biasLane, int act, float alpha,
                                                      bool pixelOk, const uint32_t (&v)[32])
        {
#pragma unroll
            for (int j = 0; j < 32; ++j)
            {
                const float b = __shfl_sync(0xffffffffu, biasLane, j);
                if (pixelOk && kBase + j < K)
                {
                    float f = __uint_as_float(v[j]) + b;
                    if constexpr (ACT == NB200_ACT_RELU) f = f > 0.f ? f : 0.f;
                    else if constexpr (ACT == NB200_ACT_LEAKY_RELU) f = f >= 0.f ? f : alpha * f;
                    else if constexpr (ACT != NB200_ACT_IDENTITY) f = apply_activation(act, alpha, f);
                    yp[(long long)(kBase + j) * strideK] = f;
                }
            }
        }

        __device__ __forceinline__ void store_chunk(float* yp, long long strideK, int kBase, int K, const float* __restrict__ bias, int lane,
                                                    int act, float alpha, bool pixelOk, const uint32_t (&v)[32])
        {
            const float bl = (bias && kBase + lane < K) ? __ldg(bias + kBase + lane) : 0.f;
            switch (act)
            {
            case NB200_ACT_IDENTITY: store_chunk_t<NB200_ACT_IDENTITY>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v); break;
            case NB200_ACT_RELU: store_chunk_t<NB200_ACT_RELU>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v); break;
            case NB200_ACT_LEAKY_RELU: store_chunk_t<NB200_ACT_LEAKY_RELU>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v); break;
            default: store_chunk_t<-1>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v); break;
            }
        }

        template <int BN, bool X3>
        __global__ void __launch_bounds__(kFpropThreads, ((BN > 128 || X3) ? 1 : 2))
        tc_fprop_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, FpropParams p,
                        const float* __restrict__ bias, float* __restrict__ y)
        {
            constexpr uint32_t kBTile = BN * kBlockC * 4;             // one filter tile
            constexpr uint32_t kBBytes = kBTile * (X3 ? 2 : 1);       // a filter stage: hi tile (+ lo tile)
            constexpr int kAStages = X3 ? 4 : a_stages(BN);
            constexpr int kACols = X3 ? 2 * kBlockC : kBlockC;        // TMEM columns per A stage
            // BN <= 128: two CTAs per SM x 256 columns; BN = 256 or 3xTF32: one CTA per SM x 512 columns
            constexpr uint32_t kTmemCols = (BN > 128 || X3) ? 512 : 256;
            static_assert(BN + kAStages * kACols <= kTmemCols, "TMEM budget");

            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            const uint32_t xBytes = (uint32_t)(kBlockC * p.HR * p.WB * 4);
            const uint32_t xBytesPad = (xBytes + 1023) & ~1023u;
            uint8_t* bRing = smem;                                   // 1 KB aligned (128B swizzle atoms)
            uint8_t* xRing = smem + p.bStages * kBBytes;
            uint64_t* bars = (uint64_t*)(xRing + p.xStages * xBytesPad);
            uint64_t* bFull = bars;            // [bStages]
            uint64_t* bEmpty = bFull + 8;      // [bStages]
            uint64_t* xFull = bEmpty + 8;      // [xStages]
            uint64_t* xEmpty = xFull + 4;      // [xStages]
            uint64_t* aFull = xEmpty + 4;      // [kAStages]
            uint64_t* aEmpty = aFull + kAStagesMax;
            uint64_t* accBar = aEmpty + kAStagesMax;
            uint64_t* accFree = accBar + 1;    // 3xTF32 only: the accumulator has been drained into registers
            uint32_t* tmemSlot = (uint32_t*)(accFree + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;

            // tile coordinates: filter tile fastest so CTAs sharing the same activation tile run together (L2 reuse)
            int t = blockIdx.x;
            const int sp = t % p.splits; t /= p.splits;           // channel split: this CTA's share of the reduction
            const int cbBegin = sp * p.cbPer;
            const int cbCount = min(p.cbPer, p.Cblocks - cbBegin);
            const int kt = t % p.tilesK; t /= p.tilesK;
            const int tw = t % p.tilesW; t /= p.tilesW;
            const int th = t % p.tilesH; t /= p.tilesH;
            const int n = t;
            const int ow0 = tw * kTileW, oh0 = th * kTileH, k0 = kt * BN;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapW);
                for (int s = 0; s < p.bStages; ++s) { ptx::mbar_init(&bFull[s], 1); ptx::mbar_init(&bEmpty[s], 1); }
                for (int s = 0; s < p.xStages; ++s) { ptx::mbar_init(&xFull[s], 1); ptx::mbar_init(&xEmpty[s], kTileH * kConvGroups); }
                for (int s = 0; s < kAStages; ++s) { ptx::mbar_init(&aFull[s], kTileH); ptx::mbar_init(&aEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::mbar_init(accFree, kTileH * kConvGroups);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, kTmemCols);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents(); // the next kernel's prologue may overlap this kernel (sm100_ptx.cuh)
            ptx::pdl_wait();              // nothing below runs before the previous kernel's memory is visible
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t tmemA = tmemAcc + BN;
            // 32-bit shared addresses of the barrier arrays, computed once (see sm100_ptx.cuh)
            const uint32_t bFull32 = ptx::smem_u32(bFull), bEmpty32 = ptx::smem_u32(bEmpty), xFull32 = ptx::smem_u32(xFull),
                           xEmpty32 = ptx::smem_u32(xEmpty), aFull32 = ptx::smem_u32(aFull), aEmpty32 = ptx::smem_u32(aEmpty),
                           accBar32 = ptx::smem_u32(accBar), accFree32 = ptx::smem_u32(accFree);

            const int taps = p.R * p.S;
            // per-CTA debug rows: [0] filter producer, [1] MMA issuer, [2] halo producer, [3] converter warp 3 lane 0
            long long* dbgBase = p.dbg ? p.dbg + (long long)blockIdx.x * 4 * kDbgSlots : nullptr;
            WaitAcc dbgP(dbgBase ? dbgBase : nullptr);
            WaitAcc dbgM(dbgBase ? dbgBase + kDbgSlots : nullptr);
            WaitAcc dbgX(dbgBase ? dbgBase + 2 * kDbgSlots : nullptr);
            WaitAcc dbgC((dbgBase && warp == kFirstConvWarp) ? dbgBase + 3 * kDbgSlots : nullptr);
            const long long tStart = clock64();

            if (warp == 0)
            {
                if (lane == 0)
                {
                    // ===== TMA producer (filters): one tile per (channel block, tap) =====
                    int bs = 0;
                    uint32_t bph = 0;
                    for (int cb = 0; cb < cbCount; ++cb)
                        for (int tap = 0; tap < taps; ++tap)
                        {
                            timed_wait(&bEmpty[bs], bph ^ 1, dbgP, kDbgBEmpty);
                            ptx::mbar_arrive_expect_tx(&bFull[bs], kBBytes);
                            ptx::tma_load_3d(bRing + bs * kBBytes, &mapW, &bFull[bs], (cbBegin + cb) * kBlockC, k0, tap);
                            if (X3) // lo parts live behind the hi parts in the repacked tensor (tap index + taps)
                                ptx::tma_load_3d(bRing + bs * kBBytes + kBTile, &mapW, &bFull[bs], (cbBegin + cb) * kBlockC, k0, taps + tap);
                            if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                        }
                }
            }
            else if (warp == 2)
            {
                if (lane == 0)
                {
                    // ===== TMA producer (activations): one halo tile per channel block, independent of the filter ring =====
                    int xs = 0;
                    uint32_t xph = 0;
                    for (int cb = 0; cb < cbCount; ++cb)
                    {
                        timed_wait(&xEmpty[xs], xph ^ 1, dbgX, kDbgXEmpty);
                        ptx::mbar_arrive_expect_tx(&xFull[xs], xBytes);
                        // x viewed as (W, H, C, N); origin 16-byte aligned in W; out-of-bounds elements read as 0 (= zero padding)
                        ptx::tma_load_4d(xRing + xs * xBytesPad, &mapX, &xFull[xs], ow0 - p.wOff, oh0 - p.padY, (cbBegin + cb) * kBlockC, n);
                        if (++xs == p.xStages) { xs = 0; xph ^= 1; }
                    }
                }
            }
            else if (warp == 1)
            {
                // ===== MMA issuer: D[128 x BN] += A[tmem 128 x 32] * B[smem BN x 32]^T per tap =====
                // The whole warp runs the (warp-uniform) loop; one elected lane issues the tcgen05 instructions.
                constexpr uint32_t idesc = ptx::idesc_tf32(128, BN, /*A K-major (TMEM)*/ 0, /*B K-major*/ 0);
                // B: K-major SW128: 8-filter groups 1 KB apart (SBO); a K=8 slice starts kk*32 B into the 128 B row
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(bRing), /*LBO*/ 16, /*SBO*/ 1024);
                int as = 0, bs = 0;
                uint32_t aph = 0, bph = 0;
                const int iters = taps * cbCount;
                const long long tLoop = dbgM.sink ? clock64() : 0;
                int tapc = 0, cbc = 0; // position inside the current channel block (3xTF32 drains the accumulator per block)
                for (int it = 0; it < iters; ++it)
                {
                    timed_wait(bFull32 + 8u * bs, bph, dbgM, kDbgBFull);
                    timed_wait(aFull32 + 8u * as, aph, dbgM, kDbgAFull);
                    if (X3 && tapc == 0 && cbc > 0)
                        ptx::mbar_wait(accFree32, (uint32_t)(cbc - 1) & 1); // previous block's partial sums are in registers
                    const long long tIssue = dbgM.sink ? clock64() : 0;
                    ptx::tc_fence_after_sync();
                    if (ptx::elect_one())
                    {
                        const uint64_t db = descB0 + (uint64_t)((bs * kBBytes) >> 4);
                        const uint32_t ta = tmemA + as * kACols;
#pragma unroll
                        for (int kk = 0; kk < kBlockC / 8; ++kk)
                        {
                            ptx::mma_tf32_ts(tmemAcc, ta + kk * 8, db + kk * 2, idesc, ((X3 ? tapc : it) | kk) != 0);        // hi * hi
                            if (X3)
                            {
                                ptx::mma_tf32_ts(tmemAcc, ta + kk * 8, db + (kBTile >> 4) + kk * 2, idesc, 1);            // hi * lo
                                ptx::mma_tf32_ts(tmemAcc, ta + kBlockC + kk * 8, db + kk * 2, idesc, 1);                  // lo * hi
                            }
                        }
                        ptx::mma_commit(aEmpty32 + 8u * as); // both slots reusable once these MMAs have consumed them
                        ptx::mma_commit(bEmpty32 + 8u * bs);
                        if (X3 && tapc == taps - 1)
                            ptx::mma_commit(accBar32); // this channel block's partial accumulator is complete
                    }
                    __syncwarp();
                    if (++tapc == taps) { tapc = 0; ++cbc; }
                    if (dbgM.sink) dbgM.v[kDbgAcc] += clock64() - tIssue;
                    if (++as == kAStages) { as = 0; aph ^= 1; }
                    if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                }
                if (dbgM.sink) dbgM.v[kDbgTotal] = clock64() - tLoop;
                if (!X3 && ptx::elect_one())
                    ptx::mma_commit(accBar32); // accumulator complete
                __syncwarp();
            }
            else
            {
                // ===== converters (then epilogue): warps 3..10 in two groups that alternate taps =====
                // TMEM lane quadrant q = warp % 4 = output row of the tile; group g handles iterations it = g (mod 2),
                // i.e. A stages {g, g+2}. Each warp overlaps the TMEM store of one tap with the loads of its next tap.
                const int q = warp & 3;
                const int g = (warp - kFirstConvWarp) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                const uint32_t chanStrideB = (uint32_t)(p.HR * p.WB * 4); // bytes between channels of the halo tile
                const uint32_t xRing32 = ptx::smem_u32(xRing);
                bool pending = false;
                int pendStage = 0;
                const long long tConv = dbgC.sink ? clock64() : 0;
                // 3xTF32: the tensor core's fp32 accumulator truncates on every add, a bias that grows with the length of
                // the accumulation chain (measured 3.8e-5 max-normalised at C = 512). So the chain is cut per channel block:
                // each block's partial sums are drained from TMEM and added, round-to-nearest, into registers.
                constexpr int kOwnChunks = X3 ? (BN / (32 * kConvGroups)) : 1;
                float racc[kOwnChunks][32];
                if (X3)
                {
#pragma unroll
                    for (int ch = 0; ch < kOwnChunks; ++ch)
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            racc[ch][j] = 0.f;
                }
                for (int cb = 0; cb < cbCount; ++cb)
                {
                    const int xs = cb % p.xStages;
                    timed_wait(xFull32 + 8u * xs, (uint32_t)(cb / p.xStages) & 1, dbgC, kDbgXFull);
                    const uint32_t xt = xRing32 + xs * xBytesPad;
                    int r = 0, s = 0;
                    for (int tap = 0; tap < taps; ++tap, (++s == p.S ? (s = 0, ++r) : 0))
                    {
                        const int it = cb * taps + tap;
                        if ((it & 1) != g)
                            continue;
                        // pixel (q, lane) of the tile, tap (r, s): halo row q + r, halo column lane + s - padX + wOff
                        const uint32_t src = xt + (uint32_t)(((q + r) * p.WB + (lane + s - p.padX + p.wOff)) << 2);
                        uint32_t v[kBlockC];
                        const long long tA = dbgC.sink ? clock64() : 0;
                        if (chanStrideB == 960u) // 3x3 filters: compile-time channel pitch -> LDS with immediate offsets
                        {
#pragma unroll
                            for (int c = 0; c < kBlockC; ++c)
                                v[c] = ptx::tf32_round_bits(ptx::lds_b32(src + c * 960u));
                        }
                        else
                        {
#pragma unroll
                            for (int c = 0; c < kBlockC; ++c)
                                v[c] = ptx::tf32_round_bits(ptx::lds_b32(src + c * chanStrideB));
                        }
                        uint32_t vlo[kBlockC]; // only live in the 3xTF32 instantiation
                        if (X3)
                        {
#pragma unroll
                            for (int c = 0; c < kBlockC; ++c)
                            {
                                const float full = __uint_as_float(v[c] - 0x1000u);           // undo the rounding add: the raw fp32 value
                                const uint32_t hi = v[c] & 0xFFFFE000u;                         // exactly what the tensor core will read
                                vlo[c] = ptx::tf32_round_bits(__float_as_uint(full - __uint_as_float(hi)));
                                v[c] = hi;
                            }
                        }
                        if (pending)
                        {
                            // the previous tap's store has had the whole load phase to land
                            if (dbgC.sink)
                            {
                                // v[] is consumed here so the clock read sits after the loads have landed
                                uint32_t x = 0;
#pragma unroll
                                for (int c = 0; c < kBlockC; ++c) x ^= v[c];
                                const long long t0 = clock64() + (x == 0x12345u ? 1 : 0);
                                dbgC.v[kDbgAFull] += t0 - tA;        // converter row: slot AFull = load + round phase
                                ptx::tmem_st_wait();
                                const long long t1 = clock64();
                                dbgC.v[kDbgBFull] += t1 - t0;        // slot BFull = tcgen05.st completion wait
                                ptx::tc_fence_before_sync();
                                __syncwarp();
                                if (lane == 0)
                                    ptx::mbar_arrive(aFull32 + 8u * pendStage);
                                __syncwarp();
                                dbgC.v[kDbgXEmpty] += clock64() - t1; // slot XEmpty = fence + arrive
                            }
                            else
                            {
                                ptx::tmem_st_wait();
                                ptx::tc_fence_before_sync();
                                __syncwarp();
                                if (lane == 0)
                                    ptx::mbar_arrive(aFull32 + 8u * pendStage);
                            }
                        }
                        const int as = it & (kAStages - 1);
                        timed_wait(aEmpty32 + 8u * as, ((uint32_t)(it / kAStages) & 1) ^ 1, dbgC, kDbgAEmpty);
                        const long long tS = dbgC.sink ? clock64() : 0;
                        ptx::tc_fence_after_sync();
                        ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * kACols, v);
                        if (X3)
                            ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * kACols + kBlockC, vlo);
                        if (dbgC.sink) dbgC.v[kDbgBEmpty] += clock64() - tS; // slot BEmpty = tcgen05.st issue
                        pending = true;
                        pendStage = as;
                    }
                    // every load of this halo tile has been consumed into registers (the stores above read them)
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(xEmpty32 + 8u * xs);
                    if (X3)
                    {
                        if (pending) // the MMAs of this block cannot finish before its last A tile is published
                        {
                            ptx::tmem_st_wait();
                            ptx::tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0)
                                ptx::mbar_arrive(aFull32 + 8u * pendStage);
                            pending = false;
                        }
                        ptx::mbar_wait(accBar32, (uint32_t)cb & 1);
                        ptx::tc_fence_after_sync();
#pragma unroll
                        for (int ch = 0; ch < kOwnChunks; ++ch)
                        {
                            uint32_t pv[32];
                            ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + (g + ch * kConvGroups) * 32, pv);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                racc[ch][j] += __uint_as_float(pv[j]);
                        }
                        ptx::tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0)
                            ptx::mbar_arrive(accFree32);
                    }
                }
                if (pending)
                {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(aFull32 + 8u * pendStage);
                }

                // ----- epilogue: the two warps of a quadrant split the filter columns -----
                const int oh = oh0 + q, ow = ow0 + lane;
                if (!X3)
                {
                    timed_wait(accBar32, 0, dbgC, kDbgAcc);
                    ptx::tc_fence_after_sync();
                }
                const bool pixelOk = oh < p.Ho && ow < p.Wo;
                const bool split = p.splits > 1; // raw partial sums; bias and activation wait for the reduce kernel
                float* yp = (split ? p.partial + sp * p.partialStride : y) + n * p.yStrideN + (long long)oh * p.Wo + ow;
                // bias + activation + one full 128-byte line per store instruction (lanes = 32 consecutive output columns)
                auto store_chunk = [&](int c0, const uint32_t (&v)[32]) {
                    nb200::store_chunk(yp, p.yStrideK, k0 + c0, p.K, split ? nullptr : bias, lane, split ? (int)NB200_ACT_IDENTITY : p.act, p.alpha,
                                       pixelOk, v);
                };
                if (X3)
                {
#pragma unroll
                    for (int ch = 0; ch < kOwnChunks; ++ch)
                    {
                        const int c0 = (g + ch * kConvGroups) * 32;
                        if (k0 + c0 < p.K)
                        {
                            uint32_t v[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = __float_as_uint(racc[ch][j]);
                            store_chunk(c0, v);
                        }
                    }
                }
                else
                {
#pragma unroll 1
                    for (int c0 = g * 32; c0 < BN; c0 += 32 * kConvGroups)
                    {
                        if (k0 + c0 >= p.K)
                            break; // warp-uniform
                        uint32_t v[32];
                        ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + c0, v);
                        ptx::tmem_ld_wait();
                        store_chunk(c0, v);
                    }
                }
                if (dbgC.sink) dbgC.v[kDbgTotal] = clock64() - tConv;
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemAcc, kTmemCols);
            }
            if (warp == 0) dbgP.flush();
            if (warp == 1) dbgM.flush();
            if (warp == 2) dbgX.flush();
            if (warp == kFirstConvWarp) dbgC.flush();
        }




        // ---------------------------------------------------------------- forward kernel, 256-pixel tile (two M halves)
        // Measured on B200 an SM ingests ~48 B/clk from L2 (profiles/r1_microbench_*): the 128-pixel tile streams a
        // 128 B x BN filter tile per 4 MMAs, which for BN = 256 is 35 KB per 512 MMA cycles -- ingest-bound at ~70 %. This
        // variant lets TWO 128-pixel halves (8 output rows x 32 columns) share every filter tile: one CTA per SM, accumulators
        // D0 | D1 (2 x BN <= 256 TMEM columns), one A ring per half (4 x 32 columns each), converter group g owns half g for
        // every tap, the MMA warp issues both halves against the same B stage. Filter bytes per MMA cycle halve
        // (23 KB per 512 cycles at BN = 128), the halo tile grows to 8+R-1 rows.
        constexpr int kHalfStages = 4;

        template <int BN>
        __global__ void __launch_bounds__(kFpropThreads, 1)
        tc_fprop_m256_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, FpropParams p,
                             const float* __restrict__ bias, float* __restrict__ y)
        {
            static_assert(BN <= 128, "two accumulators of BN columns plus two A rings must fit 512 TMEM columns");
            constexpr uint32_t kBBytes = BN * kBlockC * 4;
            constexpr uint32_t kTmemCols = 512;

            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            const uint32_t xBytes = (uint32_t)(kBlockC * p.HR * p.WB * 4);
            const uint32_t xBytesPad = (xBytes + 1023) & ~1023u;
            uint8_t* bRing = smem;
            uint8_t* xRing = smem + p.bStages * kBBytes;
            uint64_t* bars = (uint64_t*)(xRing + p.xStages * xBytesPad);
            uint64_t* bFull = bars;            // [bStages]
            uint64_t* bEmpty = bFull + 8;
            uint64_t* xFull = bEmpty + 8;      // [xStages]
            uint64_t* xEmpty = xFull + 4;
            uint64_t* aFull = xEmpty + 4;      // [2 halves][kHalfStages]
            uint64_t* aEmpty = aFull + 2 * kHalfStages;
            uint64_t* accBar = aEmpty + 2 * kHalfStages;
            uint32_t* tmemSlot = (uint32_t*)(accBar + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;

            int t = blockIdx.x;
            const int kt = t % p.tilesK; t /= p.tilesK;
            const int tw = t % p.tilesW; t /= p.tilesW;
            const int th = t % p.tilesH; t /= p.tilesH;
            const int n = t;
            const int ow0 = tw * kTileW, oh0 = th * (2 * kTileH), k0 = kt * BN;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapW);
                for (int s = 0; s < p.bStages; ++s) { ptx::mbar_init(&bFull[s], 1); ptx::mbar_init(&bEmpty[s], 1); }
                for (int s = 0; s < p.xStages; ++s) { ptx::mbar_init(&xFull[s], 1); ptx::mbar_init(&xEmpty[s], kTileH * kConvGroups); }
                for (int s = 0; s < 2 * kHalfStages; ++s) { ptx::mbar_init(&aFull[s], kTileH); ptx::mbar_init(&aEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, kTmemCols);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents(); // the next kernel's prologue may overlap this kernel (sm100_ptx.cuh)
            ptx::pdl_wait();              // nothing below runs before the previous kernel's memory is visible
            const uint32_t tmemAcc = *tmemSlot;            // D0 at +0, D1 at +BN
            const uint32_t tmemA = tmemAcc + 2 * BN;       // half h, stage s at +(h * kHalfStages + s) * 32

            const int taps = p.R * p.S;
            const int iters = taps * p.Cblocks;

            if (warp == 0)
            {
                if (lane == 0)
                {
                    int bs = 0;
                    uint32_t bph = 0;
                    for (int cb = 0; cb < p.Cblocks; ++cb)
                        for (int tap = 0; tap < taps; ++tap)
                        {
                            ptx::mbar_wait(&bEmpty[bs], bph ^ 1);
                            ptx::mbar_arrive_expect_tx(&bFull[bs], kBBytes);
                            ptx::tma_load_3d(bRing + bs * kBBytes, &mapW, &bFull[bs], cb * kBlockC, k0, tap);
                            if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                        }
                }
            }
            else if (warp == 2)
            {
                if (lane == 0)
                {
                    int xs = 0;
                    uint32_t xph = 0;
                    for (int cb = 0; cb < p.Cblocks; ++cb)
                    {
                        ptx::mbar_wait(&xEmpty[xs], xph ^ 1);
                        ptx::mbar_arrive_expect_tx(&xFull[xs], xBytes);
                        ptx::tma_load_4d(xRing + xs * xBytesPad, &mapX, &xFull[xs], ow0 - p.wOff, oh0 - p.padY, cb * kBlockC, n);
                        if (++xs == p.xStages) { xs = 0; xph ^= 1; }
                    }
                }
            }
            else if (warp == 1)
            {
                constexpr uint32_t idesc = ptx::idesc_tf32(128, BN, 0, 0);
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(bRing), 16, 1024);
                int as = 0, bs = 0;
                uint32_t aph = 0, bph = 0;
                for (int it = 0; it < iters; ++it)
                {
                    ptx::mbar_wait(&bFull[bs], bph);
                    const uint64_t db = descB0 + (uint64_t)((bs * kBBytes) >> 4);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                    {
                        ptx::mbar_wait(&aFull[h * kHalfStages + as], aph);
                        ptx::tc_fence_after_sync();
                        if (ptx::elect_one())
                        {
                            const uint32_t ta = tmemA + (h * kHalfStages + as) * kBlockC;
#pragma unroll
                            for (int kk = 0; kk < kBlockC / 8; ++kk)
                                ptx::mma_tf32_ts(tmemAcc + h * BN, ta + kk * 8, db + kk * 2, idesc, (it | kk) != 0);
                            ptx::mma_commit(&aEmpty[h * kHalfStages + as]);
                            if (h == 1)
                                ptx::mma_commit(&bEmpty[bs]); // both halves have read this filter tile
                        }
                        __syncwarp();
                    }
                    if (++as == kHalfStages) { as = 0; aph ^= 1; }
                    if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                }
                if (ptx::elect_one())
                    ptx::mma_commit(accBar);
                __syncwarp();
            }
            else
            {
                // ===== converters: group g owns tile half g (output rows 4g .. 4g+3) for every tap =====
                const int q = warp & 3;
                const int g = (warp - kFirstConvWarp) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                const uint32_t chanStrideB = (uint32_t)(p.HR * p.WB * 4);
                const uint32_t xRing32 = ptx::smem_u32(xRing);
                uint64_t* myFull = aFull + g * kHalfStages;
                uint64_t* myEmpty = aEmpty + g * kHalfStages;
                const uint32_t myA = tmemA + laneSel + g * kHalfStages * kBlockC;
                bool pending = false;
                int pendStage = 0;
                int it = 0;
                for (int cb = 0; cb < p.Cblocks; ++cb)
                {
                    const int xs = cb % p.xStages;
                    ptx::mbar_wait(&xFull[xs], (uint32_t)(cb / p.xStages) & 1);
                    const uint32_t xt = xRing32 + xs * xBytesPad;
                    int r = 0, s = 0;
                    for (int tap = 0; tap < taps; ++tap, ++it, (++s == p.S ? (s = 0, ++r) : 0))
                    {
                        const uint32_t src = xt + (uint32_t)(((g * kTileH + q + r) * p.WB + (lane + s - p.padX + p.wOff)) << 2);
                        uint32_t v[kBlockC];
#pragma unroll
                        for (int c = 0; c < kBlockC; ++c)
                            v[c] = ptx::tf32_round_bits(ptx::lds_b32(src + c * chanStrideB));
                        if (pending)
                        {
                            ptx::tmem_st_wait();
                            ptx::tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0)
                                ptx::mbar_arrive(&myFull[pendStage]);
                        }
                        const int as = it & (kHalfStages - 1);
                        ptx::mbar_wait(&myEmpty[as], ((uint32_t)(it / kHalfStages) & 1) ^ 1);
                        ptx::tc_fence_after_sync();
                        ptx::tmem_st_32x32b_x32(myA + as * kBlockC, v);
                        pending = true;
                        pendStage = as;
                    }
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(&xEmpty[xs]);
                }
                if (pending)
                {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(&myFull[pendStage]);
                }

                // ----- epilogue: this group's half, all BN columns -----
                const int oh = oh0 + g * kTileH + q, ow = ow0 + lane;
                ptx::mbar_wait(accBar, 0);
                ptx::tc_fence_after_sync();
                const bool pixelOk = oh < p.Ho && ow < p.Wo;
                float* yp = y + n * p.yStrideN + (long long)oh * p.Wo + ow;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32)
                {
                    if (k0 + c0 >= p.K)
                        break;
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + g * BN + c0, v);
                    ptx::tmem_ld_wait();
                    nb200::store_chunk(yp, p.yStrideK, k0 + c0, p.K, bias, lane, p.act, p.alpha, pixelOk, v);
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemAcc, kTmemCols);
            }
        }


        // ---------------------------------------------------------------- forward kernel, horizontal taps folded into N
        // For few filters (K <= 128) the main kernel is bound by its converter warps: every converted A tile (128 pixels x 32
        // channels, ~90 instructions per warp) feeds only 4 MMAs of N = K columns. This kernel (3x3-style filters: S = 3,
        // padX = 1) makes each A tile feed three times the columns. With lane p of a tile row standing for INPUT column
        // xc = 32 tw + p,
        //     P_s[p][k] = sum_{c,r} x[c][oh + r - padY][xc(p)] * f[k][c][r][s]            (one GEMM, N = 3 x 64 = 192)
        //     y[32 tw + i][k] = P_0[i - 1][k] + P_1[i][k] + P_2[i + 1][k]
        // so the A operand is not shifted per horizontal tap at all: one A tile per (channel block, filter ROW r), the three
        // taps of that row side by side in the B tile (192 rows x 32 channels), and the epilogue adds the three partial
        // products with lane shuffles. The tiles of a row cover input columns 0 .. W-1 exactly; the two columns they leave
        // out are the zero padding. Column 0 of a tile needs P_0 of the previous tile's lane 31 and column 31 needs P_2 of the
        // next tile's lane 0, so a CTA walks a strip (one image, 4 output rows, 64 filters) left to right, carries that one
        // P_0 value and holds back the last 8-column sector of every tile in shared memory until the next tile completes it;
        // each store instruction then writes 32 consecutive columns starting 8 left of the tile -- whole 32-byte sectors
        // (a first version that stored 30-column tiles directly ran 3x slower end to end on partial-sector writes).
        // The kernel is persistent (one CTA per SM, strips dealt round-robin) with two accumulators, so the epilogue of one
        // tile overlaps the main loop of the next: at C = 64 a tile's main loop is only ~2300 MMA cycles.
        //   warp 0 filter TMA | warp 1 MMA issue | warp 2 halo TMA | warps 4-11 converters (2 groups) | warps 12-19 epilogue
        //   TMEM: accumulators at columns [0,192) and [256,448); A stages at 192, 224, 448, 480.
        constexpr int kRtThreads = 640;
        constexpr int kRtBNK = 64;                          // filters per tile
        constexpr int kRtS = 3;                             // horizontal taps folded into N
        constexpr int kRtN = kRtBNK * kRtS;                 // MMA N
        constexpr int kRtWB = 32;                           // halo width = the tile's 32 input columns (16-byte aligned origin)
        constexpr int kRtHold = 8;                          // columns held back per tile (one 32-byte sector)
        constexpr int kRtHoldStride = kRtHold + 4;          // + P_0[lane 31] carry (two parities), P_2[lane 0], bias: see rt_epilogue_chunk
        constexpr uint32_t kRtBBytes = kRtN * kBlockC * 4;  // one filter stage
        constexpr uint32_t kRtHoldBytes = kTileH * kRtBNK * kRtHoldStride * 4; // [row][filter][8 columns + carry]
        constexpr uint32_t kRtDummyBytes = 8 * 32 * 4;      // one scratch word per epilogue thread
        constexpr int kRtFirstConvWarp = 4;
        constexpr int kRtFirstEpiWarp = 12;

        struct RowtapParams
        {
            int Cblocks, R, padX, padY, HR;
            int xStages, bStages;
            int Ho, Wo, K;
            int tilesW, tilesH, tilesK, numStrips;
            long long yStrideN, yStrideK;
            int act;
            float alpha;
            int dbgFlags; // profiling only (NB200_RT_DEBUG): 1 = epilogue skips math + stores, 2 = converters skip the loads, 4 = no global stores
        };

        __device__ __forceinline__ uint32_t rt_a_stage_col(uint32_t as) { return (as < 2 ? 192u : 448u - 64u) + as * 32u; }

        __device__ __forceinline__ float rt_activate(int act, float alpha, float f)
        {
            if (act == NB200_ACT_RELU) return f > 0.f ? f : 0.f;
            if (act == NB200_ACT_IDENTITY) return f;
            return apply_activation(act, alpha, f);
        }

        // One 16-filter chunk of the row-tap epilogue for this thread's lane (see the kernel). Lanes 0-23 store this tile's columns
        // 0-23; lanes 24-31 store the PREVIOUS tile's columns 24-31 (held in shared memory, same lane) and hold this tile's.
        // Per (tile row, filter) the warp owns 12 words of shared memory at rowAddr + 48 j:
        //     [0..7] held columns 24-31 | [8], [9] P0[lane 31] of the even / odd tiles (carry into the next tile's column 0)
        //     [10] P2[lane 0] of this tile (completes the previous tile's column 31) | [11] bias
        // The chunk is written as whole-array phases (16 independent shuffles / loads / stores each) so that no result is
        // consumed right after it is requested: the warp issues in order, and a per-filter chain of shuffle -> add -> load ->
        // store stalled on every link (ablation: the math alone cost 0.10 ms of the 0.41 ms C64->K64 layer).
        // Shared accesses are volatile asm, i.e. they stay in program order within the warp.
        template <int ACT, bool FULL>
        __device__ __forceinline__ void rt_epilogue_chunk(uint32_t (&p0)[16], uint32_t (&p1)[16], uint32_t (&p2)[16], uint32_t rowAddr, int lane,
                                                          bool first, uint32_t parity, float* yptr, bool storeOk, long long strideK, int kLeft,
                                                          int act, float alpha)
        {
            constexpr uint32_t kPitch = kRtHoldStride * 4;
            const bool upper = lane >= 32 - kRtHold;
            const uint32_t heldAddr = rowAddr + (uint32_t)((lane - (32 - kRtHold)) * 4);
            uint32_t t[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)                      // bias (broadcast load)
                t[j] = ptx::lds_b32(rowAddr + j * kPitch + 44);
            if (lane == 31)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)                  // this tile's P0[31] for the next tile
                    ptx::sts_b32(rowAddr + j * kPitch + 32 + 4 * parity, p0[j]);
            }
            if (lane == 0)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)                  // this tile's P2[0] for lane 31 below
                    ptx::sts_b32(rowAddr + j * kPitch + 40, p2[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
                p0[j] = __float_as_uint(__shfl_up_sync(0xffffffffu, __uint_as_float(p0[j]), 1));      // P0[i-1]
#pragma unroll
            for (int j = 0; j < 16; ++j)
                p2[j] = __float_as_uint(__shfl_down_sync(0xffffffffu, __uint_as_float(p2[j]), 1));    // P2[i+1]
            if (lane == 0)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)                  // column 0: the previous tile's P0[31] (zero padding for the first tile)
                    p0[j] = first ? 0u : ptx::lds_b32(rowAddr + j * kPitch + 32 + 4 * (parity ^ 1));
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                float o = (__uint_as_float(p0[j]) + __uint_as_float(p1[j])) + __uint_as_float(t[j]);
                if (lane < 31)
                    o += __uint_as_float(p2[j]);
                p1[j] = __float_as_uint(o);
            }
            if (upper)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)                  // the sector the previous tile held back
                    t[j] = ptx::lds_b32(heldAddr + j * kPitch);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    ptx::sts_b32(heldAddr + j * kPitch, p1[j]);
            }
            if (lane == 31)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    p2[j] = ptx::lds_b32(rowAddr + j * kPitch + 40);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
            {
                float val = __uint_as_float(upper ? t[j] : p1[j]);
                if (lane == 31)
                    val += __uint_as_float(p2[j]);
                if constexpr (ACT == NB200_ACT_RELU) val = val > 0.f ? val : 0.f;
                else if constexpr (ACT != NB200_ACT_IDENTITY) val = apply_activation(act, alpha, val);
                if (storeOk && (FULL || j < kLeft))
                    *yptr = val;
                yptr += strideK;
            }
        }

        __global__ void __launch_bounds__(kRtThreads, 1)
        tc_rowtap_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, RowtapParams p,
                         const float* __restrict__ bias, float* __restrict__ y)
        {
            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            const uint32_t xBytes = (uint32_t)(kBlockC * p.HR * kRtWB * 4);
            const uint32_t xBytesPad = (xBytes + 1023) & ~1023u;
            uint8_t* bRing = smem;
            uint8_t* xRing = smem + p.bStages * kRtBBytes;
            float* hold = (float*)(xRing + p.xStages * xBytesPad);
            uint64_t* bars = (uint64_t*)((uint8_t*)hold + kRtHoldBytes + kRtDummyBytes);
            uint64_t* bFull = bars;             // [bStages <= 8]
            uint64_t* bEmpty = bFull + 8;
            uint64_t* xFull = bEmpty + 8;       // [xStages <= 4]
            uint64_t* xEmpty = xFull + 4;
            uint64_t* aFull = xEmpty + 4;       // [4]
            uint64_t* aEmpty = aFull + 4;
            uint64_t* accFull = aEmpty + 4;     // [2]
            uint64_t* accEmpty = accFull + 2;
            uint32_t* tmemSlot = (uint32_t*)(accEmpty + 2);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapW);
                for (int s = 0; s < p.bStages; ++s) { ptx::mbar_init(&bFull[s], 1); ptx::mbar_init(&bEmpty[s], 1); }
                for (int s = 0; s < p.xStages; ++s) { ptx::mbar_init(&xFull[s], 1); ptx::mbar_init(&xEmpty[s], 8); }
                for (int s = 0; s < 4; ++s) { ptx::mbar_init(&aFull[s], 4); ptx::mbar_init(&aEmpty[s], 1); }
                for (int s = 0; s < 2; ++s) { ptx::mbar_init(&accFull[s], 1); ptx::mbar_init(&accEmpty[s], 8); }
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, 512);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            const uint32_t tmemBase = *tmemSlot;
            const uint32_t bFull32 = ptx::smem_u32(bFull), bEmpty32 = ptx::smem_u32(bEmpty), xFull32 = ptx::smem_u32(xFull),
                           xEmpty32 = ptx::smem_u32(xEmpty), aFull32 = ptx::smem_u32(aFull), aEmpty32 = ptx::smem_u32(aEmpty),
                           accFull32 = ptx::smem_u32(accFull), accEmpty32 = ptx::smem_u32(accEmpty);
            const int itersPerTile = p.Cblocks * p.R;

            if (warp == 0)
            {
                if (lane == 0)
                {
                    // ===== filter TMA: one 192-row tile per (channel block, filter row), again for every tile of the strip =====
                    int bs = 0;
                    uint32_t bph = 0;
                    for (int sp = blockIdx.x; sp < p.numStrips; sp += gridDim.x)
                    {
                        const int kt = sp % p.tilesK;
                        for (int tw = 0; tw < p.tilesW; ++tw)
                            for (int cb = 0; cb < p.Cblocks; ++cb)
                                for (int r = 0; r < p.R; ++r)
                                {
                                    ptx::mbar_wait(bEmpty32 + 8u * bs, bph ^ 1);
                                    ptx::mbar_arrive_expect_tx(bFull32 + 8u * bs, kRtBBytes);
                                    ptx::tma_load_3d(bRing + bs * kRtBBytes, &mapW, &bFull[bs], cb * kBlockC, kt * kRtN, r);
                                    if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                                }
                    }
                }
            }
            else if (warp == 2)
            {
                if (lane == 0)
                {
                    // ===== halo TMA: one tile per channel block (rows above / below the image read as zeros) =====
                    int xs = 0;
                    uint32_t xph = 0;
                    for (int sp = blockIdx.x; sp < p.numStrips; sp += gridDim.x)
                    {
                        int u = sp / p.tilesK;
                        const int th = u % p.tilesH;
                        const int n = u / p.tilesH;
                        for (int tw = 0; tw < p.tilesW; ++tw)
                        {
                            const int origin = tw * kTileW;
                            for (int cb = 0; cb < p.Cblocks; ++cb)
                            {
                                ptx::mbar_wait(xEmpty32 + 8u * xs, xph ^ 1);
                                ptx::mbar_arrive_expect_tx(xFull32 + 8u * xs, xBytes);
                                ptx::tma_load_4d(xRing + xs * xBytesPad, &mapX, &xFull[xs], origin, th * kTileH - p.padY, cb * kBlockC, n);
                                if (++xs == p.xStages) { xs = 0; xph ^= 1; }
                            }
                        }
                    }
                }
            }
            else if (warp == 1)
            {
                // ===== MMA issuer =====
                constexpr uint32_t idesc = ptx::idesc_tf32(128, kRtN, 0, 0);
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(bRing), 16, 1024);
                int bs = 0;
                uint32_t bph = 0, aCount = 0, tileCount = 0;
                for (int sp = blockIdx.x; sp < p.numStrips; sp += gridDim.x)
                    for (int tw = 0; tw < p.tilesW; ++tw, ++tileCount)
                    {
                        const uint32_t buf = tileCount & 1;
                        ptx::mbar_wait(accEmpty32 + 8u * buf, ((tileCount >> 1) & 1) ^ 1); // the epilogue has drained this accumulator
                        ptx::tc_fence_after_sync();
                        const uint32_t tmemAcc = tmemBase + buf * 256;
                        for (int it = 0; it < itersPerTile; ++it, ++aCount)
                        {
                            const uint32_t as = aCount & 3;
                            ptx::mbar_wait(bFull32 + 8u * bs, bph);
                            ptx::mbar_wait(aFull32 + 8u * as, (aCount >> 2) & 1);
                            ptx::tc_fence_after_sync();
                            if (ptx::elect_one())
                            {
                                const uint64_t db = descB0 + (uint64_t)((bs * kRtBBytes) >> 4);
                                const uint32_t ta = tmemBase + rt_a_stage_col(as);
#pragma unroll
                                for (int kk = 0; kk < kBlockC / 8; ++kk)
                                    ptx::mma_tf32_ts(tmemAcc, ta + kk * 8, db + kk * 2, idesc, (it | kk) != 0);
                                ptx::mma_commit(aEmpty32 + 8u * as);
                                ptx::mma_commit(bEmpty32 + 8u * bs);
                                if (it == itersPerTile - 1)
                                    ptx::mma_commit(accFull32 + 8u * buf);
                            }
                            __syncwarp();
                            if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                        }
                    }
            }
            else if (warp >= kRtFirstConvWarp && warp < kRtFirstEpiWarp)
            {
                // ===== converters: group g builds A tiles aCount = g (mod 2) into its own stages {g, g + 2} =====
                const int q = warp & 3;
                const int g = (warp - kRtFirstConvWarp) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                const uint32_t chanStrideB = (uint32_t)(p.HR * kRtWB * 4);
                const uint32_t xRing32 = ptx::smem_u32(xRing);
                bool pending = false;
                uint32_t pendStage = 0, aCount = 0, xCount = 0;
                for (int sp = blockIdx.x; sp < p.numStrips; sp += gridDim.x)
                    for (int tw = 0; tw < p.tilesW; ++tw)
                    {
                        for (int cb = 0; cb < p.Cblocks; ++cb, ++xCount)
                        {
                            const uint32_t xs = xCount % (uint32_t)p.xStages;
                            ptx::mbar_wait(xFull32 + 8u * xs, (xCount / (uint32_t)p.xStages) & 1);
                            const uint32_t xt = xRing32 + xs * xBytesPad;
                            for (int r = 0; r < p.R; ++r, ++aCount)
                            {
                                if ((int)(aCount & 1) != g)
                                    continue;
                                const uint32_t src = xt + (uint32_t)(((q + r) * kRtWB + lane) << 2);
                                uint32_t v[kBlockC];
                                if (p.dbgFlags & 2)
                                {
#pragma unroll
                                    for (int c = 0; c < kBlockC; ++c)
                                        v[c] = 0x3f800000u;
                                }
                                else if (chanStrideB == 768u) // 3 filter rows: compile-time channel pitch -> immediate offsets
                                {
#pragma unroll
                                    for (int c = 0; c < kBlockC; ++c)
                                        v[c] = ptx::tf32_round_bits(ptx::lds_b32(src + c * 768u));
                                }
                                else
                                {
#pragma unroll
                                    for (int c = 0; c < kBlockC; ++c)
                                        v[c] = ptx::tf32_round_bits(ptx::lds_b32(src + c * chanStrideB));
                                }
                                if (pending)
                                {
                                    ptx::tmem_st_wait();
                                    ptx::tc_fence_before_sync();
                                    __syncwarp();
                                    if (lane == 0)
                                        ptx::mbar_arrive(aFull32 + 8u * pendStage);
                                }
                                const uint32_t as = aCount & 3;
                                ptx::mbar_wait(aEmpty32 + 8u * as, ((aCount >> 2) & 1) ^ 1);
                                ptx::tc_fence_after_sync();
                                ptx::tmem_st_32x32b_x32(tmemBase + laneSel + rt_a_stage_col(as), v);
                                pending = true;
                                pendStage = as;
                            }
                            __syncwarp();
                            if (lane == 0)
                                ptx::mbar_arrive(xEmpty32 + 8u * xs);
                        }
                    }
                if (pending)
                {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(aFull32 + 8u * pendStage);
                }
            }
            else if (warp >= kRtFirstEpiWarp)
            {
                // ===== epilogue: warp (q, h) owns tile row q and filters [32 h, 32 h + 32) of the tile =====
                // Per filter:  o[i]   = P0[i-1] + P1[i] + P2[i+1] + bias   (P0[-1] = carry of the previous tile; lane 31 lacks P2[32])
                //              store  lanes 0-23: this tile's columns 0-23; lanes 24-31: the sector held back by the previous tile
                //                     (same lanes; + P2[0] of this tile into its column 31) -- one 128-byte span per instruction
                //              hold   this tile's columns 24-31 and P0[31] for the next tile (after the last tile the sector is
                //                     flushed as it is: the missing term multiplies zero padding)
                const int q = warp & 3;
                const int h = (warp - kRtFirstEpiWarp) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                const bool upper = lane >= 32 - kRtHold;
                const uint32_t myHold = ptx::smem_u32(hold) + (uint32_t)((q * kRtBNK + h * 32) * kRtHoldStride * 4);
                const uint32_t holdLane = myHold + (uint32_t)((lane - (32 - kRtHold)) * 4); // meaningful for the upper lanes only
                uint32_t tileCount = 0;
                for (int sp = blockIdx.x; sp < p.numStrips; sp += gridDim.x)
                {
                    int u = sp;
                    const int kt = u % p.tilesK; u /= p.tilesK;
                    const int th = u % p.tilesH;
                    const int n = u / p.tilesH;
                    const int k0 = kt * kRtBNK + h * 32;
                    const int oh = th * kTileH + q;
                    const bool rowOk = oh < p.Ho;
                    float* yrow = y + n * p.yStrideN + (long long)oh * p.Wo + (long long)k0 * p.yStrideK;
                    // this strip's 32 biases into the warp's rows (word 11 of each filter's 12)
                    ptx::sts_b32(myHold + (uint32_t)(lane * kRtHoldStride * 4) + 44,
                                 __float_as_uint((bias && k0 + lane < p.K) ? __ldg(bias + k0 + lane) : 0.f));
                    __syncwarp();
                    for (int tw = 0; tw < p.tilesW; ++tw, ++tileCount)
                    {
                        const uint32_t buf = tileCount & 1;
                        const uint32_t acc = tmemBase + laneSel + buf * 256 + h * 32;
                        const int col = (upper ? tw - 1 : tw) * kTileW + lane;        // output column this lane stores
                        const bool storeOk = rowOk && col >= 0 && col < p.Wo && !(p.dbgFlags & 4);
                        ptx::mbar_wait(accFull32 + 8u * buf, (tileCount >> 1) & 1);
                        ptx::tc_fence_after_sync();
#pragma unroll 1
                        for (int c0 = 0; c0 < 32; c0 += 16)
                        {
                            const int kb = k0 + c0;
                            uint32_t p0[16], p1[16], p2[16];
                            ptx::tmem_ld_32x32b_x16(acc + c0, p0);
                            ptx::tmem_ld_32x32b_x16(acc + kRtBNK + c0, p1);
                            ptx::tmem_ld_32x32b_x16(acc + 2 * kRtBNK + c0, p2);
                            ptx::tmem_ld_wait();
                            if (c0 == 16)
                            {
                                // every column this warp needs has been read: hand the accumulator back to the MMA warp
                                ptx::tc_fence_before_sync();
                                __syncwarp();
                                if (lane == 0)
                                    ptx::mbar_arrive(accEmpty32 + 8u * buf);
                            }
                            if (p.dbgFlags & 1)
                                continue;
                            float* yptr = yrow + (long long)c0 * p.yStrideK + col;
                            const uint32_t ra = myHold + (uint32_t)(c0 * kRtHoldStride * 4);
                            const int kLeft = p.K - kb;
                            if (kLeft >= 16 && p.act == NB200_ACT_RELU)
                                rt_epilogue_chunk<NB200_ACT_RELU, true>(p0, p1, p2, ra, lane, tw == 0, (uint32_t)tw & 1, yptr, storeOk, p.yStrideK, kLeft, p.act, p.alpha);
                            else if (kLeft >= 16 && p.act == NB200_ACT_IDENTITY)
                                rt_epilogue_chunk<NB200_ACT_IDENTITY, true>(p0, p1, p2, ra, lane, tw == 0, (uint32_t)tw & 1, yptr, storeOk, p.yStrideK, kLeft, p.act, p.alpha);
                            else
                                rt_epilogue_chunk<-1, false>(p0, p1, p2, ra, lane, tw == 0, (uint32_t)tw & 1, yptr, storeOk, p.yStrideK, kLeft, p.act, p.alpha);
                        }
                    }
                    // flush the last tile's held sector (its column 31 lacks only a term that multiplies zero padding)
                    if (!(p.dbgFlags & 1))
                    {
                        const int col = (p.tilesW - 1) * kTileW + lane;
                        if (upper && rowOk && col < p.Wo)
                        {
                            for (int j = 0; j < 32; ++j)
                                if (k0 + j < p.K)
                                    yrow[(long long)j * p.yStrideK + col] =
                                        rt_activate(p.act, p.alpha, __uint_as_float(ptx::lds_b32(holdLane + (uint32_t)(j * kRtHoldStride * 4))));
                        }
                    }
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemBase, 512);
            }
        }

        // filters for tc_rowtap_kernel: out[r][kt][s][kLocal 0..63][c] (TF32-rounded, zero padded); mode 0 forward, mode 1 input
        // gradient (filters flipped, roles of K and C exchanged: GEMM filter index = input channel, reduction = K)
        __global__ void repack_rowtap_kernel(const float* __restrict__ w, float* __restrict__ out, int K, int C, int R, int S, int ktiles,
                                             int outCp, int mode)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            const int rowsPerR = ktiles * S * kRtBNK;
            const long long total = (long long)R * rowsPerR * outCp;
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
            {
                const int col = (int)(i % outCp);
                const int row = (int)((i / outCp) % rowsPerR);
                const int r = (int)(i / ((long long)outCp * rowsPerR));
                const int kt = row / (S * kRtBNK), s = (row / kRtBNK) % S, kl = row % kRtBNK;
                const int filt = kt * kRtBNK + kl;
                float v = 0.f;
                if (mode == 0)
                {
                    if (filt < K && col < C)
                        v = w[(((long long)filt * C + col) * R + r) * S + s];
                }
                else
                {
                    if (filt < C && col < K)
                        v = w[(((long long)col * C + filt) * R + (R - 1 - r)) * S + (S - 1 - s)];
                }
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
                out[i] = __uint_as_float(t);
            }
        }

        // ---------------------------------------------------------------- forward kernel, CTA-pair form (cta_group::2)
        // Same algorithm, but two CTAs of a cluster (one TPC = two SMs) cooperate on a 256-pixel x BN-filter tile:
        //   - each CTA converts the A tile of ITS 128 pixels (4 of the pair's 8 output rows) into its own TMEM;
        //   - each CTA loads only HALF of every filter tile (BN/2 filters); the pair leader issues
        //     tcgen05.mma.cta_group::2 (M = 256), which reads B halves from both shared memories;
        //   - accumulator rows 0-127 / 128-255 stay in the leader's / peer's TMEM, so each CTA runs its own epilogue.
        // Per MMA cycle this halves the filter bytes written by TMA and read by the tensor core in every SM, which is
        // what bounds the single-CTA kernel (ncu: tensor pipe ~30 %, shared-memory pipes saturated).
        // Cross-CTA signalling: filter TMA loads of both CTAs complete_tx on the LEADER's bFull barrier; converter warps
        // of both CTAs arrive on the LEADER's aFull barrier (remote arrive through mapa); the leader's
        // tcgen05.commit multicasts to the bEmpty / aEmpty / accBar barriers of BOTH CTAs.
        template <int BN>
        __global__ void __launch_bounds__(kFpropThreads, (BN > 128 ? 1 : 2))
        tc_fprop2_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW, FpropParams p,
                         const float* __restrict__ bias, float* __restrict__ y)
        {
            constexpr int BNH = BN / 2;
            constexpr int kAStages = a_stages(BN);
            constexpr uint32_t kBBytes = BNH * kBlockC * 4;
            constexpr uint32_t kTmemCols = BN > 128 ? 512 : 256;
            static_assert(BN + kAStages * kBlockC <= kTmemCols, "TMEM budget");

            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            const uint32_t xBytes = (uint32_t)(kBlockC * p.HR * p.WB * 4);
            const uint32_t xBytesPad = (xBytes + 1023) & ~1023u;
            uint8_t* bRing = smem;
            uint8_t* xRing = smem + p.bStages * kBBytes;
            uint64_t* bars = (uint64_t*)(xRing + p.xStages * xBytesPad);
            uint64_t* bFull = bars;            // [bStages]   (leader's copy is the live one)
            uint64_t* bEmpty = bFull + 8;      // [bStages]
            uint64_t* xFull = bEmpty + 8;      // [xStages]   local
            uint64_t* xEmpty = xFull + 4;      // [xStages]   local
            uint64_t* aFull = xEmpty + 4;      // [kAStages]  (leader's copy is the live one)
            uint64_t* aEmpty = aFull + kAStagesMax;
            uint64_t* accBar = aEmpty + kAStagesMax;
            uint32_t* tmemSlot = (uint32_t*)(accBar + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;
            const uint32_t rank = ptx::cluster_ctarank();
            const bool leader = rank == 0;

            int t = blockIdx.x >> 1; // pair index; filter tile fastest (L2 reuse of the activation tile)
            const int kt = t % p.tilesK; t /= p.tilesK;
            const int tw = t % p.tilesW; t /= p.tilesW;
            const int th = t % p.tilesH; t /= p.tilesH;
            const int n = t;
            const int ow0 = tw * kTileW, oh0 = th * (2 * kTileH) + (int)rank * kTileH, k0 = kt * BN;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapW);
                for (int s = 0; s < p.bStages; ++s) { ptx::mbar_init(&bFull[s], 1); ptx::mbar_init(&bEmpty[s], 1); }
                for (int s = 0; s < p.xStages; ++s) { ptx::mbar_init(&xFull[s], 1); ptx::mbar_init(&xEmpty[s], kTileH * kConvGroups); }
                for (int s = 0; s < kAStages; ++s) { ptx::mbar_init(&aFull[s], 2 * kTileH); ptx::mbar_init(&aEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc_2sm(tmemSlot, kTmemCols);
            ptx::tc_fence_before_sync();
            ptx::cluster_sync(); // both CTAs' barriers are initialised before any remote arrive / TMA signal
            ptx::tc_fence_after_sync();
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t tmemA = tmemAcc + BN;

            const int taps = p.R * p.S;
            // per-CTA debug rows: [0] filter producer, [1] MMA issuer, [2] halo producer, [3] converter warp 3 lane 0
            long long* dbgBase = p.dbg ? p.dbg + (long long)blockIdx.x * 4 * kDbgSlots : nullptr;
            WaitAcc dbgP(dbgBase ? dbgBase : nullptr);
            WaitAcc dbgM(dbgBase ? dbgBase + kDbgSlots : nullptr);
            WaitAcc dbgX(dbgBase ? dbgBase + 2 * kDbgSlots : nullptr);
            WaitAcc dbgC((dbgBase && warp == kFirstConvWarp) ? dbgBase + 3 * kDbgSlots : nullptr);
            const long long tStart = clock64();

            if (warp == 0)
            {
                if (lane == 0)
                {
                    // ===== TMA producer (filters): this CTA's half of each tile; bytes are counted on the leader's barrier =====
                    int bs = 0;
                    uint32_t bph = 0;
                    for (int cb = 0; cb < p.Cblocks; ++cb)
                        for (int tap = 0; tap < taps; ++tap)
                        {
                            timed_wait(&bEmpty[bs], bph ^ 1, dbgP, kDbgBEmpty);
                            if (leader)
                                ptx::mbar_arrive_expect_tx(&bFull[bs], 2 * kBBytes);
                            ptx::tma_load_3d_2sm(bRing + bs * kBBytes, &mapW, ptx::mapa_u32(&bFull[bs], 0), cb * kBlockC, k0 + (int)rank * BNH, tap);
                            if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                        }
                }
            }
            else if (warp == 2)
            {
                if (lane == 0)
                {
                    // ===== TMA producer (activations): this CTA's halo tile per channel block =====
                    int xs = 0;
                    uint32_t xph = 0;
                    for (int cb = 0; cb < p.Cblocks; ++cb)
                    {
                        timed_wait(&xEmpty[xs], xph ^ 1, dbgX, kDbgXEmpty);
                        ptx::mbar_arrive_expect_tx(&xFull[xs], xBytes);
                        ptx::tma_load_4d(xRing + xs * xBytesPad, &mapX, &xFull[xs], ow0 - p.wOff, oh0 - p.padY, cb * kBlockC, n);
                        if (++xs == p.xStages) { xs = 0; xph ^= 1; }
                    }
                }
            }
            else if (warp == 1)
            {
                if (leader)
                {
                    // ===== MMA issuer (pair leader): D[256 x BN] += A[tmem of both CTAs] * B[halves in both smem]^T =====
                    constexpr uint32_t idesc = ptx::idesc_tf32(256, BN, 0, 0);
                    const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(bRing), /*LBO*/ 16, /*SBO*/ 1024);
                    int as = 0, bs = 0;
                    uint32_t aph = 0, bph = 0;
                    const int iters = taps * p.Cblocks;
                    for (int it = 0; it < iters; ++it)
                    {
                        timed_wait(&bFull[bs], bph, dbgM, kDbgBFull);
                        timed_wait(&aFull[as], aph, dbgM, kDbgAFull);
                        ptx::tc_fence_after_sync();
                        if (ptx::elect_one())
                        {
                            const uint64_t db = descB0 + (uint64_t)((bs * kBBytes) >> 4);
                            const uint32_t ta = tmemA + as * kBlockC;
#pragma unroll
                            for (int kk = 0; kk < kBlockC / 8; ++kk)
                                ptx::mma_tf32_ts_2sm(tmemAcc, ta + kk * 8, db + kk * 2, idesc, (it | kk) != 0);
                            ptx::mma_commit_2sm(&aEmpty[as], 3);
                            ptx::mma_commit_2sm(&bEmpty[bs], 3);
                        }
                        __syncwarp();
                        if (++as == kAStages) { as = 0; aph ^= 1; }
                        if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                    }
                    if (ptx::elect_one())
                        ptx::mma_commit_2sm(accBar, 3);
                    __syncwarp();
                }
            }
            else
            {
                // ===== converters (then epilogue), as in the single-CTA kernel; aFull lives in the leader =====
                const int q = warp & 3;
                const int g = (warp - kFirstConvWarp) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                const uint32_t chanStrideB = (uint32_t)(p.HR * p.WB * 4);
                const uint32_t xRing32 = ptx::smem_u32(xRing);
                bool pending = false;
                int pendStage = 0;
                for (int cb = 0; cb < p.Cblocks; ++cb)
                {
                    const int xs = cb % p.xStages;
                    timed_wait(&xFull[xs], (uint32_t)(cb / p.xStages) & 1, dbgC, kDbgXFull);
                    const uint32_t xt = xRing32 + xs * xBytesPad;
                    int r = 0, s = 0;
                    for (int tap = 0; tap < taps; ++tap, (++s == p.S ? (s = 0, ++r) : 0))
                    {
                        const int it = cb * taps + tap;
                        if ((it & 1) != g)
                            continue;
                        const uint32_t src = xt + (uint32_t)(((q + r) * p.WB + (lane + s - p.padX + p.wOff)) << 2);
                        uint32_t v[kBlockC];
                        if (chanStrideB == 960u) // 3x3 filters: compile-time channel pitch -> LDS with immediate offsets
                        {
#pragma unroll
                            for (int c = 0; c < kBlockC; ++c)
                                v[c] = ptx::tf32_round_bits(ptx::lds_b32(src + c * 960u));
                        }
                        else
                        {
#pragma unroll
                            for (int c = 0; c < kBlockC; ++c)
                                v[c] = ptx::tf32_round_bits(ptx::lds_b32(src + c * chanStrideB));
                        }
                        if (pending)
                        {
                            ptx::tmem_st_wait();
                            ptx::tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0)
                                ptx::mbar_arrive_cluster(ptx::mapa_u32(&aFull[pendStage], 0));
                        }
                        const int as = it & (kAStages - 1);
                        timed_wait(&aEmpty[as], ((uint32_t)(it / kAStages) & 1) ^ 1, dbgC, kDbgAEmpty);
                        ptx::tc_fence_after_sync();
                        ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * kBlockC, v);
                        pending = true;
                        pendStage = as;
                    }
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(&xEmpty[xs]);
                }
                if (pending)
                {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive_cluster(ptx::mapa_u32(&aFull[pendStage], 0));
                }

                // ----- epilogue: this CTA's 128 accumulator rows -----
                const int oh = oh0 + q, ow = ow0 + lane;
                timed_wait(accBar, 0, dbgC, kDbgAcc);
                ptx::tc_fence_after_sync();
                const bool pixelOk = oh < p.Ho && ow < p.Wo;
                float* yp = y + n * p.yStrideN + (long long)oh * p.Wo + ow;
#pragma unroll 1
                for (int c0 = g * 32; c0 < BN; c0 += 32 * kConvGroups)
                {
                    if (k0 + c0 >= p.K)
                        break; // warp-uniform
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + c0, v);
                    ptx::tmem_ld_wait();
                    nb200::store_chunk(yp, p.yStrideK, k0 + c0, p.K, bias, lane, p.act, p.alpha, pixelOk, v);
                }
            }

            if (warp == 0) dbgP.flush();
            if (warp == 1) dbgM.flush();
            if (warp == 2) dbgX.flush();
            if (warp == kFirstConvWarp) dbgC.flush();
            if (dbgBase && threadIdx.x == 0)
                dbgBase[kDbgTotal] = clock64() - tStart;
            // neither CTA may exit (or free TMEM) while the other can still signal its barriers or read its filter halves
            ptx::tc_fence_before_sync();
            ptx::cluster_sync();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc_2sm(tmemAcc, kTmemCols);
            }
        }


        // ---------------------------------------------------------------- gather kernel (any stride / padding / map size)
        // Same MMA pipeline as tc_fprop_kernel (filters by TMA, A operand converted into tensor memory, TS-form
        // tcgen05.mma, fused epilogue) but the A tile is GATHERED: the 128 rows of a tile are 128 consecutive entries of a
        // flattened pixel list (n, a, b), and for every filter tap each converter thread computes its own input address
        //      iy = a * iyMul + iyAdd[tap],   ix = b * ixMul + ixAdd[tap]        (out of range -> 0)
        // and loads straight from global memory / L1 (coalesced along b). This covers what the halo-tile kernel cannot:
        // strided forward convolutions, feature maps narrower than a 32-pixel row tile (the batch folds into M), widths
        // that are not multiples of 4, and -- through one launch per stride-parity class of output pixels -- the input
        // gradient of strided convolutions (= the forward of Conv2DTranspose), where each class sees only its own taps.
        struct GatherParams
        {
            int Cblocks, ntaps;
            int ntapsA;                     // class pair (tc_gather_kernel<BN, true>): taps [0, ntapsA) feed the first accumulator
            int tapsAll;                    // R * S of the filter (3xTF32: the lo parts start at this tap index of the repacked tensor)
            int C, H, W;                    // gathered tensor: channels, rows, cols
            int K, Ho, Wo;                  // produced tensor: channels, rows, cols
            int PH, PW;                     // pixel list extent per image
            int oyMul, oyAdd, oxMul, oxAdd; // produced pixel of list entry (a, b)
            int iyMul, ixMul;
            long long totalPix;             // N * PH * PW
            int tilesK, bStages;
            int act;
            float alpha;
            // channel split (few tiles, many channels: the weight-bound U-Net bottleneck layers): split sp of a tile reduces
            // channel blocks [sp*cbPer, (sp+1)*cbPer) and writes a raw partial; fprop_split_reduce_kernel adds them
            int splits, cbPer;
            float* partial;
            long long partialStride;
            short iyAdd[32], ixAdd[32], wtap[32];
            int smemSlack; // bytes the launch reserved for aligning the dynamic shared memory base to 1 KB
            int dbgFlags; // NB200_GATHER_DEBUG ablations (profiling only): 1 no gather loads, 2 no output stores, 4 empty body, 8 no TMEM either
        };

        // Several pixel lists served by ONE launch: the stride^2 parity classes of a strided input gradient (transposed
        // convolution) differ only in their tap lists and pixel-list geometry, and each alone fills a fraction of the chip.
        constexpr int kGatherMaxClasses = 9;
        struct GatherBatch
        {
            int count;
            int tileStart[kGatherMaxClasses + 1]; // first CTA of each class; tileStart[count] = grid size
            GatherParams cls[kGatherMaxClasses];
        };

        // float2 variant of store_chunk for a class PAIR (see tc_gather_kernel): v0 / v1 are the two horizontally adjacent
        // output pixels (2b, 2b + 1) of this thread, so 32 lanes write 256 contiguous bytes per channel -- whole sectors,
        // where the two classes launched separately each wrote every other float of every sector.
        template <int ACT>
        __device__ __forceinline__ void store_chunk2_t(float* yp, long long strideK, int kBase, int K, float biasLane, int act, float alpha,
                                                       bool pixelOk, const uint32_t (&v0)[32], const uint32_t (&v1)[32])
        {
#pragma unroll
            for (int j = 0; j < 32; ++j)
            {
                const float b = __shfl_sync(0xffffffffu, biasLane, j);
                if (pixelOk && kBase + j < K)
                {
                    float2 f = make_float2(__uint_as_float(v0[j]) + b, __uint_as_float(v1[j]) + b);
                    if constexpr (ACT == NB200_ACT_RELU) { f.x = f.x > 0.f ? f.x : 0.f; f.y = f.y > 0.f ? f.y : 0.f; }
                    else if constexpr (ACT == NB200_ACT_LEAKY_RELU) { f.x = f.x >= 0.f ? f.x : alpha * f.x; f.y = f.y >= 0.f ? f.y : alpha * f.y; }
                    else if constexpr (ACT != NB200_ACT_IDENTITY) { f.x = apply_activation(act, alpha, f.x); f.y = apply_activation(act, alpha, f.y); }
                    *reinterpret_cast<float2*>(yp + (long long)(kBase + j) * strideK) = f;
                }
            }
        }

        __device__ __forceinline__ void store_chunk2(float* yp, long long strideK, int kBase, int K, const float* __restrict__ bias, int lane,
                                                     int act, float alpha, bool pixelOk, const uint32_t (&v0)[32], const uint32_t (&v1)[32])
        {
            const float bl = (bias && kBase + lane < K) ? __ldg(bias + kBase + lane) : 0.f;
            switch (act)
            {
            case NB200_ACT_IDENTITY: store_chunk2_t<NB200_ACT_IDENTITY>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v0, v1); break;
            case NB200_ACT_RELU: store_chunk2_t<NB200_ACT_RELU>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v0, v1); break;
            case NB200_ACT_LEAKY_RELU: store_chunk2_t<NB200_ACT_LEAKY_RELU>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v0, v1); break;
            default: store_chunk2_t<-1>(yp, strideK, kBase, K, bl, act, alpha, pixelOk, v0, v1); break;
            }
        }

        // (A version that staged the gathers through per-thread cp.async FIFOs in shared memory -- 2-4 iterations in flight per
        // converter group instead of one -- was parity-clean but no faster: the gathers are bound by L1 wavefronts / LSU issue,
        // not by latency, and the read-back doubled the LSU work: 42.6 vs 44.2 us on a 36-iteration full-wave layer, and 33.9
        // vs 17.3 us with loads and stores ablated; profiles/r2_gather_ablate.txt.)
        template <int BN, bool PAIR, bool X3>
        struct GatherCfg
        {
            static_assert(!(PAIR && X3), "3xTF32 drains ONE accumulator per channel block");
            static_assert(!X3 || BN <= 128, "3xTF32 keeps the running sums of BN / 2 filters in registers");
            static constexpr bool kOnePerSm = BN > 128 || PAIR || X3;
            static constexpr uint32_t kTmemCols = kOnePerSm ? 512 : 256;
            static constexpr int kACols = X3 ? 2 * kBlockC : kBlockC;       // [hi | lo] A tiles
            static constexpr int kAStages = X3 ? 4 : kOnePerSm ? (PAIR && BN > 64 ? 4 : 8) : 4;
            static_assert((PAIR ? 2 : 1) * BN + kAStages * kACols <= (int)kTmemCols, "TMEM budget");
        };

        // PAIR: the CTA computes TWO pixel classes of a stride-2 input gradient (transposed convolution) for the same 128 list
        // entries -- output columns 2b and 2b + 1 of row 2a + ph -- one after the other into two accumulators (taps [0, ntapsA)
        // belong to the first class) and its epilogue interleaves them.
        // X3 (NB200_MATH_3XTF32): operands split into hi + lo TF32 parts (filters in the repack, activations here), three MMAs per
        // K slice, and the accumulation chain cut per channel block exactly as in tc_fprop_kernel<BN, true>.
        template <int BN, bool PAIR, bool X3>
        __global__ void __launch_bounds__(kFpropThreads, (GatherCfg<BN, PAIR, X3>::kOnePerSm ? 1 : 2))
        tc_gather_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ GatherBatch batch, const float* __restrict__ in,
                         const float* __restrict__ bias, float* __restrict__ out)
        {
            using Cfg = GatherCfg<BN, PAIR, X3>;
            int cls = 0;
            while (cls + 1 < batch.count && (int)blockIdx.x >= batch.tileStart[cls + 1])
                ++cls;
            const GatherParams& p = batch.cls[cls];
            const int bid = (int)blockIdx.x - batch.tileStart[cls];
            constexpr uint32_t kBTile = BN * kBlockC * 4;
            constexpr uint32_t kBBytes = kBTile * (X3 ? 2 : 1);
            constexpr int kAStages = Cfg::kAStages;
            constexpr int kACols = Cfg::kACols;
            constexpr uint32_t kTmemCols = Cfg::kTmemCols;

            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            uint8_t* bRing = smem;
            uint64_t* bars = (uint64_t*)(smem + p.bStages * kBBytes);
            if ((uint32_t)(smem - smemRaw) > (uint32_t)p.smemSlack)
            {
                // two CTAs per SM leave no room for a full kilobyte of alignment slack (gather_smem): the launch relies on the
                // dynamic shared memory window starting (nearly) 1 KB aligned, which holds for a kernel without static shared memory
                if (threadIdx.x == 0)
                    printf("nb200: tc_gather_kernel shared memory base misaligned by %u bytes (slack %d)\n", (uint32_t)(smem - smemRaw), p.smemSlack);
                __trap();
            }
            uint64_t* bFull = bars;
            uint64_t* bEmpty = bFull + 8;
            uint64_t* aFull = bEmpty + 8;
            uint64_t* aEmpty = aFull + kAStagesMax;
            uint64_t* accBar = aEmpty + kAStagesMax;
            uint64_t* accFree = accBar + 1;    // 3xTF32 only: the accumulator has been drained into registers
            uint32_t* tmemSlot = (uint32_t*)(accFree + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;
            const int split = bid % p.splits;
            const int kt = (bid / p.splits) % p.tilesK;
            const int tile = bid / (p.splits * p.tilesK);
            const int k0 = kt * BN;
            const int cbBegin = split * p.cbPer;
            const int cbEnd = min(p.Cblocks, cbBegin + p.cbPer);
            const int dbgFlags = p.dbgFlags;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapW);
                for (int s = 0; s < p.bStages; ++s) { ptx::mbar_init(&bFull[s], 1); ptx::mbar_init(&bEmpty[s], 1); }
                for (int s = 0; s < kAStages; ++s) { ptx::mbar_init(&aFull[s], kTileH); ptx::mbar_init(&aEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::mbar_init(accFree, kTileH * kConvGroups);
                ptx::fence_mbar_init();
            }
            if (warp == 1 && !(dbgFlags & 8))
                ptx::tmem_alloc(tmemSlot, kTmemCols);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents(); // the next kernel's prologue may overlap this kernel (sm100_ptx.cuh)
            ptx::pdl_wait();              // nothing below runs before the previous kernel's memory is visible
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t tmemA = tmemAcc + (PAIR ? 2 : 1) * BN;
            const int nCb = max(cbEnd - cbBegin, 0);
            const int iters = p.ntaps * nCb;
            const int ntapsA = PAIR ? p.ntapsA : p.ntaps;

            if (dbgFlags & 12)
            {
                // profiling: launch + prologue + teardown only
            }
            else if (warp == 0)
            {
                if (lane == 0)
                {
                    int bs = 0;
                    uint32_t bph = 0;
                    for (int cb = cbBegin; cb < cbEnd; ++cb)
                        for (int t = 0; t < p.ntaps; ++t)
                        {
                            ptx::mbar_wait(&bEmpty[bs], bph ^ 1);
                            ptx::mbar_arrive_expect_tx(&bFull[bs], kBBytes);
                            ptx::tma_load_3d(bRing + bs * kBBytes, &mapW, &bFull[bs], cb * kBlockC, k0, p.wtap[t]);
                            if (X3) // lo parts live behind the hi parts in the repacked tensor (tap index + all taps of the filter)
                                ptx::tma_load_3d(bRing + bs * kBBytes + kBTile, &mapW, &bFull[bs], cb * kBlockC, k0, p.tapsAll + p.wtap[t]);
                            if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                        }
                }
            }
            else if (warp == 1)
            {
                constexpr uint32_t idesc = ptx::idesc_tf32(128, BN, 0, 0);
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(bRing), 16, 1024);
                int as = 0, bs = 0, t = 0, cbc = 0;
                uint32_t aph = 0, bph = 0;
                bool first0 = true, first1 = true;   // the first MMA into an accumulator overwrites it
                for (int it = 0; it < iters; ++it)
                {
                    ptx::mbar_wait(&bFull[bs], bph);
                    ptx::mbar_wait(&aFull[as], aph);
                    if (X3 && t == 0 && cbc > 0)
                        ptx::mbar_wait(accFree, (uint32_t)(cbc - 1) & 1); // the previous block's partial sums are in registers
                    ptx::tc_fence_after_sync();
                    const bool second = PAIR && t >= ntapsA;
                    const bool fresh = X3 ? t == 0 : second ? first1 : first0;
                    if (ptx::elect_one())
                    {
                        const uint64_t db = descB0 + (uint64_t)((bs * kBBytes) >> 4);
                        const uint32_t ta = tmemA + as * kACols;
                        const uint32_t acc = tmemAcc + (second ? BN : 0);
#pragma unroll
                        for (int kk = 0; kk < kBlockC / 8; ++kk)
                        {
                            ptx::mma_tf32_ts(acc, ta + kk * 8, db + kk * 2, idesc, !(fresh && kk == 0));                       // hi * hi
                            if (X3)
                            {
                                ptx::mma_tf32_ts(acc, ta + kk * 8, db + (kBTile >> 4) + kk * 2, idesc, 1);                     // hi * lo
                                ptx::mma_tf32_ts(acc, ta + kBlockC + kk * 8, db + kk * 2, idesc, 1);                           // lo * hi
                            }
                        }
                        ptx::mma_commit(&aEmpty[as]);
                        ptx::mma_commit(&bEmpty[bs]);
                        if (X3 && t == p.ntaps - 1)
                            ptx::mma_commit(accBar); // this channel block's partial accumulator is complete
                    }
                    __syncwarp();
                    if (second) first1 = false; else first0 = false;
                    if (++t == p.ntaps) { t = 0; ++cbc; }
                    if (++as == kAStages) { as = 0; aph ^= 1; }
                    if (++bs == p.bStages) { bs = 0; bph ^= 1; }
                }
                if (!X3 && ptx::elect_one())
                    ptx::mma_commit(accBar);
                __syncwarp();
            }
            else if (warp >= kFirstConvWarp)
            {
                const int q = warp & 3;
                const int g = (warp - kFirstConvWarp) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                // this thread's entry of the pixel list (32-bit arithmetic: lists beyond 2^31 entries are rejected on the host)
                const int pix = tile * 128 + q * 32 + lane;
                const bool pixOk = pix < (int)p.totalPix;
                const int b = pix % p.PW;
                const int a = (pix / p.PW) % p.PH;
                const int n = pix / (p.PW * p.PH);
                const long long plane = (long long)p.H * p.W;
                const float* inN = in + (long long)n * p.C * plane;

                bool pending = false;
                int pendStage = 0;
                // 3xTF32: running sums of this warp's filter chunks (see tc_fprop_kernel)
                constexpr int kOwnChunks = X3 ? (BN / (32 * kConvGroups)) : 1;
                float racc[kOwnChunks][32];
                if (X3)
                {
#pragma unroll
                    for (int ch = 0; ch < kOwnChunks; ++ch)
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            racc[ch][j] = 0.f;
                }
                auto publish = [&]() {
                    if (pending)
                    {
                        ptx::tmem_st_wait();
                        ptx::tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0)
                            ptx::mbar_arrive(&aFull[pendStage]);
                        pending = false;
                    }
                };
                int it = 0;
                for (int cbRel = 0; cbRel < nCb; ++cbRel)
                {
                    const int cbase = (cbBegin + cbRel) * kBlockC;
                    for (int t = 0; t < p.ntaps; ++t, ++it)
                    {
                        if ((it & 1) != g)
                            continue;
                        const int iy = a * p.iyMul + p.iyAdd[t], ix = b * p.ixMul + p.ixAdd[t];
                        const bool ok = pixOk && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && !(dbgFlags & 1);
                        uint32_t v[kBlockC];
                        if (ok)
                        {
                            const float* src = inN + cbase * plane + (long long)iy * p.W + ix;
                            if (cbase + kBlockC <= p.C)
                            {
#pragma unroll
                                for (int c = 0; c < kBlockC; ++c)
                                    v[c] = ptx::tf32_round_bits(__float_as_uint(__ldg(src + c * plane)));
                            }
                            else
                            {
#pragma unroll
                                for (int c = 0; c < kBlockC; ++c)
                                    v[c] = cbase + c < p.C ? ptx::tf32_round_bits(__float_as_uint(__ldg(src + c * plane))) : 0x1000u;
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int c = 0; c < kBlockC; ++c)
                                v[c] = 0x1000u; // tf32_round_bits(0): the tensor core reads the upper 19 bits only
                        }
                        uint32_t vlo[kBlockC]; // only live in the 3xTF32 instantiation
                        if (X3)
                        {
#pragma unroll
                            for (int c = 0; c < kBlockC; ++c)
                            {
                                const float full = __uint_as_float(v[c] - 0x1000u);           // undo the rounding add: the raw fp32 value
                                const uint32_t hi = v[c] & 0xFFFFE000u;                         // exactly what the tensor core will read
                                vlo[c] = ptx::tf32_round_bits(__float_as_uint(full - __uint_as_float(hi)));
                                v[c] = hi;
                            }
                        }
                        publish();
                        const int as = it & (kAStages - 1);
                        ptx::mbar_wait(&aEmpty[as], ((uint32_t)(it / kAStages) & 1) ^ 1);
                        ptx::tc_fence_after_sync();
                        ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * kACols, v);
                        if (X3)
                            ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * kACols + kBlockC, vlo);
                        pending = true;
                        pendStage = as;
                    }
                    if (X3 && p.ntaps > 0)
                    {
                        publish(); // the MMAs of this block cannot finish before its last A tile is published
                        ptx::mbar_wait(accBar, (uint32_t)cbRel & 1);
                        ptx::tc_fence_after_sync();
#pragma unroll
                        for (int ch = 0; ch < kOwnChunks; ++ch)
                        {
                            uint32_t pv[32];
                            ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + (g + ch * kConvGroups) * 32, pv);
                            ptx::tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                racc[ch][j] += __uint_as_float(pv[j]);
                        }
                        ptx::tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0)
                            ptx::mbar_arrive(accFree);
                    }
                }
                publish();

                // ----- epilogue -----
                const int oy = a * p.oyMul + p.oyAdd, ox = b * p.oxMul + p.oxAdd;
                const bool outOk = pixOk && oy < p.Ho && ox < p.Wo && !(dbgFlags & 2);
                const long long oplane = (long long)p.Ho * p.Wo;
                const bool raw = p.splits > 1; // partial sums: no bias, no activation
                float* op = (raw ? p.partial + split * p.partialStride : out) + (long long)n * p.K * oplane + (long long)oy * p.Wo + ox;
                const float* eb = raw ? nullptr : bias;
                const int eact = raw ? NB200_ACT_IDENTITY : p.act;
                if (!X3)
                {
                    ptx::mbar_wait(accBar, 0);
                    ptx::tc_fence_after_sync();
                }
                const bool have0 = nCb > 0 && ntapsA > 0, have1 = PAIR && nCb > 0 && p.ntaps > ntapsA;
                if constexpr (X3)
                {
#pragma unroll
                    for (int ch = 0; ch < kOwnChunks; ++ch)
                    {
                        const int c0 = (g + ch * kConvGroups) * 32;
                        if (k0 + c0 < p.K)
                        {
                            uint32_t v[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = __float_as_uint(racc[ch][j]);
                            store_chunk(op, oplane, k0 + c0, p.K, eb, lane, eact, p.alpha, outOk, v);
                        }
                    }
                }
                else
#pragma unroll 1
                for (int c0 = g * 32; c0 < BN; c0 += 32 * kConvGroups)
                {
                    if (k0 + c0 >= p.K)
                        break;
                    uint32_t v[32];
                    if (have0)
                        ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + c0, v);
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0u; // a pixel class no tap reaches (e.g. 1x1 filters, stride 2)
                    }
                    if constexpr (PAIR)
                    {
                        uint32_t v1[32];
                        if (have1)
                            ptx::tmem_ld_32x32b_x32(tmemAcc + BN + laneSel + c0, v1);
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v1[j] = 0u;
                        }
                        ptx::tmem_ld_wait();
                        store_chunk2(op, oplane, k0 + c0, p.K, eb, lane, eact, p.alpha, outOk, v, v1);
                    }
                    else
                    {
                        ptx::tmem_ld_wait();
                        store_chunk(op, oplane, k0 + c0, p.K, eb, lane, eact, p.alpha, outOk, v);
                    }
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1 && !(dbgFlags & 8))
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemAcc, kTmemCols);
            }
        }

        // ---------------------------------------------------------------- kernel-gradient kernel
        //   dw[k][c][r][s] = sum over pixels  dy[n][k][oh][ow] * x[n][c][oh+r-pY][ow+s-pX]           (stride 1)
        //   GEMM view      D_s[M = 128 channels][N = BN filters] += A_s[channel][pixel] * B[filter][pixel], reduction = pixels.
        //   CTA            one (channel tile, filter tile, filter row r) and a contiguous slice of the (n, oh) output rows
        //                  (split-K); it keeps the S accumulators of its filter row in TMEM (S x BN columns).
        //   step           32 consecutive output pixels of one row. TMA brings dy[BN k][32 pix] (K-major SWIZZLE_128B,
        //                  used as-is as the B operand of all S taps) and the matching x row segment [128 c][44 w].
        //   A operand      thread = channel: reads its 40-float row segment with 10 conflict-free LDS.128 (channel pitch
        //                  44 floats = odd multiple of 16 B), then builds the S tap-shifted 32-pixel windows from REGISTERS,
        //                  rounds to TF32 and tcgen05.st's them as 128-lane x 32-column K-major A tiles in TMEM.
        //   epilogue       coalesced partials ws[split][tap][k][c]; a second kernel adds the splits in fixed order and
        //                  transposes to KCRS (deterministic, no atomics).
        struct WgradParams
        {
            int R, S, padX, padY;
            int wOff;            // x segment starts at ow0 - wOff
            int N, Ho, Wo;       // dy extent
            int C, K;
            int segs;            // ceil(Wo / 32)
            int tilesC, tilesK, splits;
            int rowsPerSplit;    // output rows (n, oh) per split
            int stages;
            int pack;            // C <= 64: two filter taps share one A tile (lanes 0-63 tap 2g, lanes 64-127 tap 2g+1)
            int groups;          // accumulators per CTA: S, or ceil(S/2) when packed
            uint32_t xBytes;     // bytes of the x box (128 or 64 channel rows)
        };

        constexpr int kWgXW = 44;                               // x segment width in floats (pitch: 176 B)
        constexpr uint32_t kWgXBytes = 128 * kWgXW * 4;         // 22528
        constexpr int kWgAStages = 4;

        // OFF0 >= 0: three taps with compile-time window offsets OFF0 + s into the row segment (3-column filters, padX = 1:
        // OFF0 = wOff - padX = 3). The segment is rounded once and every window is a fixed register range, which removes the
        // runtime compare chain and two thirds of the rounding adds: ~175 instead of ~330 converter instructions per step
        // (the converter warps run at ~5 cycles per instruction and were what held this kernel at ~70 % of the MMA rate).
        template <int BN, bool PACK, int OFF0>
        __global__ void __launch_bounds__(kThreads, 1)
        tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDy, WgradParams p,
                        float* __restrict__ ws)
        {
            constexpr uint32_t kBBytes = BN * 32 * 4;
            constexpr uint32_t kStageBytes = ((kWgXBytes + kBBytes) + 1023) & ~1023u;
            constexpr uint32_t kTmemCols = 512;

            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            uint64_t* bars = (uint64_t*)(smem + p.stages * kStageBytes);
            uint64_t* full = bars;                 // [stages] TMA landed
            uint64_t* empty = full + 8;            // [stages] 4 converter warps + 1 MMA commit
            uint64_t* aFull = empty + 8;           // [kWgAStages]
            uint64_t* aEmpty = aFull + kWgAStages;
            uint64_t* accBar = aEmpty + kWgAStages;
            uint32_t* tmemSlot = (uint32_t*)(accBar + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;

            int t = blockIdx.x;
            const int split = t % p.splits; t /= p.splits;
            const int kt = t % p.tilesK; t /= p.tilesK;
            const int ct = t % p.tilesC; t /= p.tilesC;
            const int r = t;
            const int c0 = ct * 128, k0 = kt * BN;
            const int totalRows = p.N * p.Ho;
            const int rowBegin = split * p.rowsPerSplit;
            const int rowEnd = min(rowBegin + p.rowsPerSplit, totalRows);
            const int steps = max(rowEnd - rowBegin, 0) * p.segs;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapDy);
                for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], kTileH + 1); }
                for (int s = 0; s < kWgAStages; ++s) { ptx::mbar_init(&aFull[s], kTileH); ptx::mbar_init(&aEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, kTmemCols);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents(); // the next kernel's prologue may overlap this kernel (sm100_ptx.cuh)
            ptx::pdl_wait();              // nothing below runs before the previous kernel's memory is visible
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t tmemA = tmemAcc + kTmemCols - kWgAStages * 32;

            if (warp == 0)
            {
                if (lane == 0)
                {
                    // ===== TMA producer =====
                    int st = 0;
                    uint32_t ph = 0;
                    for (int row = rowBegin; row < rowEnd; ++row)
                    {
                        const int n = row / p.Ho, oh = row - n * p.Ho;
                        for (int seg = 0; seg < p.segs; ++seg)
                        {
                            const int ow0 = seg * 32;
                            ptx::mbar_wait(&empty[st], ph ^ 1);
                            uint8_t* xs = smem + st * kStageBytes;
                            uint8_t* bs = xs + kWgXBytes;
                            ptx::mbar_arrive_expect_tx(&full[st], p.xBytes + kBBytes);
                            // x viewed as (W, C, H, N): box {44, 128, 1, 1}; rows above/below the image read as 0
                            ptx::tma_load_4d(xs, &mapX, &full[st], ow0 - p.wOff, c0, oh + r - p.padY, n);
                            // dy viewed as (Wo, K, Ho, N): box {32, BN, 1, 1} -> [k][32 pixels], 128-byte rows, swizzled
                            ptx::tma_load_4d(bs, &mapDy, &full[st], ow0, k0, oh, n);
                            if (++st == p.stages) { st = 0; ph ^= 1; }
                        }
                    }
                }
            }
            else if (warp == 1)
            {
                // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
                constexpr uint32_t idesc = ptx::idesc_tf32(128, BN, 0, 0);
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(smem + kWgXBytes), 16, 1024);
                int st = 0, as = 0;
                uint32_t ph = 0, aph = 0;
                for (int it = 0; it < steps; ++it)
                {
                    ptx::mbar_wait(&full[st], ph);
                    const uint64_t db = descB0 + (uint64_t)((st * kStageBytes) >> 4);
                    for (int s = 0; s < p.groups; ++s)
                    {
                        ptx::mbar_wait(&aFull[as], aph);
                        ptx::tc_fence_after_sync();
                        if (ptx::elect_one())
                        {
                            const uint32_t ta = tmemA + as * 32;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                ptx::mma_tf32_ts(tmemAcc + s * BN, ta + kk * 8, db + kk * 2, idesc, (it | kk) != 0);
                            ptx::mma_commit(&aEmpty[as]);
                            if (s == p.groups - 1)
                                ptx::mma_commit(&empty[st]); // dy tile consumed
                        }
                        __syncwarp();
                        if (++as == kWgAStages) { as = 0; aph ^= 1; }
                    }
                    if (++st == p.stages) { st = 0; ph ^= 1; }
                }
                if (ptx::elect_one())
                    ptx::mma_commit(accBar);
                __syncwarp();
            }
            else
            {
                // ===== converters: thread = channel (TMEM lane); two groups of 4 warps alternate steps =====
                const int q = warp & 3;
                const int g = (warp - 2) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                const int cl = q * 32 + lane;
                const uint32_t smem32 = ptx::smem_u32(smem);
                bool pending = false;
                int pendStage = 0;
                for (int it = g; it < steps; it += 2)
                {
                    const int st = it % p.stages;
                    ptx::mbar_wait(&full[st], (uint32_t)(it / p.stages) & 1);
                    const uint32_t rowp = smem32 + st * kStageBytes + (PACK ? (cl & 63) : cl) * (kWgXW * 4);
                    uint32_t row[40];
#pragma unroll
                    for (int i = 0; i < 10; ++i)
                        ptx::lds_v4(rowp + i * 16, row[4 * i + 0], row[4 * i + 1], row[4 * i + 2], row[4 * i + 3]);
                    if constexpr (OFF0 >= 0)
                    {
#pragma unroll
                        for (int i = 0; i < 40; ++i)
                            row[i] = ptx::tf32_round_bits(row[i]);
                    }
#pragma unroll
                    for (int grp = 0; grp < (OFF0 >= 0 ? (PACK ? 2 : 3) : 8); ++grp)
                    {
                        if (OFF0 < 0 && grp >= p.groups)
                            break;
                        uint32_t v[32];
                        if constexpr (OFF0 >= 0)
                        {
                            // fixed windows: tap s = grp, or (packed) taps 2 grp / 2 grp + 1 in the two lane halves
                            if (!PACK || (cl >> 6) == 0)
                            {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    v[j] = row[j + OFF0 + (PACK ? 2 * grp : grp)];
                            }
                            else if (grp == 0)
                            {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    v[j] = row[j + OFF0 + 1];
                            }
                            else
                            {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    v[j] = 0u; // the fourth tap does not exist
                            }
                        }
                        else
                        {
                        // tap column handled by this lane in this group (packed: the two lane halves take adjacent taps)
                        const int s = PACK ? 2 * grp + (cl >> 6) : grp;
                        const int off = s - p.padX + p.wOff; // 0..8 for a real tap
                        if (PACK)
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = 0u; // lanes whose tap does not exist (odd S, upper half of the last group) feed zeros
                        }
#pragma unroll
                        for (int o = 0; o <= 8; ++o)
                            if ((!PACK || s < p.S) && off == o)
                            {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    v[j] = ptx::tf32_round_bits(row[j + o]);
                            }
                        }
                        if (pending)
                        {
                            ptx::tmem_st_wait();
                            ptx::tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0)
                                ptx::mbar_arrive(&aFull[pendStage]);
                        }
                        const int j = it * p.groups + grp; // global A-tile sequence number = MMA consumption order
                        const int as = j & (kWgAStages - 1);
                        ptx::mbar_wait(&aEmpty[as], ((uint32_t)(j / kWgAStages) & 1) ^ 1);
                        ptx::tc_fence_after_sync();
                        ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * 32, v);
                        pending = true;
                        pendStage = as;
                    }
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(&empty[st]); // x segment fully consumed into registers / TMEM stores issued
                }
                if (pending)
                {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(&aFull[pendStage]);
                }

                // ----- epilogue: partial[split][tap][k][c], lanes = consecutive channels -> coalesced -----
                ptx::mbar_wait(accBar, 0);
                ptx::tc_fence_after_sync();
                const int c = c0 + (PACK ? (cl & 63) : cl);
                const int taps = p.R * p.S;
                for (int grp = 0; grp < p.groups; ++grp)
                {
                    const int s = PACK ? 2 * grp + (cl >> 6) : grp;
                    const bool tapOk = s < p.S;
                    float* dst = ws + ((long long)(split * taps + r * p.S + (tapOk ? s : 0)) * p.K) * p.C + c;
#pragma unroll 1
                    for (int j0 = g * 32; j0 < BN; j0 += 64)
                    {
                        if (k0 + j0 >= p.K)
                            break;
                        uint32_t v[32];
                        if (steps > 0)
                        {
                            ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + grp * BN + j0, v);
                            ptx::tmem_ld_wait();
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 0u;
                        }
                        if (tapOk && c < p.C)
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (k0 + j0 + j < p.K)
                                    dst[(long long)(k0 + j0 + j) * p.C] = __uint_as_float(v[j]);
                        }
                    }
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemAcc, kTmemCols);
            }
        }



        // ---------------------------------------------------------------- kernel gradient, filter rows folded into N (C <= 64)
        // With C <= 64 and K <= 64 a converted x tile feeds only 4 MMAs of N = 64 and tc_wgrad_kernel is converter-bound
        // (220 TFLOP/s on VGG block1_conv2). A filter-row shift is a whole-row shift of dy, which TMA can express, so here
        // the R filter rows share every converted tile: a step is one INPUT row y and one 32-column segment,
        //     A_t  = the packed x windows of the step (lanes 0-63 tap s = 2t, lanes 64-127 tap s = 2t + 1), t = 0, 1
        //     B    = [dy[oh = y + padY - r][segment], r = 0 .. R-1]: R tiles of 64 filters landed back to back = one operand of
        //            64 R rows (rows outside the image: TMA zero fill)
        //     D[t][lane][r 64 + k] += A_t . B^T                                  one N = 64 R MMA per K slice
        // i.e. R times the MMA work per converted tile and a third of the MMA instructions. A CTA walks its (image, input row)
        // range row by row, all segments of a row in turn: consecutive steps then read adjacent 128-byte pieces of the same
        // DRAM pages. (A first version walked y fastest inside a segment and kept the dy tiles in a ring for R steps: it was
        // bound by DRAM page misses -- 0.44 ms with neither conversions nor MMAs issued.) Each dy tile is fetched R times, two
        // of them from L2. 3-column filters, padX = 1, R <= 3; grid = (row split, filter tile); partials and reduce as in
        // tc_wgrad_kernel.
        constexpr int kRfBN = 64;
        constexpr int kRfXBytes = 64 * kWgXW * 4;          // 11264
        constexpr int kRfDyBytes = kRfBN * 32 * 4;         // 8192

        struct RowfoldParams
        {
            int R, padY, N, H, Ho, Wo, C, K, segs, tilesK, splits, rowsPerSplit, stages;
            uint32_t stageBytes;
            int dbgFlags; // profiling only (NB200_RF_DEBUG): 1 = converters skip loads and rounding, 2 = no MMAs are issued
        };

        __global__ void __launch_bounds__(kThreads, 1)
        tc_wgrad_rowfold_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDy, RowfoldParams p,
                                float* __restrict__ ws)
        {
            constexpr int S = 3, OFF0 = 3;
            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            // stage = [R dy tiles (1 KB aligned, SW128)][x segment]
            uint64_t* bars = (uint64_t*)(smem + p.stages * p.stageBytes);
            uint64_t* full = bars;                     // [stages <= 8]
            uint64_t* empty = full + 8;                // [stages] 4 converter warps + 1 MMA commit
            uint64_t* aFull = empty + 8;               // [4]
            uint64_t* aEmpty = aFull + 4;              // [4]
            uint64_t* accBar = aEmpty + 4;
            uint32_t* tmemSlot = (uint32_t*)(accBar + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;
            const int split = blockIdx.x % p.splits;
            const int kt = blockIdx.x / p.splits;
            const int k0 = kt * kRfBN;
            const int totalRows = p.N * p.H;
            const int rowBegin = split * p.rowsPerSplit;
            const int rowEnd = min(rowBegin + p.rowsPerSplit, totalRows);
            const int steps = max(rowEnd - rowBegin, 0) * p.segs;
            const int R = p.R;
            const uint32_t xOff = (uint32_t)R * kRfDyBytes;   // x segment behind the dy tiles of the stage

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapDy);
                for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], kTileH + 1); }
                for (int s = 0; s < 4; ++s) { ptx::mbar_init(&aFull[s], kTileH); ptx::mbar_init(&aEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, 512);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents(); // the next kernel's prologue may overlap this kernel (sm100_ptx.cuh)
            ptx::pdl_wait();              // nothing below runs before the previous kernel's memory is visible
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t tmemA = tmemAcc + 512 - 4 * 32;
            const uint32_t full32 = ptx::smem_u32(full), empty32 = ptx::smem_u32(empty), aFull32 = ptx::smem_u32(aFull),
                           aEmpty32 = ptx::smem_u32(aEmpty), accBar32 = ptx::smem_u32(accBar);

            if (warp == 0)
            {
                if (lane == 0)
                {
                    // ===== TMA producer: per step the R dy tiles and the x segment of one (row, segment) =====
                    int st = 0;
                    uint32_t ph = 0;
                    for (int row = rowBegin; row < rowEnd; ++row)
                    {
                        const int n = row / p.H, y = row - n * p.H;
                        for (int seg = 0; seg < p.segs; ++seg)
                        {
                            const int ow0 = seg * 32;
                            ptx::mbar_wait(empty32 + 8u * st, ph ^ 1);
                            uint8_t* stage = smem + st * p.stageBytes;
                            ptx::mbar_arrive_expect_tx(full32 + 8u * st, xOff + kRfXBytes);
                            for (int r = 0; r < R; ++r) // dy viewed as (Wo, K, Ho, N): [64 filters][32 pixels] of output row y + padY - r
                                ptx::tma_load_4d(stage + r * kRfDyBytes, &mapDy, &full[st], ow0, k0, y + p.padY - r, n);
                            // x viewed as (W, C, H, N): box {44, 64, 1, 1}; columns left / right of the image read as zeros
                            ptx::tma_load_4d(stage + xOff, &mapX, &full[st], ow0 - 4, 0, y, n);
                            if (++st == p.stages) { st = 0; ph ^= 1; }
                        }
                    }
                }
            }
            else if (warp == 1)
            {
                // ===== MMA issuer: per step two A tiles x one B operand of 64 R rows =====
                const uint32_t idesc = ptx::idesc_tf32(128, kRfBN * R, 0, 0);
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(smem), 16, 1024);
                int st = 0;
                uint32_t ph = 0, aCount = 0;
                for (int it = 0; it < steps; ++it)
                {
                    ptx::mbar_wait(full32 + 8u * st, ph);
                    const uint64_t db = descB0 + (uint64_t)((st * p.stageBytes) >> 4);
#pragma unroll 1
                    for (int t = 0; t < 2; ++t, ++aCount)
                    {
                        const uint32_t as = aCount & 3;
                        ptx::mbar_wait(aFull32 + 8u * as, (aCount >> 2) & 1);
                        ptx::tc_fence_after_sync();
                        if (ptx::elect_one())
                        {
                            const uint32_t ta = tmemA + as * 32;
                            if (!(p.dbgFlags & 2))
                            {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk)
                                    ptx::mma_tf32_ts(tmemAcc + t * R * kRfBN, ta + kk * 8, db + kk * 2, idesc, (it | kk) != 0);
                            }
                            ptx::mma_commit(aEmpty32 + 8u * as);
                            if (t == 1)
                                ptx::mma_commit(empty32 + 8u * st); // dy tiles consumed
                        }
                        __syncwarp();
                    }
                    if (++st == p.stages) { st = 0; ph ^= 1; }
                }
                if (ptx::elect_one())
                    ptx::mma_commit(accBar32);
                __syncwarp();
            }
            else
            {
                // ===== converters: thread = (tap half, channel); two groups of 4 warps alternate steps =====
                const int q = warp & 3;
                const int g = (warp - 2) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                const int cl = q * 32 + lane;
                const uint32_t smem32 = ptx::smem_u32(smem);
                bool pending = false;
                uint32_t pendStage = 0;
                for (int it = g; it < steps; it += 2)
                {
                    const int st = it % p.stages;
                    ptx::mbar_wait(full32 + 8u * st, (uint32_t)(it / p.stages) & 1);
                    const uint32_t rowp = smem32 + st * p.stageBytes + xOff + (cl & 63) * (kWgXW * 4);
                    uint32_t row[40];
                    if (p.dbgFlags & 1)
                    {
#pragma unroll
                        for (int i = 0; i < 40; ++i) row[i] = 0x3f800000u;
                    }
                    else
                    {
#pragma unroll
                        for (int i = 0; i < 10; ++i)
                            ptx::lds_v4(rowp + i * 16, row[4 * i + 0], row[4 * i + 1], row[4 * i + 2], row[4 * i + 3]);
#pragma unroll
                        for (int i = 0; i < 40; ++i)
                            row[i] = ptx::tf32_round_bits(row[i]);
                    }
#pragma unroll
                    for (int t = 0; t < 2; ++t)
                    {
                        uint32_t v[32];
                        if ((cl >> 6) == 0)
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = row[j + OFF0 + 2 * t];
                        }
                        else if (t == 0)
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = row[j + OFF0 + 1];
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                v[j] = 0u; // the fourth tap does not exist
                        }
                        if (pending)
                        {
                            ptx::tmem_st_wait();
                            ptx::tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0)
                                ptx::mbar_arrive(aFull32 + 8u * pendStage);
                        }
                        const uint32_t j = (uint32_t)it * 2 + t; // global A-tile number = MMA consumption order
                        const uint32_t as = j & 3;               // group g only ever touches stages 2g, 2g + 1
                        ptx::mbar_wait(aEmpty32 + 8u * as, ((j >> 2) & 1) ^ 1);
                        ptx::tc_fence_after_sync();
                        ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * 32, v);
                        pending = true;
                        pendStage = as;
                    }
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(empty32 + 8u * st); // x segment is in registers / tensor memory
                }
                if (pending)
                {
                    ptx::tmem_st_wait();
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(aFull32 + 8u * pendStage);
                }

                // ----- epilogue: partial[split][tap = r*3 + s][k][c], lanes = consecutive channels -----
                ptx::mbar_wait(accBar32, 0);
                ptx::tc_fence_after_sync();
                const int c = cl & 63;
                for (int t = 0; t < 2; ++t)
                {
                    const int s = 2 * t + (cl >> 6);
                    const bool tapOk = s < S;
                    for (int r = 0; r < R; ++r)
                    {
                        float* dst = ws + ((long long)(split * R * S + r * S + (tapOk ? s : 0)) * p.K) * p.C + c;
#pragma unroll 1
                        for (int j0 = g * 32; j0 < kRfBN; j0 += 64)
                        {
                            if (k0 + j0 >= p.K)
                                break;
                            uint32_t v[32];
                            if (steps > 0)
                            {
                                ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + (t * R + r) * kRfBN + j0, v);
                                ptx::tmem_ld_wait();
                            }
                            else
                            {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = 0u;
                            }
                            if (tapOk && c < p.C)
                            {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (k0 + j0 + j < p.K)
                                        dst[(long long)(k0 + j0 + j) * p.C] = __uint_as_float(v[j]);
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
                ptx::tmem_dealloc(tmemAcc, 512);
            }
        }

        // ---------------------------------------------------------------- gathered kernel gradient (any stride / map size)
        //   dw[k][c][r][s] = sum over (n, oh, ow)  dy[n][k][oh][ow] * x[n][c][oh*st-pY+r][ow*st-pX+s]
        // Same GEMM view as tc_wgrad_kernel (M = 128 channels, N = BN filters, reduction = pixels, one accumulator per tap
        // of the CTA's tap group, split-K over pixel chunks, deterministic two-pass reduction), but built for the shapes
        // the row-segment kernel cannot take: strides, and maps so small that a 32-pixel reduction chunk spans whole
        // images. A chunk is 32 consecutive entries of the flattened (n, oh*Wo+ow) list: PXI = min(32, Ho*Wo) pixels from
        // each of 32/PXI images. dy arrives by one 3-D TMA box {PXI, BN, 32/PXI} as 32/PXI K-major sub-tiles (128-, 64- or
        // 32-byte rows, matching swizzle). x is gathered with lanes = pixels (coalesced), then each warp transposes its
        // 32 channels x 32 pixels block through a private padded shared-memory scratch so that thread = channel holds
        // the 32 pixels the TMEM A tile wants.
        struct WgatherParams
        {
            int C, H, W, K, Ho, Wo, N;
            int S, stride, padX, padY;
            int PXI;              // pixels per image per chunk
            int chunksPerImg;     // Ho*Wo / PXI
            long long chunks;     // total 32-pixel chunks
            int tilesC, tilesK, groups, tapsPerGroup, ntaps, splits;
            long long chunksPerSplit;
            int stages;
            uint32_t rowBytes, layoutType;
        };

        constexpr int kWgScratchFloats = 32 * 33;

        // X3 (NB200_MATH_3XTF32): one tap per CTA; A = [hi | lo] x windows (split by the converters, as in the forward kernels),
        // B = dy as TMA lands it (the tensor core reads its upper 19 bits = hi) plus a lo tile = tf32(dy - hi) that the
        // converter warps write behind it in the same swizzled layout; D += hi*hi + hi*lo + lo*hi. The tensor core's fp32
        // accumulator truncates on every add -- a bias that grows with the chain length, and a kernel gradient's chain is
        // thousands of MMAs long -- so the chain is cut every kWgX3Segment steps: the accumulator is drained from TMEM and
        // added, round-to-nearest, into registers (96 accumulations per segment, like the 108 of a 3x3 channel block in
        // tc_fprop_kernel<BN, true>).
        constexpr int kWgX3Segment = 8;

        // (A TF32 variant that wrote the A tiles straight into shared memory as swizzled K-major tiles -- SS-form MMA, no transposing
        // scratch, no TMEM store -- was parity-clean and SLOWER: 128 -> 128 @16x16, batch 128: 0.138 -> 0.200 ms; 256 -> 512 @31x31:
        // 0.257 -> 0.379 ms. An M128 x N128 x K8 SS MMA reads 8 KB of operands from shared memory per 64 cycles = the whole 128 B/clk
        // of the SM while the converters' stores and the TMA writes want it too; the TS form keeps A off that path.)
        template <int BN, bool X3>
        __global__ void __launch_bounds__(kThreads, 1)
        tc_wgrad_gather_kernel(const __grid_constant__ CUtensorMap mapDy, const __grid_constant__ WgatherParams p, const float* __restrict__ x,
                               float* __restrict__ ws)
        {
            constexpr uint32_t kBTile = BN * 32 * 4;
            constexpr uint32_t kBBytes = kBTile * (X3 ? 2 : 1);
            constexpr uint32_t kTmemCols = 512;
            constexpr int kACols = X3 ? 64 : 32;

            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            float* scratch = (float*)(smem + p.stages * kBBytes);              // [8 warps][32][33]
            uint64_t* bars = (uint64_t*)(scratch + 8 * kWgScratchFloats);
            uint64_t* full = bars;
            uint64_t* empty = full + 8;
            uint64_t* aFull = empty + 8;
            uint64_t* aEmpty = aFull + kWgAStages;
            uint64_t* accBar = aEmpty + kWgAStages;
            uint64_t* accFree = accBar + 1;     // X3: the accumulator has been drained into registers
            uint64_t* loFull = accFree + 1;     // X3 [stages]: the lo tile of a dy stage has been written
            uint32_t* tmemSlot = (uint32_t*)(loFull + 8);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;

            int t = blockIdx.x;
            const int split = t % p.splits; t /= p.splits;
            const int kt = t % p.tilesK; t /= p.tilesK;
            const int ct = t % p.tilesC; t /= p.tilesC;
            const int grp = t;
            const int c0 = ct * 128, k0 = kt * BN;
            const int tap0 = grp * p.tapsPerGroup;
            const int ntap = X3 ? 1 : min(p.tapsPerGroup, p.ntaps - tap0);
            const long long chunkBegin = split * p.chunksPerSplit;
            const long long chunkEnd = min(chunkBegin + p.chunksPerSplit, p.chunks);
            const int steps = (int)max(chunkEnd - chunkBegin, 0ll);
            const int imgsPerChunk = 32 / p.PXI;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapDy);
                for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); ptx::mbar_init(&loFull[s], 8); }
                for (int s = 0; s < kWgAStages; ++s) { ptx::mbar_init(&aFull[s], kTileH); ptx::mbar_init(&aEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::mbar_init(accFree, 8);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, kTmemCols);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents(); // the next kernel's prologue may overlap this kernel (sm100_ptx.cuh)
            ptx::pdl_wait();              // nothing below runs before the previous kernel's memory is visible
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t tmemA = tmemAcc + kTmemCols - kWgAStages * kACols;

            if (warp == 0)
            {
                if (lane == 0)
                {
                    int st = 0;
                    uint32_t ph = 0;
                    for (long long ch = chunkBegin; ch < chunkEnd; ++ch)
                    {
                        const long long img0 = ch / p.chunksPerImg * imgsPerChunk; // first image of the chunk
                        const int hw0 = (int)(ch % p.chunksPerImg) * p.PXI;
                        ptx::mbar_wait(&empty[st], ph ^ 1);
                        ptx::mbar_arrive_expect_tx(&full[st], kBTile);
                        // dy viewed as (Ho*Wo, K, N): box {PXI, BN, 32/PXI}; images / filters past the end read as 0
                        ptx::tma_load_3d(smem + st * kBBytes, &mapDy, &full[st], hw0, k0, (int)img0);
                        if (++st == p.stages) { st = 0; ph ^= 1; }
                    }
                }
            }
            else if (warp == 1)
            {
                constexpr uint32_t idesc = ptx::idesc_tf32(128, BN, 0, 0);
                const uint64_t descB0 = ptx::smem_desc_kmajor(ptx::smem_u32(smem), 8 * p.rowBytes, p.layoutType);
                const uint32_t subTile = BN * p.rowBytes;   // bytes of one image's sub-tile
                int st = 0, as = 0;
                uint32_t ph = 0, aph = 0;
                for (int it = 0; it < steps; ++it)
                {
                    ptx::mbar_wait(&full[st], ph);
                    if (X3)
                    {
                        ptx::mbar_wait(&loFull[st], ph);
                        if (it > 0 && it % kWgX3Segment == 0)
                            ptx::mbar_wait(accFree, (uint32_t)(it / kWgX3Segment - 1) & 1); // the previous segment is in registers
                    }
                    const uint32_t stageOff = st * kBBytes;
                    for (int tp = 0; tp < ntap; ++tp)
                    {
                        ptx::mbar_wait(&aFull[as], aph);
                        ptx::tc_fence_after_sync();
                        if (ptx::elect_one())
                        {
                            const uint32_t ta = tmemA + as * kACols;
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                            {
                                // reduction elements kk*8 .. kk*8+7: sub-tile (kk*8)/PXI, byte offset ((kk*8)%PXI)*4 inside the row
                                const uint32_t off = stageOff + ((kk * 8) / p.PXI) * subTile + (((kk * 8) % p.PXI) << 2);
                                const uint32_t accumulate = X3 ? ((it % kWgX3Segment) | kk) != 0 : (it | kk) != 0;
                                ptx::mma_tf32_ts(tmemAcc + tp * BN, ta + kk * 8, descB0 + (uint64_t)(off >> 4), idesc, accumulate);          // hi * hi
                                if (X3)
                                {
                                    ptx::mma_tf32_ts(tmemAcc, ta + kk * 8, descB0 + (uint64_t)((off + kBTile) >> 4), idesc, 1);              // hi * lo
                                    ptx::mma_tf32_ts(tmemAcc, ta + 32 + kk * 8, descB0 + (uint64_t)(off >> 4), idesc, 1);                    // lo * hi
                                }
                            }
                            ptx::mma_commit(&aEmpty[as]);
                            if (tp == ntap - 1)
                                ptx::mma_commit(&empty[st]);
                            if (X3 && ((it + 1) % kWgX3Segment == 0 || it == steps - 1))
                                ptx::mma_commit(accBar); // this segment's partial accumulator is complete
                        }
                        __syncwarp();
                        if (++as == kWgAStages) { as = 0; aph ^= 1; }
                    }
                    if (++st == p.stages) { st = 0; ph ^= 1; }
                }
                if (!X3 && ptx::elect_one())
                    ptx::mma_commit(accBar);
                __syncwarp();
            }
            else
            {
                const int q = warp & 3;
                const int g = (warp - 2) >> 2;
                const uint32_t laneSel = (uint32_t)(q * 32) << 16;
                float* myScratch = scratch + (warp - 2) * kWgScratchFloats;
                const long long plane = (long long)p.H * p.W;
                const int cw0 = c0 + q * 32;   // this warp's 32 channels
                const int hwTotal = p.Ho * p.Wo;
                bool pending = false;
                int pendStage = 0;
                const long long nTiles = (long long)steps * ntap;
                constexpr int kOwnChunks = X3 ? BN / 64 : 1;
                float racc[kOwnChunks][32];
                if (X3)
                {
#pragma unroll
                    for (int ch = 0; ch < kOwnChunks; ++ch)
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            racc[ch][j] = 0.f;
                }
                auto publish = [&]() {
                    if (pending)
                    {
                        ptx::tmem_st_wait();
                        ptx::tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0)
                            ptx::mbar_arrive(&aFull[pendStage]);
                        pending = false;
                    }
                };
                // X3: every converter warp writes its share of the lo tile of EVERY dy stage (all 8 warps, both groups)
                int loStage = 0;
                uint32_t loPhase = 0;
                auto write_lo = [&]() {
                    ptx::mbar_wait(&full[loStage], loPhase);
                    const uint32_t* hiT = (const uint32_t*)(smem + loStage * kBBytes);
                    uint32_t* loT = (uint32_t*)(smem + loStage * kBBytes + kBTile);
                    const int w8 = warp - 2;
#pragma unroll 4
                    for (int i = w8 * 32 + lane; i < (int)(kBTile / 4); i += 256)
                    {
                        const uint32_t raw = hiT[i];
                        const uint32_t hi = raw & 0xFFFFE000u;   // what the tensor core reads of the raw fp32 value
                        loT[i] = ptx::tf32_round_bits(__float_as_uint(__uint_as_float(raw) - __uint_as_float(hi)));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the MMA's async proxy
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(&loFull[loStage]);
                    if (++loStage == p.stages) { loStage = 0; loPhase ^= 1; }
                };
                auto drain = [&](int seg) {
                    publish();
                    ptx::mbar_wait(accBar, (uint32_t)seg & 1);
                    ptx::tc_fence_after_sync();
#pragma unroll
                    for (int ch = 0; ch < kOwnChunks; ++ch)
                    {
                        uint32_t pv[32];
                        ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + (g + 2 * ch) * 32, pv);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            racc[ch][j] += __uint_as_float(pv[j]);
                    }
                    ptx::tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(accFree);
                };
                for (long long j = X3 ? 0 : g; j < nTiles; j += X3 ? 1 : 2)
                {
                    const int it = (int)(j / ntap), tp = (int)(j - (long long)it * ntap);
                    if (X3)
                    {
                        write_lo();               // step `it`'s dy stage (ntap == 1: j == it); the empty barrier orders stage reuse
                        if ((it & 1) != g)
                        {
                            if ((it + 1) % kWgX3Segment == 0 || it == steps - 1)
                                drain(it / kWgX3Segment);
                            continue;
                        }
                    }
                    const int tap = tap0 + tp;
                    const int r = tap / p.S, s = tap - r * p.S;
                    // lane = pixel of the chunk
                    const long long ch = chunkBegin + it;
                    const long long img = ch / p.chunksPerImg * imgsPerChunk + lane / p.PXI;
                    const int hw = (int)(ch % p.chunksPerImg) * p.PXI + lane % p.PXI;
                    const int oh = hw / p.Wo, ow = hw - oh * p.Wo;
                    const int iy = oh * p.stride - p.padY + r, ix = ow * p.stride - p.padX + s;
                    const bool ok = img < p.N && hw < hwTotal && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                    const float* src = x + (img * p.C + cw0) * plane + (long long)iy * p.W + ix;
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                    {
                        const float f = (ok && cw0 + c < p.C) ? __ldg(src + c * plane) : 0.f;
                        myScratch[c * 33 + lane] = f;              // row = channel, column = pixel
                    }
                    __syncwarp();
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        v[i] = ptx::tf32_round_bits(__float_as_uint(myScratch[lane * 33 + i])); // lane = channel, i = pixel
                    __syncwarp();
                    uint32_t vlo[32]; // only live in the 3xTF32 instantiation
                    if (X3)
                    {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                        {
                            const float fullv = __uint_as_float(v[i] - 0x1000u);
                            const uint32_t hi = v[i] & 0xFFFFE000u;
                            vlo[i] = ptx::tf32_round_bits(__float_as_uint(fullv - __uint_as_float(hi)));
                            v[i] = hi;
                        }
                    }
                    publish();
                    const int as = (int)(j & (kWgAStages - 1));
                    ptx::mbar_wait(&aEmpty[as], ((uint32_t)(j / kWgAStages) & 1) ^ 1);
                    ptx::tc_fence_after_sync();
                    ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * kACols, v);
                    if (X3)
                        ptx::tmem_st_32x32b_x32(tmemA + laneSel + as * kACols + 32, vlo);
                    pending = true;
                    pendStage = as;
                    if (X3 && ((it + 1) % kWgX3Segment == 0 || it == steps - 1))
                        drain(it / kWgX3Segment);
                }
                publish();

                // ----- epilogue: partial[split][tap][k][c] -----
                if (!X3)
                {
                    ptx::mbar_wait(accBar, 0);
                    ptx::tc_fence_after_sync();
                }
                const int c = c0 + q * 32 + lane;
                for (int tp = 0; tp < ntap; ++tp)
                {
                    float* dst = ws + ((long long)(split * p.ntaps + tap0 + tp) * p.K) * p.C + c;
#pragma unroll 1
                    for (int j0 = g * 32; j0 < BN; j0 += 64)
                    {
                        if (k0 + j0 >= p.K)
                            break;
                        uint32_t v[32];
                        if (X3)
                        {
#pragma unroll
                            for (int ch = 0; ch < kOwnChunks; ++ch)
                                if (j0 == (g + 2 * ch) * 32)
                                {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(racc[ch][j]);
                                }
                        }
                        else if (steps > 0)
                        {
                            ptx::tmem_ld_32x32b_x32(tmemAcc + laneSel + tp * BN + j0, v);
                            ptx::tmem_ld_wait();
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 0u;
                        }
                        if (c < p.C)
                        {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (k0 + j0 + j < p.K)
                                    dst[(long long)(k0 + j0 + j) * p.C] = __uint_as_float(v[j]);
                        }
                    }
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemAcc, kTmemCols);
            }
        }

        // dw[k][c][tap] = sum over splits of partial[split][tap][k][c]
        __global__ void wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int K, int C, int taps, int splits)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            const long long total = (long long)K * C * taps;
            const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= total)
                return;
            // i enumerates (tap, k, c) with c fastest, matching the partial layout -> coalesced reads
            const int c = (int)(i % C);
            const int k = (int)((i / C) % K);
            const int tap = (int)(i / ((long long)C * K));
            float acc = 0.f;
            for (int s = 0; s < splits; ++s)
                acc += ws[(long long)s * total + i];
            dw[((long long)k * C + c) * taps + tap] = acc;
        }


        // ---------------------------------------------------------------- kernel gradient, few input channels (C <= 4, 3x3)
        // First layers (RGB input): dw has only K x 27 elements but the reduction runs over every output pixel, so the op is a
        // stream over dy (0.5 GB for VGG16 block1_conv1 at batch 8) -- HBM-bound, and far too slow on CUDA cores (measured
        // 0.72 ms against a 0.09 ms stream time). GEMM view with the roles the tensor core wants:
        //     D[M = 128 filters][N = round16(C*9)] += A[filter][pixel] * B[j = (c, r, s)][pixel],  32 pixels per step
        //     A = the dy tile exactly as TMA lands it (K-major, SWIZZLE_128B): SS-form MMA straight from shared memory
        //     B = the im2col rows of the step: J x 32 floats, written by a converter warp from the (C x 3 rows x 40 columns)
        //         x segment TMA brings with the dy tile; TF32-rounded; laid out as the same swizzled K-major tile
        // One CTA per SM walks a contiguous share of the (image, output row, 32-column segment) steps and writes one partial
        // per CTA; a reduce kernel adds the partials in fixed order (deterministic, no atomics). dy is truncated to TF32 by the
        // tensor core (its bits are never touched by a thread); x is rounded to nearest by the converters.
        //   warp 0 TMA | warp 1 MMA issue | warps 2-5 converters (step it -> warp it % 4, B stage it % 4), then epilogue
        constexpr int kScThreads = 192;
        constexpr int kScXW = 40;                                // x segment: columns ow0 - 4 .. ow0 + 35
        constexpr int kScABytes = 128 * 32 * 4;                  // dy tile
        constexpr int kScStageBytes = kScABytes + 2048;          // + x segment (C <= 4: 1920 B) rounded up; keeps every A tile 1 KB aligned
        constexpr int kScBStages = 4;
        constexpr int kScBBytes = 48 * 128;                      // im2col tile: Jpad <= 48 rows of 128 bytes (6 KB, 1 KB aligned)

        struct ScWgradParams
        {
            int C, K, J, Jpad;          // J = C * 9 live rows, Jpad = MMA N
            int N, Ho, Wo, segs, padX, padY;
            int stages, tilesK;
            long long steps;            // N * Ho * segs
            uint32_t xBytes;
        };

        __global__ void __launch_bounds__(kScThreads, 1)
        tc_smallc_wgrad_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapDy, ScWgradParams p,
                               float* __restrict__ ws)
        {
            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            uint8_t* bRing = smem + p.stages * kScStageBytes;         // kScBStages x 5 KB
            uint64_t* bars = (uint64_t*)(bRing + kScBStages * kScBBytes);
            uint64_t* full = bars;               // [stages <= 8] TMA landed
            uint64_t* empty = full + 8;          // [stages] converter done with x + MMA done with dy
            uint64_t* bFull = empty + 8;         // [4]
            uint64_t* bEmpty = bFull + 4;        // [4]
            uint64_t* accBar = bEmpty + 4;
            uint32_t* tmemSlot = (uint32_t*)(accBar + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;
            const int kt = blockIdx.x % p.tilesK;
            const int split = blockIdx.x / p.tilesK;
            const int splits = gridDim.x / p.tilesK;
            const long long per = (p.steps + splits - 1) / splits;
            const long long begin = (long long)split * per;
            const long long end = begin + per < p.steps ? begin + per : p.steps;
            const int steps = end > begin ? (int)(end - begin) : 0;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapX);
                ptx::prefetch_tensormap(&mapDy);
                for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 2); }
                for (int s = 0; s < kScBStages; ++s) { ptx::mbar_init(&bFull[s], 1); ptx::mbar_init(&bEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, 64);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents(); // the next kernel's prologue may overlap this kernel (sm100_ptx.cuh)
            ptx::pdl_wait();              // nothing below runs before the previous kernel's memory is visible
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t full32 = ptx::smem_u32(full), empty32 = ptx::smem_u32(empty), bFull32 = ptx::smem_u32(bFull),
                           bEmpty32 = ptx::smem_u32(bEmpty), accBar32 = ptx::smem_u32(accBar);

            if (warp == 0)
            {
                if (lane == 0)
                {
                    int st = 0;
                    uint32_t ph = 0;
                    long long g = begin;
                    int seg = (int)(g % p.segs);
                    long long row = g / p.segs;
                    int oh = (int)(row % p.Ho), n = (int)(row / p.Ho);
                    for (int it = 0; it < steps; ++it)
                    {
                        ptx::mbar_wait(empty32 + 8u * st, ph ^ 1);
                        uint8_t* a = smem + st * kScStageBytes;
                        ptx::mbar_arrive_expect_tx(full32 + 8u * st, kScABytes + p.xBytes);
                        // dy viewed as (Wo, K, Ho, N): [128 filters][32 pixels], rows past K read as zeros
                        ptx::tma_load_4d(a, &mapDy, &full[st], seg * 32, kt * 128, oh, n);
                        // x viewed as (W, H, C, N): [C][3 rows][40 columns]; everything outside the image reads as zeros
                        ptx::tma_load_4d(a + kScABytes, &mapX, &full[st], seg * 32 - 4, oh - p.padY, 0, n);
                        if (++st == p.stages) { st = 0; ph ^= 1; }
                        if (++seg == p.segs) { seg = 0; if (++oh == p.Ho) { oh = 0; ++n; } }
                    }
                }
            }
            else if (warp == 1)
            {
                const uint32_t idesc = ptx::idesc_tf32(128, p.Jpad, 0, 0);
                const uint64_t descA0 = ptx::smem_desc_sw128(ptx::smem_u32(smem), 16, 1024);
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(bRing), 16, 1024);
                int st = 0;
                uint32_t ph = 0;
                for (int it = 0; it < steps; ++it)
                {
                    const uint32_t bs = it & (kScBStages - 1);
                    ptx::mbar_wait(full32 + 8u * st, ph);
                    ptx::mbar_wait(bFull32 + 8u * bs, (it / kScBStages) & 1);
                    ptx::tc_fence_after_sync();
                    if (ptx::elect_one())
                    {
                        const uint64_t da = descA0 + (uint64_t)((st * kScStageBytes) >> 4);
                        const uint64_t db = descB0 + (uint64_t)((bs * kScBBytes) >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            ptx::mma_tf32_ss(tmemAcc, da + kk * 2, db + kk * 2, idesc, (it | kk) != 0);
                        ptx::mma_commit(bEmpty32 + 8u * bs);
                        ptx::mma_commit(empty32 + 8u * st);
                    }
                    __syncwarp();
                    if (++st == p.stages) { st = 0; ph ^= 1; }
                }
                if (ptx::elect_one())
                    ptx::mma_commit(accBar32);
                __syncwarp();
            }
            else
            {
                // ===== converters: lane = pixel of the step; row j = (c, r, s) of the im2col tile =====
                const int cw = warp - 2;                       // also this warp's B stage
                const uint32_t smem32 = ptx::smem_u32(smem);
                const uint32_t bTile = ptx::smem_u32(bRing) + cw * kScBBytes;
                // swizzled position of pixel `lane` in row j: 16-byte chunk (lane / 4) ^ (j % 8)
                const uint32_t laneLo = (uint32_t)(lane & 3) * 4, laneChunk = (uint32_t)(lane >> 2);
                if (steps > 0)
                {
                    // rows J .. Jpad-1 stay zero for the whole kernel (no other writer touches them)
                    for (int j = p.J; j < p.Jpad; ++j)
                        ptx::sts_b32(bTile + j * 128 + ((laneChunk ^ (uint32_t)(j & 7)) << 4) + laneLo, 0u);
                }
                for (int it = cw; it < steps; it += kScBStages)
                {
                    const int st = it % p.stages;
                    ptx::mbar_wait(full32 + 8u * st, (uint32_t)(it / p.stages) & 1);
                    ptx::mbar_wait(bEmpty32 + 8u * cw, ((uint32_t)(it / kScBStages) & 1) ^ 1);
                    const uint32_t xs = smem32 + st * kScStageBytes + kScABytes + (uint32_t)(lane + 4 - p.padX) * 4;
                    int j = 0;
                    for (int c = 0; c < p.C; ++c)
#pragma unroll
                        for (int r = 0; r < 3; ++r)
#pragma unroll
                            for (int s = 0; s < 3; ++s, ++j)
                            {
                                const uint32_t v = ptx::tf32_round_bits(ptx::lds_b32(xs + (uint32_t)((c * 3 + r) * kScXW + s) * 4));
                                ptx::sts_b32(bTile + j * 128 + ((laneChunk ^ (uint32_t)(j & 7)) << 4) + laneLo, v);
                            }
                    // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0)
                    {
                        ptx::mbar_arrive(bFull32 + 8u * cw);
                        ptx::mbar_arrive(empty32 + 8u * st);     // x segment consumed
                    }
                }

                // ----- epilogue: partial[split][kt][filter][Jpad] -----
                const int q = warp & 3;
                ptx::mbar_wait(accBar32, 0);
                ptx::tc_fence_after_sync();
                float* dst = ws + (((long long)split * p.tilesK + kt) * 128 + q * 32 + lane) * p.Jpad;
                for (int j0 = 0; j0 < p.Jpad; j0 += 8)
                {
                    uint32_t v[8];
                    if (steps > 0)
                    {
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                                     : "r"(tmemAcc + ((uint32_t)(q * 32) << 16) + j0)
                                     : "memory");
                        ptx::tmem_ld_wait();
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j0 + j] = __uint_as_float(v[j]);
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemAcc, 64);
            }
        }

        // ---------------------------------------------------------------- kernel gradient, few input channels, ANY stride / filter
        // The same GEMM with the roles swapped (D[128 filters][N = round16(C*R*S)] += dy tile * im2col rows) for the first layers
        // the kernel above cannot take: strided convolutions (pix2pix enc1, DCGAN D conv1: 3 channels, 3x3, stride 2), 4x4
        // filters with 6 channels on an input whose width is not a multiple of 4 (PatchGAN d1 on the 259 x 259 padded pair: TMA
        // cannot address x there), any padding. dy arrives by TMA exactly as above (SS-form A operand); the im2col rows are
        // gathered by the converter warps STRAIGHT from global memory / L1 (lane = output pixel, one 4-byte load per (c, r, s);
        // every x element is reused by up to R*S/stride^2 rows and stays in L1), rounded to TF32 and written as the swizzled
        // K-major B tile. C*R*S <= 96 columns of TMEM. Bound: HBM in principle (4*(C*stride^2 + K) bytes per output pixel),
        // the converters' LSU issue in practice.
        constexpr int kSgBBytes = 96 * 128;      // im2col tile: Jpad <= 96 rows of 128 bytes (12 KB, 1 KB aligned)
        constexpr int kSgTmemCols = 128;

        struct SgWgradParams
        {
            int C, K, J, Jpad, R, S, stride, padX, padY;
            int N, H, W, Ho, Wo, segs;
            int stages, tilesK;
            long long steps;            // N * Ho * segs
        };

        __global__ void __launch_bounds__(kScThreads, 1)
        tc_smallc_wgrad_gather_kernel(const __grid_constant__ CUtensorMap mapDy, SgWgradParams p, const float* __restrict__ x, float* __restrict__ ws)
        {
            extern __shared__ uint8_t smemRaw[];
            uint8_t* smem = (uint8_t*)(((uintptr_t)smemRaw + 1023) & ~(uintptr_t)1023);
            uint8_t* bRing = smem + p.stages * kScABytes;
            uint64_t* bars = (uint64_t*)(bRing + kScBStages * kSgBBytes);
            uint64_t* full = bars;               // [stages <= 8] TMA landed
            uint64_t* empty = full + 8;          // [stages] MMA done with dy
            uint64_t* bFull = empty + 8;         // [4]
            uint64_t* bEmpty = bFull + 4;        // [4]
            uint64_t* accBar = bEmpty + 4;
            uint32_t* tmemSlot = (uint32_t*)(accBar + 1);

            const int warp = threadIdx.x >> 5;
            const int lane = threadIdx.x & 31;
            const int kt = blockIdx.x % p.tilesK;
            const int split = blockIdx.x / p.tilesK;
            const int splits = gridDim.x / p.tilesK;
            const long long per = (p.steps + splits - 1) / splits;
            const long long begin = (long long)split * per;
            const long long end = begin + per < p.steps ? begin + per : p.steps;
            const int steps = end > begin ? (int)(end - begin) : 0;

            if (warp == 0 && lane == 0)
            {
                ptx::prefetch_tensormap(&mapDy);
                for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
                for (int s = 0; s < kScBStages; ++s) { ptx::mbar_init(&bFull[s], 1); ptx::mbar_init(&bEmpty[s], 1); }
                ptx::mbar_init(accBar, 1);
                ptx::fence_mbar_init();
            }
            if (warp == 1)
                ptx::tmem_alloc(tmemSlot, kSgTmemCols);
            ptx::tc_fence_before_sync();
            __syncthreads();
            ptx::tc_fence_after_sync();
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            const uint32_t tmemAcc = *tmemSlot;
            const uint32_t full32 = ptx::smem_u32(full), empty32 = ptx::smem_u32(empty), bFull32 = ptx::smem_u32(bFull),
                           bEmpty32 = ptx::smem_u32(bEmpty), accBar32 = ptx::smem_u32(accBar);

            if (warp == 0)
            {
                if (lane == 0)
                {
                    int st = 0;
                    uint32_t ph = 0;
                    long long g = begin;
                    int seg = (int)(g % p.segs);
                    long long row = g / p.segs;
                    int oh = (int)(row % p.Ho), n = (int)(row / p.Ho);
                    for (int it = 0; it < steps; ++it)
                    {
                        ptx::mbar_wait(empty32 + 8u * st, ph ^ 1);
                        ptx::mbar_arrive_expect_tx(full32 + 8u * st, kScABytes);
                        // dy viewed as (Wo, K, Ho, N): [128 filters][32 pixels], rows past K / columns past Wo read as zeros
                        ptx::tma_load_4d(smem + st * kScABytes, &mapDy, &full[st], seg * 32, kt * 128, oh, n);
                        if (++st == p.stages) { st = 0; ph ^= 1; }
                        if (++seg == p.segs) { seg = 0; if (++oh == p.Ho) { oh = 0; ++n; } }
                    }
                }
            }
            else if (warp == 1)
            {
                const uint32_t idesc = ptx::idesc_tf32(128, p.Jpad, 0, 0);
                const uint64_t descA0 = ptx::smem_desc_sw128(ptx::smem_u32(smem), 16, 1024);
                const uint64_t descB0 = ptx::smem_desc_sw128(ptx::smem_u32(bRing), 16, 1024);
                int st = 0;
                uint32_t ph = 0;
                for (int it = 0; it < steps; ++it)
                {
                    const uint32_t bs = it & (kScBStages - 1);
                    ptx::mbar_wait(full32 + 8u * st, ph);
                    ptx::mbar_wait(bFull32 + 8u * bs, (it / kScBStages) & 1);
                    ptx::tc_fence_after_sync();
                    if (ptx::elect_one())
                    {
                        const uint64_t da = descA0 + (uint64_t)((st * kScABytes) >> 4);
                        const uint64_t db = descB0 + (uint64_t)((bs * kSgBBytes) >> 4);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            ptx::mma_tf32_ss(tmemAcc, da + kk * 2, db + kk * 2, idesc, (it | kk) != 0);
                        ptx::mma_commit(bEmpty32 + 8u * bs);
                        ptx::mma_commit(empty32 + 8u * st);
                    }
                    __syncwarp();
                    if (++st == p.stages) { st = 0; ph ^= 1; }
                }
                if (ptx::elect_one())
                    ptx::mma_commit(accBar32);
                __syncwarp();
            }
            else
            {
                // ===== converters: lane = output pixel of the step; row j = (c, r, s) of the im2col tile =====
                const int cw = warp - 2;                       // also this warp's B stage
                const uint32_t bTile = ptx::smem_u32(bRing) + cw * kSgBBytes;
                // swizzled position of pixel `lane` in row j: 16-byte chunk (lane / 4) ^ (j % 8)
                const uint32_t laneLo = (uint32_t)(lane & 3) * 4, laneChunk = (uint32_t)(lane >> 2);
                if (steps > 0)
                {
                    // rows J .. Jpad-1 stay zero for the whole kernel (no other writer touches them)
                    for (int j = p.J; j < p.Jpad; ++j)
                        ptx::sts_b32(bTile + j * 128 + ((laneChunk ^ (uint32_t)(j & 7)) << 4) + laneLo, 0u);
                }
                const long long plane = (long long)p.H * p.W;
                for (int it = cw; it < steps; it += kScBStages)
                {
                    const long long g = begin + it;
                    const int seg = (int)(g % p.segs);
                    const long long row = g / p.segs;
                    const int oh = (int)(row % p.Ho), n = (int)(row / p.Ho);
                    const int ow = seg * 32 + lane;
                    const int iy0 = oh * p.stride - p.padY, ix0 = ow * p.stride - p.padX;
                    const bool pixOk = ow < p.Wo;
                    ptx::mbar_wait(bEmpty32 + 8u * cw, ((uint32_t)(it / kScBStages) & 1) ^ 1);
                    const float* xn = x + (long long)n * p.C * plane;
                    int j = 0;
                    for (int c = 0; c < p.C; ++c)
                        for (int r = 0; r < p.R; ++r)
                        {
                            const int iy = iy0 + r;
                            const bool rowOk = pixOk && iy >= 0 && iy < p.H;
                            const float* xr = xn + c * plane + (long long)iy * p.W;
                            for (int s = 0; s < p.S; ++s, ++j)
                            {
                                const int ix = ix0 + s;
                                const float f = (rowOk && ix >= 0 && ix < p.W) ? __ldg(xr + ix) : 0.f;
                                ptx::sts_b32(bTile + j * 128 + ((laneChunk ^ (uint32_t)(j & 7)) << 4) + laneLo, ptx::tf32_round_bits(__float_as_uint(f)));
                            }
                        }
                    // generic-proxy writes -> visible to the tensor core's async-proxy reads
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0)
                        ptx::mbar_arrive(bFull32 + 8u * cw);
                }

                // ----- epilogue: partial[split][kt][filter][Jpad] -----
                const int q = warp & 3;
                ptx::mbar_wait(accBar32, 0);
                ptx::tc_fence_after_sync();
                float* dst = ws + (((long long)split * p.tilesK + kt) * 128 + q * 32 + lane) * p.Jpad;
                for (int j0 = 0; j0 < p.Jpad; j0 += 8)
                {
                    uint32_t v[8];
                    if (steps > 0)
                    {
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                                     : "r"(tmemAcc + ((uint32_t)(q * 32) << 16) + j0)
                                     : "memory");
                        ptx::tmem_ld_wait();
                    }
                    else
                    {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0u;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j0 + j] = __uint_as_float(v[j]);
                }
            }

            ptx::tc_fence_before_sync();
            __syncthreads();
            if (warp == 1)
            {
                ptx::tc_fence_after_sync();
                ptx::tmem_dealloc(tmemAcc, kSgTmemCols);
            }
        }

        // dw[k][j] = sum over splits of partial[split][k / 128][k % 128][j]   (j = (c, r, s) is already the KCRS order)
        __global__ void smallc_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, int K, int J, int Jpad, int tilesK, int splits)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            const int i = blockIdx.x * blockDim.x + threadIdx.x;
            if (i >= K * J)
                return;
            const int k = i / J, j = i - k * J;
            const long long off = ((long long)(k / 128) * 128 + (k % 128)) * Jpad + j;
            const long long stride = (long long)tilesK * 128 * Jpad;
            float acc = 0.f;
            for (int s = 0; s < splits; ++s)
                acc += ws[s * stride + off];
            dw[i] = acc;
        }

        // y = activation(sum over channel splits of partial + bias): the second pass of a channel-split forward call
        __global__ void fprop_split_reduce_kernel(const float* __restrict__ partial, long long stride, int splits, const float* __restrict__ bias,
                                                  int act, float alpha, float* __restrict__ y, long long total, long long plane, int K)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
            {
                float acc = 0.f;
                for (int sp = 0; sp < splits; ++sp)
                    acc += partial[sp * stride + i];
                if (bias)
                    acc += __ldg(bias + (int)((i / plane) % K));
                y[i] = apply_activation(act, alpha, acc);
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

        int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* stridesBytes, const cuuint32_t* box,
                     CUtensorMapSwizzle swizzle)
        {
            EncodeTiledFn fn = encode_fn();
            if (!fn)
                return fail(NB200_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
            cuuint32_t es[5] = {1, 1, 1, 1, 1};
            CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, stridesBytes, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS)
                return fail(NB200_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
            return NB200_OK;
        }

        inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

        // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) property of a kernel, so the one-time
        // opt-in is keyed by the device ordinal: a process that drives several GPUs (one host thread per device) opts in on
        // each of them. The flag is only ever set after a successful call and the call is idempotent, so two host threads
        // racing on the same device at worst both make it.
        constexpr int kMaxDevices = 64;
        struct DeviceOnce { unsigned char done[kMaxDevices]; };

        template <typename Kernel>
        int opt_in_smem(DeviceOnce& once, Kernel kernel, int bytes)
        {
            int dev = 0;
            NB200_CUDA_TRY(cudaGetDevice(&dev));
            if (dev >= 0 && dev < kMaxDevices && __atomic_load_n(&once.done[dev], __ATOMIC_ACQUIRE))
                return NB200_OK;
            NB200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            if (dev >= 0 && dev < kMaxDevices)
                __atomic_store_n(&once.done[dev], (unsigned char)1, __ATOMIC_RELEASE);
            return NB200_OK;
        }

        // SM count of the current device (persistent grids are sized from it), cached per device ordinal.
        int device_sms(int* out)
        {
            static int cache[kMaxDevices];
            int dev = 0;
            NB200_CUDA_TRY(cudaGetDevice(&dev));
            int v = (dev >= 0 && dev < kMaxDevices) ? __atomic_load_n(&cache[dev], __ATOMIC_RELAXED) : 0;
            if (!v)
            {
                NB200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
                if (dev >= 0 && dev < kMaxDevices)
                    __atomic_store_n(&cache[dev], v, __ATOMIC_RELAXED);
            }
            *out = v;
            return NB200_OK;
        }

        int pick_bn(int K)
        {
            // Shared-memory traffic per MMA cycle falls with the tile's N (the converted A tile is reused across more
            // filters), so take the widest accumulator the filter count fills.
            return K <= 64 ? 64 : K <= 128 ? 128 : 256;
        }

        // Forward-shaped problem: act tensor `in` (N, Cin, Hin, Win) -> out (N, Kout, Hout, Wout), filters repacked by `mode`.
        struct FwdShape
        {
            int N, Cin, Hin, Win, Kout, Hout, Wout, R, S, padX, padY;
            int x3; // 3xTF32 operand split
        };

        struct Plan
        {
            int BN, wOff, WB, HR, xStages, bStages;
            bool pair; // CTA-pair (cta_group::2) kernel
            bool m256; // 256-pixel tile: two M halves share each filter tile (one CTA per SM)
            int splits, cbPer; // channel split for grids that leave most SMs idle (small batches, small maps)
            size_t smemBytes;
            bool ok;
        };

        Plan make_plan(const FwdShape& f)
        {
            Plan pl{};
            // The CTA-pair (cta_group::2) kernel is correct but, as measured on B200 (profiles/), slower than the
            // single-CTA kernel for every VGG/DCGAN shape, so it is opt-in (NB200_FPROP_PAIR=1) until that is understood.
            static const bool usePair = getenv("NB200_FPROP_PAIR") != nullptr;
            pl.pair = usePair && !f.x3;
            pl.BN = pick_bn(f.Kout);
            if (f.x3 && pl.BN > 128)
                pl.BN = 128; // hi+lo filter stages are twice as large; keep the ring deep enough
            {
                // A 256-wide tile runs one CTA per SM; if that leaves most SMs idle (small feature maps), halve the tile
                // width to double the number of CTAs instead.
                const long long spatial = (long long)f.N * ceil_div(f.Hout, kTileH) * ceil_div(f.Wout, kTileW);
                if (pl.BN == 256 && spatial * ceil_div(f.Kout, 256) <= 96)
                    pl.BN = 128;
            }
            {
                // 256-pixel tiles halve the filter bytes per MMA cycle. NB200_FPROP_M256=1 opts in (profiling).
                static const char* env = getenv("NB200_FPROP_M256");
                const long long tiles256 = (long long)f.N * ceil_div(f.Hout, 2 * kTileH) * ceil_div(f.Wout, kTileW) * ceil_div(f.Kout, 128);
                (void)tiles256;
                pl.m256 = false; // measured slower than the 128-pixel tile on every VGG16 layer (profiles/); opt-in only
                if (env) pl.m256 = !f.x3 && !pl.pair && env[0] == '1';
                if (pl.m256 && pl.BN > 128) pl.BN = 128;
            }
            pl.wOff = round_up(f.padX, 4);
            const int right = f.S - 1 - f.padX > 0 ? f.S - 1 - f.padX : 0;
            pl.WB = round_up(kTileW + pl.wOff + right, 4);
            pl.HR = (pl.m256 ? 2 * kTileH : kTileH) + f.R - 1;
            const size_t xBytes = ((size_t)kBlockC * pl.HR * pl.WB * 4 + 1023) & ~(size_t)1023;
            const size_t bBytes = (size_t)(pl.pair ? pl.BN / 2 : pl.BN) * kBlockC * 4 * (f.x3 ? 2 : 1);
            const size_t fixed = 1024 /*alignment slack*/ + 512 /*barriers*/;
            // Grids of at most one CTA per SM (small batches, small maps) are latency-bound on the filter ring -- nothing else
            // on the SM hides a TMA round trip -- so they take the whole shared memory for a deeper ring.
            const long long ctasTotal = (long long)f.N * ceil_div(f.Hout, kTileH) * ceil_div(f.Wout, kTileW) * ceil_div(f.Kout, pl.BN);
            {
                // Too few tiles for the chip (batch-1 style transfer on the deep 32x32 / 64x64 layers): split the channel blocks
                // of every tile over several CTAs; each writes a raw partial and a second kernel adds them in fixed order.
                static const char* env = getenv("NB200_FPROP_SPLIT"); // 0 disables (profiling)
                const int Cblocks = round_up(f.Cin, kBlockC) / kBlockC;
                pl.splits = 1;
                if (!f.x3 && !pl.pair && !pl.m256 && ctasTotal <= 74 && Cblocks >= 4 && !(env && env[0] == '0'))
                {
                    int want = (int)(148 / ctasTotal);
                    if (want > Cblocks / 2) want = Cblocks / 2;
                    if (want > 8) want = 8;
                    if (want > 1) pl.splits = want;
                }
                pl.cbPer = ceil_div(Cblocks, pl.splits);
                pl.splits = ceil_div(Cblocks, pl.cbPer);
            }
            const long long budget = (pl.BN > 128 || f.x3 || pl.m256 || ctasTotal <= 148) ? kSmemBudget1 : kSmemBudget2;
            pl.ok = false;
            // prefer two (three when alone on the SM) halo stages; give the rest to the filter ring (at least 2, at most 8)
            for (int xs = 2; xs >= 1 && !pl.ok; --xs)
            {
                const long long rest = budget - (long long)fixed - (long long)xs * (long long)xBytes;
                int bs = (int)(rest / (long long)bBytes);
                if (bs > 8) bs = 8;
                if (bs >= 2)
                {
                    pl.xStages = xs; pl.bStages = bs; pl.ok = true;
                    pl.smemBytes = fixed + xs * xBytes + bs * bBytes;
                }
            }
            return pl;
        }

        bool shape_ok(const FwdShape& f)
        {
            if (!(f.Win % 4 == 0 && f.Win >= 4 && f.Hin >= 1 && f.Cin >= 8 && f.Kout >= 8 && f.padX >= 0 && f.padY >= 0 && f.N >= 1))
                return false;
            if (f.R * f.S > 64 || f.padX > 16 || f.S - 1 - f.padX > 16)
                return false;
            const Plan pl = make_plan(f);
            return pl.ok && pl.WB <= 256 && pl.HR <= 256;
        }

        size_t repack_bytes(const FwdShape& f)
        {
            // 256-byte aligned: what follows in the workspace (channel-split partials) wants the alignment, and the launcher
            // and nb200_conv2d_workspace_bytes must agree on ONE number (odd R*S*Kout*Cp/32 is only a multiple of 128)
            return ((size_t)f.R * f.S * f.Kout * round_up(f.Cin, kBlockC) * sizeof(float) * (f.x3 ? 2 : 1) + 255) & ~(size_t)255;
        }

        template <int BN, bool X3>
        int launch_fprop(const FwdShape& f, const Plan& pl, const CUtensorMap& mapX, const CUtensorMap& mapW, const FpropParams& p,
                         const float* bias, float* out, cudaStream_t st)
        {
            static DeviceOnce attrSet{};
            if (const int rcAttr = opt_in_smem(attrSet, tc_fprop_kernel<BN, X3>, kSmemBudget1)) return rcAttr;
            const long long tiles = (long long)p.tilesK * p.tilesW * p.tilesH * f.N * p.splits;
            if (tiles > 0x3FFFFFFFll)
                return fail(NB200_E_UNSUPPORTED, "too many tiles");
            static const bool debugWaits = getenv("NB200_DEBUG_WAITS") != nullptr;
            const long long ctas = pl.pair ? 2 * tiles : tiles;
            FpropParams pd = p;
            long long* dbgDev = nullptr;
            if (debugWaits)
            {
                NB200_CUDA_TRY(cudaMalloc(&dbgDev, ctas * 4 * kDbgSlots * sizeof(long long)));
                NB200_CUDA_TRY(cudaMemset(dbgDev, 0, ctas * 4 * kDbgSlots * sizeof(long long)));
                pd.dbg = dbgDev;
            }
            if (pl.m256)
            {
                if constexpr (BN <= 128 && !X3)
                {
                    static DeviceOnce attrSetM{};
                    if (const int rcAttr = opt_in_smem(attrSetM, tc_fprop_m256_kernel<BN>, kSmemBudget1)) return rcAttr;
                    NB200_CUDA_TRY(launch_kernel(tc_fprop_m256_kernel<BN>, dim3((unsigned)tiles), dim3(kFpropThreads), pl.smemBytes, st, mapX, mapW, pd, bias, out));
                }
            }
            else if (pl.pair && !X3)
            {
                static DeviceOnce attrSet2{};
                if (const int rcAttr = opt_in_smem(attrSet2, tc_fprop2_kernel<BN>, BN > 128 ? kSmemBudget1 : kSmemBudget2)) return rcAttr;
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3((unsigned)(2 * tiles));
                cfg.blockDim = dim3(kFpropThreads);
                cfg.dynamicSmemBytes = pl.smemBytes;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                NB200_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_fprop2_kernel<BN>, mapX, mapW, pd, bias, out));
            }
            else
                NB200_CUDA_TRY(launch_kernel(tc_fprop_kernel<BN, X3>, dim3((unsigned)tiles), dim3(kFpropThreads), pl.smemBytes, st, mapX, mapW, pd, bias, out));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            if (debugWaits)
            {
                NB200_CUDA_TRY(cudaStreamSynchronize(st));
                std::vector<long long> h((size_t)ctas * 4 * kDbgSlots);
                NB200_CUDA_TRY(cudaMemcpy(h.data(), dbgDev, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                cudaFree(dbgDev);
                double sum[4][kDbgSlots] = {};
                for (long long c = 0; c < ctas; ++c)
                    for (int r = 0; r < 4; ++r)
                        for (int k = 0; k < kDbgSlots; ++k)
                            sum[r][k] += (double)h[((size_t)c * 4 + r) * kDbgSlots + k];
                const double inv = 1.0 / (double)ctas;
                fprintf(stderr, "[nb200 waits] BN=%d pair=%d ctas=%lld iters=%d | total %.0f cyc/CTA | filterTMA: bEmpty %.0f | MMA: loop %.0f issue %.0f bFull %.0f aFull %.0f | haloTMA: xEmpty %.0f | conv(w3): total %.0f xFull %.0f aEmpty %.0f load %.0f stWait %.0f arrive %.0f stIssue %.0f acc %.0f\n",
                        BN, (int)pl.pair, ctas, p.Cblocks * p.R * p.S, sum[0][kDbgTotal] * inv, sum[0][kDbgBEmpty] * inv, sum[1][kDbgTotal] * inv, sum[1][kDbgAcc] * inv, sum[1][kDbgBFull] * inv,
                        sum[1][kDbgAFull] * inv, sum[2][kDbgXEmpty] * inv, sum[3][kDbgTotal] * inv, sum[3][kDbgXFull] * inv, sum[3][kDbgAEmpty] * inv, sum[3][kDbgAFull] * inv, sum[3][kDbgBFull] * inv, sum[3][kDbgXEmpty] * inv, sum[3][kDbgBEmpty] * inv, sum[3][kDbgAcc] * inv);
            }
            return NB200_OK;
        }


        // ---- horizontal-taps-in-N kernel (tc_rowtap_kernel): 3-column filters with few output filters ----
        bool rowtap_wanted(const FwdShape& f)
        {
            static const char* env = getenv("NB200_ROWTAP"); // 0 / 1 override for profiling
            if (f.x3 || f.S != kRtS || f.padX != 1 || f.R > 7)
                return false;
            if (env)
                return env[0] == '1';
            return f.Kout <= 64 && f.Wout >= 64 && f.Wout % 8 == 0 && (long long)f.N * ceil_div(f.Hout, kTileH) * ceil_div(f.Kout, kRtBNK) >= 24; // strips for the persistent grid
        }

        size_t rowtap_bytes(const FwdShape& f)
        {
            return (size_t)f.R * ceil_div(f.Kout, kRtBNK) * kRtN * round_up(f.Cin, kBlockC) * sizeof(float);
        }

        int run_rowtap(const FwdShape& f, int repackMode, int wK, int wC, const float* in, const float* w, const float* bias, int act,
                       float alpha, float* out, void* ws, size_t wsBytes, cudaStream_t st)
        {
            const size_t need = rowtap_bytes(f);
            if (wsBytes < need || !ws)
                return fail(NB200_E_WORKSPACE, "tcgen05 conv needs %zu workspace bytes, got %zu", need, wsBytes);
            if (((uintptr_t)in & 15) || ((uintptr_t)ws & 15))
                return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");
            const int Cp = round_up(f.Cin, kBlockC);
            const int ktiles = ceil_div(f.Kout, kRtBNK);
            const int rowsPerR = ktiles * kRtN;
            float* wr = (float*)ws;
            if (g_tcFilterMode != kFiltersReady)
            {
                const long long total = (long long)f.R * rowsPerR * Cp;
                const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
                NB200_CUDA_TRY(launch_kernel(repack_rowtap_kernel, dim3(blocks), dim3(256), 0, st, w, wr, wK, wC, f.R, f.S, ktiles, Cp, repackMode));
                NB200_CUDA_TRY(cudaGetLastError());
                count_launch();
            }
            if (g_tcFilterMode == kFiltersOnly)
                return NB200_OK;
            RowtapParams p;
            p.Cblocks = Cp / kBlockC; p.R = f.R; p.padX = f.padX; p.padY = f.padY; p.HR = kTileH + f.R - 1;
            const size_t xBytes = ((size_t)kBlockC * p.HR * kRtWB * 4 + 1023) & ~(size_t)1023;
            p.xStages = 2;
            long long rest = (long long)kSmemBudget1 - 1536 - kRtHoldBytes - kRtDummyBytes - 2 * (long long)xBytes;
            p.bStages = (int)(rest / kRtBBytes);
            if (p.bStages > 8) p.bStages = 8;
            if (p.bStages < 2)
                return fail(NB200_E_UNSUPPORTED, "no shared-memory plan for this filter size");
            const size_t smemBytes = 1536 + kRtHoldBytes + kRtDummyBytes + 2 * xBytes + (size_t)p.bStages * kRtBBytes;
            p.Ho = f.Hout; p.Wo = f.Wout; p.K = f.Kout;
            p.tilesW = ceil_div(f.Wout, kTileW); p.tilesH = ceil_div(f.Hout, kTileH); p.tilesK = ktiles;
            const long long tiles = (long long)f.N * p.tilesH * p.tilesK; // strips
            if (tiles > 0x7fffffffLL)
                return fail(NB200_E_UNSUPPORTED, "too many tiles");
            p.numStrips = (int)tiles;
            p.yStrideK = (long long)f.Hout * f.Wout;
            p.yStrideN = p.yStrideK * f.Kout;
            p.act = act; p.alpha = alpha;
            {
                static const char* dbgEnv = getenv("NB200_RT_DEBUG");
                p.dbgFlags = dbgEnv ? atoi(dbgEnv) : 0;
            }

            CUtensorMap mapX, mapW;
            {
                cuuint64_t dims[4] = {(cuuint64_t)f.Win, (cuuint64_t)f.Hin, (cuuint64_t)f.Cin, (cuuint64_t)f.N};
                cuuint64_t strides[3] = {(cuuint64_t)f.Win * 4, (cuuint64_t)f.Hin * f.Win * 4, (cuuint64_t)f.Cin * f.Hin * f.Win * 4};
                cuuint32_t box[4] = {(cuuint32_t)kRtWB, (cuuint32_t)p.HR, kBlockC, 1};
                int rc = make_map(&mapX, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
                if (rc) return rc;
            }
            {
                cuuint64_t dims[3] = {(cuuint64_t)Cp, (cuuint64_t)rowsPerR, (cuuint64_t)f.R};
                cuuint64_t strides[2] = {(cuuint64_t)Cp * 4, (cuuint64_t)Cp * rowsPerR * 4};
                cuuint32_t box[3] = {kBlockC, (cuuint32_t)kRtN, 1};
                int rc = make_map(&mapW, wr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
                if (rc) return rc;
            }
            static DeviceOnce attrSet{};
            if (const int rcAttr = opt_in_smem(attrSet, tc_rowtap_kernel, kSmemBudget1)) return rcAttr;
            int smCount = 0;
            if (const int rcSm = device_sms(&smCount)) return rcSm;
            const unsigned grid = (unsigned)(tiles < smCount ? tiles : smCount);
            NB200_CUDA_TRY(launch_kernel(tc_rowtap_kernel, dim3(grid), dim3(kRtThreads), smemBytes, st, mapX, mapW, p, bias, out));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }

        // out[tap][outRows][outCp] from w[wK][wC][R][S]; mode 0 forward, 1 flipped + transposed (input gradient), 2 transposed
        int launch_repack(const float* w, float* out, int wK, int wC, int R, int S, int outRows, int outCp, int mode, int x3, cudaStream_t st)
        {
            if (g_tcFilterMode == kFiltersReady)
                return NB200_OK; // the caller's workspace already holds this layer's repacked filters (nb200_conv2d_prepare_filters)
            const int taps = R * S;
            if (taps <= 32 && outCp % 32 == 0)
            {
                if (mode == 0)
                {
                    dim3 grid((unsigned)(outCp / 32), (unsigned)wK);
                    NB200_CUDA_TRY(launch_kernel(repack_fwd_tiled_kernel, dim3(grid), dim3(288), 32 * taps * sizeof(float), st, w, out, wK, wC, taps, outCp, x3));
                }
                else
                {
                    dim3 grid((unsigned)ceil_div(outRows, 8), (unsigned)(outCp / 32));
                    NB200_CUDA_TRY(launch_kernel(repack_dgrad_tiled_kernel, dim3(grid), dim3(256), 32 * (8 * taps + 1) * sizeof(float), st, w, out, wK, wC, R, S, outRows, outCp, mode == 1, x3));
                }
            }
            else
            {
                const long long total = (long long)taps * outRows * outCp;
                const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
                NB200_CUDA_TRY(launch_kernel(repack_filters_kernel, dim3(blocks), dim3(256), 0, st, w, out, wK, wC, R, S, outRows, outCp, mode, x3));
            }
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }

        int launch_wgrad_reduce(const float* ws, float* dw, int K, int C, int taps, int splits, cudaStream_t st)
        {
            if (taps <= 32)
            {
                dim3 grid((unsigned)ceil_div(C, 32), (unsigned)K);
                NB200_CUDA_TRY(launch_kernel(wgrad_reduce_tiled_kernel, dim3(grid), dim3(288), 32 * taps * sizeof(float), st, ws, dw, K, C, taps, splits));
            }
            else
            {
                const long long total = (long long)K * C * taps;
                NB200_CUDA_TRY(launch_kernel(wgrad_reduce_kernel, dim3(ceil_div(total, 256)), dim3(256), 0, st, ws, dw, K, C, taps, splits));
            }
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }

        int run_fwd_shaped(const FwdShape& f, int repackMode, int wK, int wC, const float* in, const float* w, const float* bias, int act,
                           float alpha, float* out, void* ws, size_t wsBytes, cudaStream_t st)
        {
            if (rowtap_wanted(f))
                return run_rowtap(f, repackMode, wK, wC, in, w, bias, act, alpha, out, ws, wsBytes, st);
            const Plan pl = make_plan(f);
            if (!pl.ok)
                return fail(NB200_E_UNSUPPORTED, "no shared-memory plan for this filter size");
            const size_t repackBytes = (repack_bytes(f) + 255) & ~(size_t)255;
            const long long outElems = (long long)f.N * f.Kout * f.Hout * f.Wout;
            const size_t need = repackBytes + (pl.splits > 1 ? (size_t)pl.splits * outElems * sizeof(float) : 0);
            if (wsBytes < need || !ws)
                return fail(NB200_E_WORKSPACE, "tcgen05 conv needs %zu workspace bytes, got %zu", need, wsBytes);
            if (((uintptr_t)in & 15) || ((uintptr_t)ws & 15))
                return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");

            const int Cp = round_up(f.Cin, kBlockC);
            float* wr = (float*)ws;
            {
                const int rc = launch_repack(w, wr, wK, wC, f.R, f.S, f.Kout, Cp, repackMode, f.x3, st);
                if (rc) return rc;
            }
            if (g_tcFilterMode == kFiltersOnly)
                return NB200_OK;

            CUtensorMap mapX, mapW;
            {
                // activation in its natural (W, H, C, N) order; the box is the halo tile of one 32-channel block
                cuuint64_t dims[4] = {(cuuint64_t)f.Win, (cuuint64_t)f.Hin, (cuuint64_t)f.Cin, (cuuint64_t)f.N};
                cuuint64_t strides[3] = {(cuuint64_t)f.Win * 4, (cuuint64_t)f.Hin * f.Win * 4, (cuuint64_t)f.Cin * f.Hin * f.Win * 4};
                cuuint32_t box[4] = {(cuuint32_t)pl.WB, (cuuint32_t)pl.HR, kBlockC, 1};
                int rc = make_map(&mapX, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
                if (rc) return rc;
            }
            {
                cuuint64_t dims[3] = {(cuuint64_t)Cp, (cuuint64_t)f.Kout, (cuuint64_t)(f.R * f.S * (f.x3 ? 2 : 1))};
                cuuint64_t strides[2] = {(cuuint64_t)Cp * 4, (cuuint64_t)Cp * f.Kout * 4};
                cuuint32_t box[3] = {kBlockC, (cuuint32_t)(pl.pair ? pl.BN / 2 : pl.BN), 1};
                int rc = make_map(&mapW, wr, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
                if (rc) return rc;
            }

            FpropParams p;
            p.Cblocks = Cp / kBlockC;
            p.R = f.R; p.S = f.S; p.padX = f.padX; p.padY = f.padY;
            p.wOff = pl.wOff; p.WB = pl.WB; p.HR = pl.HR; p.xStages = pl.xStages; p.bStages = pl.bStages;
            p.Ho = f.Hout; p.Wo = f.Wout; p.K = f.Kout;
            p.tilesW = ceil_div(f.Wout, kTileW); p.tilesH = ceil_div(f.Hout, (pl.pair || pl.m256) ? 2 * kTileH : kTileH); p.tilesK = ceil_div(f.Kout, pl.BN);
            p.act = act; p.alpha = alpha; p.dbg = nullptr;
            p.yStrideK = (long long)f.Hout * f.Wout;
            p.yStrideN = p.yStrideK * f.Kout;
            p.splits = pl.splits; p.cbPer = pl.cbPer;
            p.partial = (float*)((uint8_t*)ws + repackBytes); p.partialStride = outElems;
            if (f.x3)
                return pl.BN == 64 ? launch_fprop<64, true>(f, pl, mapX, mapW, p, bias, out, st) : launch_fprop<128, true>(f, pl, mapX, mapW, p, bias, out, st);
            const int rc = pl.BN == 64 ? launch_fprop<64, false>(f, pl, mapX, mapW, p, bias, out, st)
                         : pl.BN == 128 ? launch_fprop<128, false>(f, pl, mapX, mapW, p, bias, out, st)
                                        : launch_fprop<256, false>(f, pl, mapX, mapW, p, bias, out, st);
            if (rc || pl.splits == 1)
                return rc;
            const int blocks = (int)((outElems + 255) / 256 > 148 * 8 ? 148 * 8 : (outElems + 255) / 256);
            NB200_CUDA_TRY(launch_kernel(fprop_split_reduce_kernel, dim3(blocks), dim3(256), 0, st, p.partial, outElems, pl.splits, bias, act, alpha, out, outElems, p.yStrideK, f.Kout));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }


        // ---------------------------------------------------------------- gather kernel, host side
        // Shared memory of tc_gather_kernel<BN, PAIR>: filter ring + barriers (+ 1 KB to align the ring).
        struct GatherSmem { int bStages, slack; size_t bytes; int maxBytes; };
        GatherSmem gather_smem(int BN, bool pair, bool x3)
        {
            const bool onePerSm = BN > 128 || pair || x3;
            const long long cap = onePerSm ? kSmemBudget1 : kSmemBudget2;
            const long long stage = (long long)BN * kBlockC * 4 * (x3 ? 2 : 1);
            GatherSmem g{};
            g.slack = 1024;
            g.maxBytes = (int)cap;
            long long bs = (cap - g.slack - 512) / stage;
            g.bStages = (int)(bs > 8 ? 8 : bs);
            g.bytes = (size_t)(g.slack + 512 + g.bStages * stage);
            return g;
        }

        template <int BN, bool PAIR, bool X3>
        int launch_gather(const CUtensorMap& mapW, const GatherBatch& b, const float* in, const float* bias, float* out, cudaStream_t st)
        {
            const GatherSmem sm = gather_smem(BN, PAIR, X3);
            static DeviceOnce attrSet{};
            if (const int rcAttr = opt_in_smem(attrSet, tc_gather_kernel<BN, PAIR, X3>, sm.maxBytes)) return rcAttr;
            const long long tiles = b.tileStart[b.count];
            NB200_CUDA_TRY(launch_kernel(tc_gather_kernel<BN, PAIR, X3>, dim3((unsigned)tiles), dim3(kFpropThreads), sm.bytes, st, mapW, b, in, bias, out));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }

        // Tile width and channel split of a gathered forward / input gradient. Few pixel tiles (small maps, small batches) with
        // many channels are bound by streaming the filters: with BN = 256 two CTAs would each read half of a 19 MB filter
        // tensor. So the filter tile narrows until the grid covers the chip, and what is still missing comes from splitting the
        // channel blocks of every tile over several CTAs (raw partials behind the repacked filters, fixed-order reduce).
        struct GatherPlan
        {
            int BN, splits, cbPer, Cblocks;
            bool x3;   // NB200_MATH_3XTF32: hi + lo operand tiles, one CTA per SM, BN <= 128
            bool pair; // stride-2 input gradient: the two column-parity classes of a row class share a CTA (whole-sector stores)
            long long outElems;
            size_t repackBytes, wsBytes;
        };

        GatherPlan gather_plan(int op, const nb200_conv_desc& d)
        {
            GatherPlan pl{};
            const bool fwd = op == NB200_OP_FORWARD;
            const int Kout = fwd ? d.K : d.C, Cin = fwd ? d.C : d.K;
            long long tilesM = 0;
            {
                static const char* pairEnv = getenv("NB200_GATHER_PAIR"); // 0 disables (profiling)
                pl.x3 = d.math == NB200_MATH_3XTF32;
                pl.pair = !fwd && !pl.x3 && d.stride == 2 && d.W % 2 == 0 && d.W >= 2 && !(pairEnv && pairEnv[0] == '0');
            }
            if (fwd)
                tilesM = ((long long)d.N * d.Ho * d.Wo + 127) / 128;
            else if (pl.pair)
                for (int ph = 0; ph < 2 && ph < d.H; ++ph)
                    tilesM += ((long long)d.N * ((d.H - ph + 1) / 2) * (d.W / 2) + 127) / 128;
            else
                for (int ph = 0; ph < d.stride && ph < d.H; ++ph)
                    for (int pw = 0; pw < d.stride && pw < d.W; ++pw)
                        tilesM += ((long long)d.N * ((d.H - ph + d.stride - 1) / d.stride) * ((d.W - pw + d.stride - 1) / d.stride) + 127) / 128;
            const int Cp = round_up(Cin, kBlockC);
            pl.Cblocks = Cp / kBlockC;
            pl.BN = pick_bn(Kout);
            if ((pl.pair || pl.x3) && pl.BN > 128)
                pl.BN = 128; // two accumulators (or [hi | lo] A tiles) + the A ring in 512 TMEM columns
            pl.splits = 1;
            static const char* env = getenv("NB200_GATHER_SPLIT"); // 0 disables (profiling)
            if (!(env && env[0] == '0'))
            {
                // (only below half a wave: 122 CTAs of BN = 256 beat 244 of BN = 128 -- 256 -> 512 @ 31x31: 0.093 vs 0.120 ms)
                while (pl.BN > 64 && tilesM * ceil_div(Kout, pl.BN) <= 74)
                    pl.BN /= 2;
                const long long ctas = tilesM * ceil_div(Kout, pl.BN);
                if (ctas <= 74 && pl.Cblocks >= 4)
                {
                    int want = (int)(148 / (ctas > 0 ? ctas : 1));
                    if (want > pl.Cblocks / 2) want = pl.Cblocks / 2;
                    if (want > 8) want = 8;
                    if (want > 1) pl.splits = want;
                }
            }
            pl.cbPer = ceil_div(pl.Cblocks, pl.splits);
            pl.splits = ceil_div(pl.Cblocks, pl.cbPer);
            pl.outElems = fwd ? (long long)d.N * d.K * d.Ho * d.Wo : (long long)d.N * d.C * d.H * d.W;
            pl.repackBytes = ((size_t)d.R * d.S * Kout * Cp * sizeof(float) * (pl.x3 ? 2 : 1) + 255) & ~(size_t)255;
            pl.wsBytes = pl.repackBytes + (pl.splits > 1 ? (size_t)pl.splits * pl.outElems * sizeof(float) : 0);
            return pl;
        }

        // Repack filters and build the filter tensor map.
        int gather_prepare(const GatherPlan& pl, int Kout, int Cin, int R, int S, int repackMode, int wK, int wC, const float* w, void* ws, size_t wsBytes,
                           cudaStream_t st, CUtensorMap* mapW, int* BN, int* bStages, int* Cblocks)
        {
            const int Cp = round_up(Cin, kBlockC);
            const size_t need = pl.wsBytes;
            if (wsBytes < need || !ws)
                return fail(NB200_E_WORKSPACE, "tcgen05 gather conv needs %zu workspace bytes, got %zu", need, wsBytes);
            if ((uintptr_t)ws & 15)
                return fail(NB200_E_INVALID, "workspace must be 16-byte aligned for TMA");
            {
                const int rc = launch_repack(w, (float*)ws, wK, wC, R, S, Kout, Cp, repackMode, pl.x3 ? 1 : 0, st);
                if (rc) return rc;
            }
            *BN = pl.BN;
            cuuint64_t dims[3] = {(cuuint64_t)Cp, (cuuint64_t)Kout, (cuuint64_t)(R * S * (pl.x3 ? 2 : 1))};
            cuuint64_t strides[2] = {(cuuint64_t)Cp * 4, (cuuint64_t)Cp * Kout * 4};
            cuuint32_t box[3] = {kBlockC, (cuuint32_t)*BN, 1};
            int rc = make_map(mapW, ws, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
            *bStages = gather_smem(*BN, pl.pair, pl.x3).bStages;
            *Cblocks = Cp / kBlockC;
            return NB200_OK;
        }

        bool gather_ok(const nb200_conv_desc& d)
        {
            // C or K below 8 (first / last layers of the GAN configs) run with zero-padded operand tiles: wasteful for the
            // tensor core, irrelevant in time (these layers are bound by the gather), and far faster than CUDA-core loops.
            return d.fmt == NB200_NCHW && (d.math == NB200_MATH_TF32 || d.math == NB200_MATH_3XTF32) && d.C >= 1 && d.K >= 1 && d.R * d.S <= 32 && d.N >= 1 &&
                   d.H >= 1 && d.W >= 1 && d.Ho >= 1 && d.Wo >= 1 && d.R <= 127 && d.S <= 127 && d.padX <= 127 && d.padY <= 127;
        }

        // append one pixel list to a batch; false when the batch (or the grid) is full
        bool gather_batch_add(GatherBatch& b, const GatherParams& p)
        {
            const long long tiles = (p.totalPix + 127) / 128 * p.tilesK * p.splits;
            if (b.count == kGatherMaxClasses || (long long)b.tileStart[b.count] + tiles > 0x3FFFFFFFll || p.totalPix > 0x7FFF0000ll)
                return false; // (the kernel indexes its pixel list with 32-bit integers)
            b.cls[b.count] = p;
            b.tileStart[b.count + 1] = b.tileStart[b.count] + (int)tiles;
            ++b.count;
            return true;
        }

        int dispatch_gather(int BN, bool pair, bool x3, const CUtensorMap& mapW, const GatherBatch& b, const float* in, const float* bias, float* out, cudaStream_t st)
        {
            if (b.count == 0 || b.tileStart[b.count] == 0)
                return NB200_OK;
            if (x3)
                return BN == 64 ? launch_gather<64, false, true>(mapW, b, in, bias, out, st) : launch_gather<128, false, true>(mapW, b, in, bias, out, st);
            if (pair)
                return BN == 64 ? launch_gather<64, true, false>(mapW, b, in, bias, out, st) : launch_gather<128, true, false>(mapW, b, in, bias, out, st);
            return BN == 64 ? launch_gather<64, false, false>(mapW, b, in, bias, out, st)
                 : BN == 128 ? launch_gather<128, false, false>(mapW, b, in, bias, out, st)
                             : launch_gather<256, false, false>(mapW, b, in, bias, out, st);
        }

        int dispatch_gather(int BN, bool x3, const CUtensorMap& mapW, const GatherParams& p, const float* in, const float* bias, float* out, cudaStream_t st)
        {
            GatherBatch b{};
            if (!gather_batch_add(b, p))
                return fail(NB200_E_UNSUPPORTED, "too many tiles");
            return dispatch_gather(BN, false, x3, mapW, b, in, bias, out, st);
        }

        FwdShape fwd_shape(const nb200_conv_desc& d)
        {
            return FwdShape{d.N, d.C, d.H, d.W, d.K, d.Ho, d.Wo, d.R, d.S, d.padX, d.padY, d.math == NB200_MATH_3XTF32};
        }

        // stride-1 input gradient as a forward conv of dy: pad' = F-1-pad, flipped taps, filter/channel roles swapped
        FwdShape dgrad_shape(const nb200_conv_desc& d)
        {
            return FwdShape{d.N, d.K, d.Ho, d.Wo, d.C, d.H, d.W, d.R, d.S, d.S - 1 - d.padX, d.R - 1 - d.padY, d.math == NB200_MATH_3XTF32};
        }

        // ---------------------------------------------------------------- kernel gradient, host side
        struct WgradPlan
        {
            int BN, wOff, tilesC, tilesK, splits, rowsPerSplit, stages, pack, groups;
            size_t smemBytes, wsBytes;
        };

        bool wgrad_shape_ok(const nb200_conv_desc& d)
        {
            if (d.fmt != NB200_NCHW || d.math != NB200_MATH_TF32 || d.stride != 1)
                return false;
            if (d.W % 4 || d.Wo % 4 || d.C < 8 || d.K < 8 || d.N < 1 || d.Ho < 1 || d.Wo < 1 || d.H < 1)
                return false;
            if (d.S > 6 || d.R > 16 || d.padX > 4 || d.S - 1 - d.padX > 4)
                return false;
            // A-ring discipline of tc_wgrad_kernel (see the kernel): at most 3 A tiles per step, i.e. S <= 3, or S <= 6 when
            // two taps share a tile (C <= 64). Wider filters take the gathered kernel.
            if ((d.C <= 64 ? (d.S + 1) / 2 : d.S) > 3)
                return false;
            return true;
        }

        WgradPlan wgrad_plan(const nb200_conv_desc& d)
        {
            WgradPlan pl{};
            pl.pack = d.C <= 64;
            pl.groups = pl.pack ? (d.S + 1) / 2 : d.S;
            pl.BN = (pl.groups <= 3 && d.K > 64) ? 128 : 64;
            pl.wOff = round_up(d.padX, 4);
            pl.tilesC = ceil_div(d.C, 128);
            pl.tilesK = ceil_div(d.K, pl.BN);
            const int combos = pl.tilesC * pl.tilesK * d.R;
            const int rows = d.N * d.Ho;
            int splits = 148 / combos;
            if (splits < 1) splits = 1;
            if (splits > rows) splits = rows;
            pl.rowsPerSplit = ceil_div(rows, splits);
            pl.splits = ceil_div(rows, pl.rowsPerSplit);
            const size_t stage = ((size_t)kWgXBytes + (size_t)pl.BN * 128 + 1023) & ~(size_t)1023;
            pl.stages = (int)((200 * 1024) / stage);
            if (pl.stages > 8) pl.stages = 8;
            pl.smemBytes = 1024 + 512 + pl.stages * stage;
            pl.wsBytes = (size_t)pl.splits * d.R * d.S * d.K * d.C * sizeof(float);
            return pl;
        }

        template <int BN, bool PACK, int OFF0>
        int launch_wgrad_t(const WgradPlan& pl, const CUtensorMap& mapX, const CUtensorMap& mapDy, const WgradParams& p, float* ws, cudaStream_t st)
        {
            static DeviceOnce attrSet{};
            if (const int rcAttr = opt_in_smem(attrSet, tc_wgrad_kernel<BN, PACK, OFF0>, 220 * 1024)) return rcAttr;
            const long long ctas = (long long)p.splits * p.tilesK * p.tilesC * p.R;
            NB200_CUDA_TRY(launch_kernel(tc_wgrad_kernel<BN, PACK, OFF0>, dim3((unsigned)ctas), dim3(kThreads), pl.smemBytes, st, mapX, mapDy, p, ws));
            NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
            return NB200_OK;
        }

        template <int BN, bool PACK>
        int launch_wgrad(const WgradPlan& pl, const CUtensorMap& mapX, const CUtensorMap& mapDy, const WgradParams& p, float* ws, cudaStream_t st)
        {
            static const bool generic = getenv("NB200_WGRAD_GENERIC") != nullptr; // profiling: force the runtime-offset converters
            if (!generic && p.S == 3 && p.wOff - p.padX == 3)
                return launch_wgrad_t<BN, PACK, 3>(pl, mapX, mapDy, p, ws, st);
            return launch_wgrad_t<BN, PACK, -1>(pl, mapX, mapDy, p, ws, st);
        }
    }


    static int gather_debug_flags()
    {
        static const char* env = getenv("NB200_GATHER_DEBUG");
        return env ? atoi(env) : 0;
    }

    // ---- gather kernel entry points ----
    bool tc_gather_forward_supported(const nb200_conv_desc& d) { return gather_ok(d); }
    bool tc_gather_input_gradient_supported(const nb200_conv_desc& d) { return gather_ok(d); }

    size_t tc_gather_workspace_bytes(int op, const nb200_conv_desc& d) { return gather_plan(op, d).wsBytes; }

    static int gather_split_reduce(const GatherPlan& pl, void* ws, const float* bias, int act, float alpha, float* out, long long plane, int Kout,
                                   cudaStream_t st)
    {
        if (pl.splits == 1)
            return NB200_OK;
        const float* partial = (const float*)((const uint8_t*)ws + pl.repackBytes);
        const int blocks = (int)((pl.outElems + 255) / 256 > 148 * 8 ? 148 * 8 : (pl.outElems + 255) / 256);
        NB200_CUDA_TRY(launch_kernel(fprop_split_reduce_kernel, dim3(blocks), dim3(256), 0, st, partial, pl.outElems, pl.splits, bias, act, alpha, out, pl.outElems, plane, Kout));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int tc_gather_forward(const nb200_conv_desc& d, const float* x, const float* w, const float* bias, int act, float alpha, float* y, void* ws,
                          size_t wsBytes, cudaStream_t st)
    {
        CUtensorMap mapW;
        int BN, bStages, Cblocks;
        const GatherPlan pl = gather_plan(NB200_OP_FORWARD, d);
        int rc = gather_prepare(pl, d.K, d.C, d.R, d.S, 0, d.K, d.C, w, ws, wsBytes, st, &mapW, &BN, &bStages, &Cblocks);
        if (rc) return rc;
        if (g_tcFilterMode == kFiltersOnly) return NB200_OK;
        GatherParams p{};
        p.splits = pl.splits; p.cbPer = pl.cbPer; p.partial = (float*)((uint8_t*)ws + pl.repackBytes); p.partialStride = pl.outElems;
        p.Cblocks = Cblocks; p.ntaps = d.R * d.S;
        p.C = d.C; p.H = d.H; p.W = d.W; p.K = d.K; p.Ho = d.Ho; p.Wo = d.Wo;
        p.PH = d.Ho; p.PW = d.Wo; p.oyMul = 1; p.oyAdd = 0; p.oxMul = 1; p.oxAdd = 0; p.iyMul = d.stride; p.ixMul = d.stride;
        p.totalPix = (long long)d.N * d.Ho * d.Wo;
        p.tilesK = ceil_div(d.K, BN); p.bStages = bStages; p.act = act; p.alpha = alpha;
        p.dbgFlags = gather_debug_flags();
        p.smemSlack = gather_smem(BN, false, pl.x3).slack;
        p.ntapsA = d.R * d.S; p.tapsAll = d.R * d.S;
        for (int r = 0; r < d.R; ++r)
            for (int s2 = 0; s2 < d.S; ++s2)
            {
                const int t = r * d.S + s2;
                p.iyAdd[t] = (short)(r - d.padY); p.ixAdd[t] = (short)(s2 - d.padX); p.wtap[t] = (short)t;
            }
        rc = dispatch_gather(BN, pl.x3, mapW, p, x, bias, y, st);
        if (rc) return rc;
        return gather_split_reduce(pl, ws, bias, act, alpha, y, (long long)d.Ho * d.Wo, d.K, st);
    }

    int tc_gather_input_gradient(const nb200_conv_desc& d, const float* dy, const float* w, float* dx, void* ws, size_t wsBytes, cudaStream_t st)
    {
        // produced tensor = dx (C channels, H x W); gathered tensor = dy (K channels, Ho x Wo); filters [tap][c][k]
        CUtensorMap mapW;
        int BN, bStages, Cblocks;
        const GatherPlan pl = gather_plan(NB200_OP_INPUT_GRADIENT, d);
        int rc = gather_prepare(pl, d.C, d.K, d.R, d.S, 2, d.K, d.C, w, ws, wsBytes, st, &mapW, &BN, &bStages, &Cblocks);
        if (rc) return rc;
        if (g_tcFilterMode == kFiltersOnly) return NB200_OK;
        const int st2 = d.stride;
        // NB200_GATHER_BATCH=0: one launch per parity class (profiling)
        static const char* batchEnv = getenv("NB200_GATHER_BATCH");
        const bool batched = !(batchEnv && batchEnv[0] == '0');
        GatherBatch batch{};
        // taps that reach parity class (ph, pw): (ph + padY - r) divisible by the stride (and likewise in x); appended to p
        auto add_taps = [&](GatherParams& p, int ph, int pw) {
            int nt = p.ntaps;
            for (int r = 0; r < d.R; ++r)
            {
                const int ty = ph + d.padY - r;
                if (((ty % st2) + st2) % st2 != 0) continue;
                for (int s2 = 0; s2 < d.S; ++s2)
                {
                    const int tx = pw + d.padX - s2;
                    if (((tx % st2) + st2) % st2 != 0) continue;
                    // exact division (ty, tx may be negative)
                    p.iyAdd[nt] = (short)(ty >= 0 ? ty / st2 : -((-ty) / st2));
                    p.ixAdd[nt] = (short)(tx >= 0 ? tx / st2 : -((-tx) / st2));
                    p.wtap[nt] = (short)(r * d.S + s2);
                    ++nt;
                }
            }
            p.ntaps = nt;
        };
        for (int ph = 0; ph < st2; ++ph)
            for (int pw = 0; pw < (pl.pair ? 1 : st2); ++pw)
            {
                if (ph >= d.H || pw >= d.W)
                    continue;
                GatherParams p{};
                p.Cblocks = Cblocks;
                p.splits = pl.splits; p.cbPer = pl.cbPer; p.partial = (float*)((uint8_t*)ws + pl.repackBytes); p.partialStride = pl.outElems;
                p.C = d.K; p.H = d.Ho; p.W = d.Wo; p.K = d.C; p.Ho = d.H; p.Wo = d.W;
                p.PH = (d.H - ph + st2 - 1) / st2; p.PW = pl.pair ? d.W / 2 : (d.W - pw + st2 - 1) / st2;
                p.oyMul = st2; p.oyAdd = ph; p.oxMul = st2; p.oxAdd = pw; p.iyMul = 1; p.ixMul = 1;
                p.totalPix = (long long)d.N * p.PH * p.PW;
                p.tilesK = ceil_div(d.C, BN); p.bStages = bStages; p.act = NB200_ACT_IDENTITY; p.alpha = 0.f;
                p.dbgFlags = gather_debug_flags();
                p.smemSlack = gather_smem(BN, pl.pair, pl.x3).slack;
                p.tapsAll = d.R * d.S;
                add_taps(p, ph, pw);
                p.ntapsA = p.ntaps;
                if (pl.pair)
                    add_taps(p, ph, 1); // the class of output columns 2b + 1, into the second accumulator
                if (batched && gather_batch_add(batch, p))
                    continue;
                // batch full (stride > 3) or batching disabled: flush what is queued, then start over with this class
                rc = dispatch_gather(BN, pl.pair, pl.x3, mapW, batch, dy, nullptr, dx, st);
                if (rc) return rc;
                batch = GatherBatch{};
                if (!gather_batch_add(batch, p))
                    return fail(NB200_E_UNSUPPORTED, "too many tiles");
            }
        rc = dispatch_gather(BN, pl.pair, pl.x3, mapW, batch, dy, nullptr, dx, st);
        if (rc) return rc;
        return gather_split_reduce(pl, ws, nullptr, NB200_ACT_IDENTITY, 0.f, dx, (long long)d.H * d.W, d.C, st);
    }


    // ---- gathered kernel gradient ----
    namespace
    {
        struct WgatherPlan
        {
            int BN, PXI, tapsPerGroup, groups, tilesC, tilesK, splits, stages;
            long long chunks, chunksPerSplit;
            size_t wsBytes, smemBytes, partialBytes;
            int hwPad; // dy plane pitch the tensor map needs (multiple of 4 floats); != Ho*Wo => a pitched copy of dy in the workspace
            bool ok;
        };

        // dst[plane][0 .. hwPad) = src[plane][0 .. hw), tail zero: gives dy planes the 16-byte pitch TMA requires
        __global__ void pitch_planes_kernel(const float* __restrict__ src, float* __restrict__ dst, long long planes, int hw, int hwPad)
        {
            ptx::pdl_launch_dependents();
            ptx::pdl_wait();
            const long long total = planes * hwPad;
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
            {
                const long long pl = i / hwPad;
                const int j = (int)(i - pl * hwPad);
                dst[i] = j < hw ? __ldg(src + pl * hw + j) : 0.f;
            }
        }

        WgatherPlan wgather_plan(const nb200_conv_desc& d)
        {
            WgatherPlan pl{};
            const int hw = d.Ho * d.Wo;
            // maps of > 16 pixels: 32-pixel chunks inside one image; smaller maps: 16 or 8 pixel slots from each of 2 or 4
            // images. Slots past the end of a map are zero on the dy side (TMA fills out-of-range box elements) and masked on
            // the x side. TMA needs a 16-byte plane pitch: when Ho*Wo % 4 != 0 (31x31 PatchGAN maps, 1x1 U-Net bottleneck) dy
            // is first copied into the workspace with its planes pitched to a multiple of 4 floats.
            pl.PXI = hw > 16 ? 32 : hw > 8 ? 16 : 8;
            pl.hwPad = round_up(hw, 4);
            const bool x3 = d.math == NB200_MATH_3XTF32;
            pl.ok = d.fmt == NB200_NCHW && (d.math == NB200_MATH_TF32 || x3) && hw >= 1 &&
                    d.R * d.S <= 32 && d.C >= 1 && d.K >= 1 && d.N >= 1 && d.H >= 1 && d.W >= 1;
            {
                // tiny channel counts fill a sliver of the 128 x BN tile; below this K*C the CUDA-core kernel is used instead
                static const char* env = getenv("NB200_WGRAD_GATHER_MIN_KC");
                static const long long minKC = env ? atoll(env) : 128; // measured: 8 -> 8 @7x7 batch 256: 0.032 ms direct, 0.056 ms gathered
                if ((long long)d.K * d.C < minKC)
                    pl.ok = false;
            }
            if (!pl.ok)
                return pl;
            pl.BN = d.K > 64 ? 128 : 64;
            const int ntaps = d.R * d.S;
            pl.tapsPerGroup = 384 / pl.BN; // accumulator columns (512 - 128 for the A ring) / BN
            if (pl.tapsPerGroup > ntaps) pl.tapsPerGroup = ntaps;
            if (x3) pl.tapsPerGroup = 1; // one accumulator per CTA: its running sums live in registers between segments
            pl.groups = ceil_div(ntaps, pl.tapsPerGroup);
            pl.tilesC = ceil_div(d.C, 128);
            pl.tilesK = ceil_div(d.K, pl.BN);
            const long long imgGroups = ((long long)d.N + 32 / pl.PXI - 1) / (32 / pl.PXI);
            pl.chunks = imgGroups * ceil_div(hw, pl.PXI);
            const int combos = pl.tilesC * pl.tilesK * pl.groups;
            long long splits = 148 / combos;
            if (splits < 1) splits = 1;
            if (splits > pl.chunks) splits = pl.chunks;
            pl.chunksPerSplit = (pl.chunks + splits - 1) / splits;
            pl.splits = (int)((pl.chunks + pl.chunksPerSplit - 1) / pl.chunksPerSplit);
            pl.stages = x3 ? 5 : 6;     // 3xTF32 stages carry a hi and a lo tile
            pl.smemBytes = 1024 + 512 + (size_t)pl.stages * pl.BN * 128 * (x3 ? 2 : 1) + 8 * kWgScratchFloats * sizeof(float);
            pl.partialBytes = ((size_t)pl.splits * ntaps * d.K * d.C * sizeof(float) + 255) & ~(size_t)255;
            pl.wsBytes = pl.partialBytes + (pl.hwPad != hw ? (size_t)d.N * d.K * pl.hwPad * sizeof(float) : 0);
            return pl;
        }

        template <int BN, bool X3>
        int launch_wgather(const WgatherPlan& pl, const CUtensorMap& mapDy, const WgatherParams& p, const float* x, float* ws, cudaStream_t st)
        {
            static DeviceOnce attrSet{};
            if (const int rcAttr = opt_in_smem(attrSet, tc_wgrad_gather_kernel<BN, X3>, 220 * 1024)) return rcAttr;
            const long long ctas = (long long)p.splits * p.tilesK * p.tilesC * p.groups;
            NB200_CUDA_TRY(launch_kernel(tc_wgrad_gather_kernel<BN, X3>, dim3((unsigned)ctas), dim3(kThreads), pl.smemBytes, st, mapDy, p, x, ws));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            return NB200_OK;
        }
    }

    bool tc_gather_kernels_gradient_supported(const nb200_conv_desc& d) { return wgather_plan(d).ok; }
    size_t tc_gather_kernels_gradient_workspace(const nb200_conv_desc& d) { return wgather_plan(d).wsBytes; }

    int tc_gather_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st)
    {
        const WgatherPlan pl = wgather_plan(d);
        if (!pl.ok)
            return fail(NB200_E_UNSUPPORTED, "gathered kernel gradient does not take this shape");
        if (wsBytes < pl.wsBytes || !ws)
            return fail(NB200_E_WORKSPACE, "tcgen05 kernel gradient needs %zu workspace bytes, got %zu", pl.wsBytes, wsBytes);
        if (((uintptr_t)dy & 15) || ((uintptr_t)ws & 15))
            return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");
        const int hw = d.Ho * d.Wo;
        if (pl.hwPad != hw)
        {
            float* pitched = (float*)((uint8_t*)ws + pl.partialBytes);
            const long long total = (long long)d.N * d.K * pl.hwPad;
            const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
            NB200_CUDA_TRY(launch_kernel(pitch_planes_kernel, dim3(blocks), dim3(256), 0, st, dy, pitched, (long long)d.N * d.K, hw, pl.hwPad));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
            dy = pitched;
        }
        CUtensorMap mapDy;
        {
            cuuint64_t dims[3] = {(cuuint64_t)hw, (cuuint64_t)d.K, (cuuint64_t)d.N};
            cuuint64_t strides[2] = {(cuuint64_t)pl.hwPad * 4, (cuuint64_t)d.K * pl.hwPad * 4};
            cuuint32_t box[3] = {(cuuint32_t)pl.PXI, (cuuint32_t)pl.BN, (cuuint32_t)(32 / pl.PXI)};
            const CUtensorMapSwizzle sw = pl.PXI == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : pl.PXI == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
            int rc = make_map(&mapDy, dy, 3, dims, strides, box, sw);
            if (rc) return rc;
        }
        WgatherParams p{};
        p.C = d.C; p.H = d.H; p.W = d.W; p.K = d.K; p.Ho = d.Ho; p.Wo = d.Wo; p.N = d.N;
        p.S = d.S; p.stride = d.stride; p.padX = d.padX; p.padY = d.padY;
        p.PXI = pl.PXI; p.chunksPerImg = ceil_div(hw, pl.PXI); p.chunks = pl.chunks;
        p.tilesC = pl.tilesC; p.tilesK = pl.tilesK; p.groups = pl.groups; p.tapsPerGroup = pl.tapsPerGroup; p.ntaps = d.R * d.S;
        p.splits = pl.splits; p.chunksPerSplit = pl.chunksPerSplit; p.stages = pl.stages;
        p.rowBytes = (uint32_t)pl.PXI * 4; p.layoutType = pl.PXI == 32 ? 2u : pl.PXI == 16 ? 4u : 6u;
        const bool x3 = d.math == NB200_MATH_3XTF32;
        int rc = x3 ? (pl.BN == 64 ? launch_wgather<64, true>(pl, mapDy, p, x, (float*)ws, st) : launch_wgather<128, true>(pl, mapDy, p, x, (float*)ws, st))
                    : (pl.BN == 64 ? launch_wgather<64, false>(pl, mapDy, p, x, (float*)ws, st) : launch_wgather<128, false>(pl, mapDy, p, x, (float*)ws, st));
        if (rc) return rc;
        return launch_wgrad_reduce((const float*)ws, dw, d.K, d.C, d.R * d.S, pl.splits, st);
    }


    // ---- small-channel kernel gradient on the tensor cores (tc_smallc_wgrad_kernel) ----
    static int sc_splits(const nb200_conv_desc& d)
    {
        const int tilesK = ceil_div(d.K, 128);
        const long long steps = (long long)d.N * d.Ho * ceil_div(d.Wo, 32);
        long long splits = 148 / tilesK;
        if (splits < 1) splits = 1;
        if (splits > steps) splits = steps > 0 ? steps : 1;
        return (int)splits;
    }

    bool tc_smallc_wgrad_supported(const nb200_conv_desc& d)
    {
        return d.math == NB200_MATH_TF32 && d.fmt == NB200_NCHW && d.C >= 1 && d.C <= 4 && d.R == 3 && d.S == 3 && d.stride == 1 &&
               d.padX <= 2 && d.padY <= 2 && d.W % 4 == 0 && d.Wo % 4 == 0 && d.K >= 1 && d.Ho == d.H + 2 * d.padY - 2 &&
               d.Wo == d.W + 2 * d.padX - 2 && (long long)d.N * d.Ho * d.Wo >= 64 * 1024;
    }

    size_t tc_smallc_wgrad_workspace(const nb200_conv_desc& d)
    {
        const int Jpad = round_up(d.C * 9, 16);
        return (size_t)sc_splits(d) * ceil_div(d.K, 128) * 128 * Jpad * sizeof(float);
    }

    int tc_smallc_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st)
    {
        const size_t need = tc_smallc_wgrad_workspace(d);
        if (wsBytes < need || !ws)
            return fail(NB200_E_WORKSPACE, "kernel gradient needs %zu workspace bytes, got %zu", need, wsBytes);
        if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15))
            return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");
        ScWgradParams p;
        p.C = d.C; p.K = d.K; p.J = d.C * 9; p.Jpad = round_up(p.J, 16);
        p.N = d.N; p.Ho = d.Ho; p.Wo = d.Wo; p.segs = ceil_div(d.Wo, 32); p.padX = d.padX; p.padY = d.padY;
        p.tilesK = ceil_div(d.K, 128);
        p.steps = (long long)d.N * d.Ho * p.segs;
        p.xBytes = (uint32_t)(d.C * 3 * kScXW * 4);
        p.stages = 8;
        const size_t smemBytes = 1024 + (size_t)p.stages * kScStageBytes + kScBStages * kScBBytes + 512;
        CUtensorMap mapX, mapDy;
        {
            cuuint64_t dims[4] = {(cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.C, (cuuint64_t)d.N};
            cuuint64_t strides[3] = {(cuuint64_t)d.W * 4, (cuuint64_t)d.H * d.W * 4, (cuuint64_t)d.C * d.H * d.W * 4};
            cuuint32_t box[4] = {(cuuint32_t)kScXW, 3, (cuuint32_t)d.C, 1};
            int rc = make_map(&mapX, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
        }
        {
            cuuint64_t dims[4] = {(cuuint64_t)d.Wo, (cuuint64_t)d.K, (cuuint64_t)d.Ho, (cuuint64_t)d.N};
            cuuint64_t strides[3] = {(cuuint64_t)d.Ho * d.Wo * 4, (cuuint64_t)d.Wo * 4, (cuuint64_t)d.K * d.Ho * d.Wo * 4};
            cuuint32_t box[4] = {32, 128, 1, 1};
            int rc = make_map(&mapDy, dy, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
        static DeviceOnce attrSet{};
        if (const int rcAttr = opt_in_smem(attrSet, tc_smallc_wgrad_kernel, 220 * 1024)) return rcAttr;
        const int splits = sc_splits(d);
        NB200_CUDA_TRY(launch_kernel(tc_smallc_wgrad_kernel, dim3((unsigned)(splits * p.tilesK)), dim3(kScThreads), smemBytes, st, mapX, mapDy, p, (float*)ws));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        NB200_CUDA_TRY(launch_kernel(smallc_wgrad_reduce_kernel, dim3(ceil_div((long long)d.K * p.J, 256)), dim3(256), 0, st, (const float*)ws, dw, d.K, p.J, p.Jpad, p.tilesK, splits));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    // ---- small-channel kernel gradient, any stride / filter size (tc_smallc_wgrad_gather_kernel) ----
    bool tc_smallc_wgrad_gather_supported(const nb200_conv_desc& d)
    {
        static const char* env = getenv("NB200_SMALLC_WGRAD_GATHER"); // 0 disables (profiling)
        if (env && env[0] == '0')
            return false;
        return d.math == NB200_MATH_TF32 && d.fmt == NB200_NCHW && d.C >= 1 && d.C <= 8 && d.R >= 1 && d.S >= 1 && d.C * d.R * d.S <= 96 &&
               d.stride >= 1 && d.stride <= 4 && d.Wo % 4 == 0 && d.K >= 8 && d.N >= 1 && d.H >= 1 && d.W >= 1 && d.Ho >= 1 &&
               (long long)d.N * d.Ho * d.Wo >= 16 * 1024 && !tc_smallc_wgrad_supported(d);
    }

    static int sg_splits(const nb200_conv_desc& d)
    {
        const int tilesK = ceil_div(d.K, 128);
        const long long steps = (long long)d.N * d.Ho * ceil_div(d.Wo, 32);
        long long splits = 148 / tilesK;
        if (splits < 1) splits = 1;
        if (splits > steps) splits = steps > 0 ? steps : 1;
        return (int)splits;
    }

    size_t tc_smallc_wgrad_gather_workspace(const nb200_conv_desc& d)
    {
        const int Jpad = round_up(d.C * d.R * d.S, 16);
        return (size_t)sg_splits(d) * ceil_div(d.K, 128) * 128 * Jpad * sizeof(float);
    }

    int tc_smallc_gather_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st)
    {
        const size_t need = tc_smallc_wgrad_gather_workspace(d);
        if (wsBytes < need || !ws)
            return fail(NB200_E_WORKSPACE, "kernel gradient needs %zu workspace bytes, got %zu", need, wsBytes);
        if ((uintptr_t)dy & 15)
            return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");
        SgWgradParams p;
        p.C = d.C; p.K = d.K; p.J = d.C * d.R * d.S; p.Jpad = round_up(p.J, 16); p.R = d.R; p.S = d.S; p.stride = d.stride; p.padX = d.padX; p.padY = d.padY;
        p.N = d.N; p.H = d.H; p.W = d.W; p.Ho = d.Ho; p.Wo = d.Wo; p.segs = ceil_div(d.Wo, 32);
        p.tilesK = ceil_div(d.K, 128);
        p.steps = (long long)d.N * d.Ho * p.segs;
        p.stages = 8;
        const size_t smemBytes = 1024 + (size_t)p.stages * kScABytes + kScBStages * kSgBBytes + 512;
        CUtensorMap mapDy;
        {
            cuuint64_t dims[4] = {(cuuint64_t)d.Wo, (cuuint64_t)d.K, (cuuint64_t)d.Ho, (cuuint64_t)d.N};
            cuuint64_t strides[3] = {(cuuint64_t)d.Ho * d.Wo * 4, (cuuint64_t)d.Wo * 4, (cuuint64_t)d.K * d.Ho * d.Wo * 4};
            cuuint32_t box[4] = {32, 128, 1, 1};
            int rc = make_map(&mapDy, dy, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
        static DeviceOnce attrSet{};
        if (const int rcAttr = opt_in_smem(attrSet, tc_smallc_wgrad_gather_kernel, 220 * 1024)) return rcAttr;
        const int splits = sg_splits(d);
        NB200_CUDA_TRY(launch_kernel(tc_smallc_wgrad_gather_kernel, dim3((unsigned)(splits * p.tilesK)), dim3(kScThreads), smemBytes, st, mapDy, p, x, (float*)ws));
        count_launch();
        NB200_CUDA_TRY(launch_kernel(smallc_wgrad_reduce_kernel, dim3(ceil_div((long long)d.K * p.J, 256)), dim3(256), 0, st, (const float*)ws, dw, d.K, p.J, p.Jpad, p.tilesK, splits));
        count_launch();
        return NB200_OK;
    }

    bool tc_forward_supported(const nb200_conv_desc& d)
    {
        return d.fmt == NB200_NCHW && (d.math == NB200_MATH_TF32 || d.math == NB200_MATH_3XTF32) && d.stride == 1 && shape_ok(fwd_shape(d));
    }

    bool tc_input_gradient_supported(const nb200_conv_desc& d)
    {
        if (d.fmt != NB200_NCHW || (d.math != NB200_MATH_TF32 && d.math != NB200_MATH_3XTF32) || d.stride != 1)
            return false;
        // the gather needs dx extent == what a forward conv of dy with pad' = F-1-pad produces
        if (d.H != d.Ho + d.R - 1 - 2 * d.padY || d.W != d.Wo + d.S - 1 - 2 * d.padX)
            return false;
        return shape_ok(dgrad_shape(d));
    }

    bool tc_kernels_gradient_supported(const nb200_conv_desc& d) { return wgrad_shape_ok(d); }

    static bool rowfold_wanted(const nb200_conv_desc& d);
    static int rowfold_splits(const nb200_conv_desc& d);

    static size_t fwd_ws_bytes(const FwdShape& f)
    {
        size_t a = repack_bytes(f);
        if (shape_ok(f))
        {
            const Plan pl = make_plan(f);
            if (pl.ok && pl.splits > 1) // channel-split forward: partials behind the (256-byte aligned) repacked filters
                a = ((a + 255) & ~(size_t)255) + (size_t)pl.splits * f.N * f.Kout * f.Hout * f.Wout * sizeof(float);
        }
        const size_t b = rowtap_wanted(f) ? rowtap_bytes(f) : 0;
        return a > b ? a : b;
    }

    bool tc_uses_rowfold(const nb200_conv_desc& d) { return rowfold_wanted(d); }

    bool tc_uses_rowtap(int op, const nb200_conv_desc& d)
    {
        return op == NB200_OP_FORWARD ? rowtap_wanted(fwd_shape(d)) : op == NB200_OP_INPUT_GRADIENT ? rowtap_wanted(dgrad_shape(d)) : false;
    }

    size_t tc_workspace_bytes(int op, const nb200_conv_desc& d)
    {
        switch (op)
        {
        case NB200_OP_FORWARD: return fwd_ws_bytes(fwd_shape(d));
        case NB200_OP_INPUT_GRADIENT: return fwd_ws_bytes(dgrad_shape(d));
        default:
        {
            const size_t a = wgrad_plan(d).wsBytes;
            const size_t b = rowfold_wanted(d) ? (size_t)rowfold_splits(d) * d.R * d.S * d.K * d.C * sizeof(float) : 0;
            return a > b ? a : b;
        }
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


    // ---- kernel gradient with the filter rows folded into N (tc_wgrad_rowfold_kernel) ----
    static bool rowfold_wanted(const nb200_conv_desc& d)
    {
        static const char* env = getenv("NB200_WGRAD_ROWFOLD"); // 0 / 1 override for profiling
        if (!wgrad_shape_ok(d) || d.S != 3 || d.padX != 1 || d.R > 3 || d.C > 64)
            return false;
        if (env)
            return env[0] == '1';
        return d.K <= 128 && (long long)d.N * d.H >= 64;
    }

    static int rowfold_splits(const nb200_conv_desc& d)
    {
        const int tilesK = ceil_div(d.K, kRfBN);
        int splits = 148 / tilesK;
        if (splits < 1) splits = 1;
        const int rows = d.N * d.H;
        if (splits > rows) splits = rows;
        const int per = ceil_div(rows, splits);
        return ceil_div(rows, per);
    }

    static int tc_rowfold_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st)
    {
        RowfoldParams p;
        p.R = d.R; p.padY = d.padY; p.N = d.N; p.H = d.H; p.Ho = d.Ho; p.Wo = d.Wo; p.C = d.C; p.K = d.K;
        p.segs = ceil_div(d.Wo, 32); p.tilesK = ceil_div(d.K, kRfBN);
        p.splits = rowfold_splits(d);
        {
            static const char* dbgEnv = getenv("NB200_RF_DEBUG");
            p.dbgFlags = dbgEnv ? atoi(dbgEnv) : 0;
        }
        p.rowsPerSplit = ceil_div(d.N * d.H, p.splits);
        const size_t need = (size_t)p.splits * d.R * d.S * d.K * d.C * sizeof(float);
        if (wsBytes < need || !ws)
            return fail(NB200_E_WORKSPACE, "tcgen05 kernel gradient needs %zu workspace bytes, got %zu", need, wsBytes);
        if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15))
            return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");
        CUtensorMap mapX, mapDy;
        {
            cuuint64_t dims[4] = {(cuuint64_t)d.W, (cuuint64_t)d.C, (cuuint64_t)d.H, (cuuint64_t)d.N};
            cuuint64_t strides[3] = {(cuuint64_t)d.H * d.W * 4, (cuuint64_t)d.W * 4, (cuuint64_t)d.C * d.H * d.W * 4};
            cuuint32_t box[4] = {kWgXW, 64, 1, 1};
            int rc = make_map(&mapX, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
        }
        {
            cuuint64_t dims[4] = {(cuuint64_t)d.Wo, (cuuint64_t)d.K, (cuuint64_t)d.Ho, (cuuint64_t)d.N};
            cuuint64_t strides[3] = {(cuuint64_t)d.Ho * d.Wo * 4, (cuuint64_t)d.Wo * 4, (cuuint64_t)d.K * d.Ho * d.Wo * 4};
            cuuint32_t box[4] = {32, (cuuint32_t)kRfBN, 1, 1};
            int rc = make_map(&mapDy, dy, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
        static DeviceOnce attrSet{};
        if (const int rcAttr = opt_in_smem(attrSet, tc_wgrad_rowfold_kernel, 220 * 1024)) return rcAttr;
        p.stageBytes = (uint32_t)((d.R * kRfDyBytes + kRfXBytes + 1023) & ~1023);
        p.stages = (int)((200 * 1024) / p.stageBytes);
        if (p.stages > 8) p.stages = 8;
        const size_t smemBytes = 1024 + (size_t)p.stages * p.stageBytes + 512;
        NB200_CUDA_TRY(launch_kernel(tc_wgrad_rowfold_kernel, dim3((unsigned)(p.splits * p.tilesK)), dim3(kThreads), smemBytes, st, mapX, mapDy, p, (float*)ws));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return launch_wgrad_reduce((const float*)ws, dw, d.K, d.C, d.R * d.S, p.splits, st);
    }

    int tc_kernels_gradient(const nb200_conv_desc& d, const float* x, const float* dy, float* dw, void* ws, size_t wsBytes, cudaStream_t st)
    {
        if (rowfold_wanted(d))
            return tc_rowfold_kernels_gradient(d, x, dy, dw, ws, wsBytes, st);
        const WgradPlan pl = wgrad_plan(d);
        if (wsBytes < pl.wsBytes || !ws)
            return fail(NB200_E_WORKSPACE, "tcgen05 kernel gradient needs %zu workspace bytes, got %zu", pl.wsBytes, wsBytes);
        if (((uintptr_t)x & 15) || ((uintptr_t)dy & 15))
            return fail(NB200_E_INVALID, "tensor base addresses must be 16-byte aligned for TMA");
        CUtensorMap mapX, mapDy;
        {
            cuuint64_t dims[4] = {(cuuint64_t)d.W, (cuuint64_t)d.C, (cuuint64_t)d.H, (cuuint64_t)d.N};
            cuuint64_t strides[3] = {(cuuint64_t)d.H * d.W * 4, (cuuint64_t)d.W * 4, (cuuint64_t)d.C * d.H * d.W * 4};
            cuuint32_t box[4] = {kWgXW, (cuuint32_t)(pl.pack ? 64 : 128), 1, 1};
            int rc = make_map(&mapX, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
        }
        {
            cuuint64_t dims[4] = {(cuuint64_t)d.Wo, (cuuint64_t)d.K, (cuuint64_t)d.Ho, (cuuint64_t)d.N};
            cuuint64_t strides[3] = {(cuuint64_t)d.Ho * d.Wo * 4, (cuuint64_t)d.Wo * 4, (cuuint64_t)d.K * d.Ho * d.Wo * 4};
            cuuint32_t box[4] = {32, (cuuint32_t)pl.BN, 1, 1};
            int rc = make_map(&mapDy, dy, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
        }
        WgradParams p;
        p.R = d.R; p.S = d.S; p.padX = d.padX; p.padY = d.padY; p.wOff = pl.wOff;
        p.N = d.N; p.Ho = d.Ho; p.Wo = d.Wo; p.C = d.C; p.K = d.K;
        p.segs = ceil_div(d.Wo, 32);
        p.tilesC = pl.tilesC; p.tilesK = pl.tilesK; p.splits = pl.splits; p.rowsPerSplit = pl.rowsPerSplit; p.stages = pl.stages;
        p.pack = pl.pack; p.groups = pl.groups; p.xBytes = (uint32_t)((pl.pack ? 64 : 128) * kWgXW * 4);
        int rc = pl.pack ? (pl.BN == 64 ? launch_wgrad<64, true>(pl, mapX, mapDy, p, (float*)ws, st) : launch_wgrad<128, true>(pl, mapX, mapDy, p, (float*)ws, st))
                         : (pl.BN == 64 ? launch_wgrad<64, false>(pl, mapX, mapDy, p, (float*)ws, st) : launch_wgrad<128, false>(pl, mapX, mapDy, p, (float*)ws, st));
        if (rc) return rc;
        return launch_wgrad_reduce((const float*)ws, dw, d.K, d.C, d.R * d.S, pl.splits, st);
    }
}
