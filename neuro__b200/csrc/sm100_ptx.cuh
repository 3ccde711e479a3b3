// Thin inline-PTX wrappers for the sm_100a features the implicit-GEMM kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / fences), UMMA descriptors.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nb200
{
    namespace ptx
    {
        __device__ __forceinline__ uint32_t smem_u32(const void* p)
        {
            return (uint32_t)__cvta_generic_to_shared(p);
        }

        __device__ __forceinline__ uint32_t lane_id()
        {
            return threadIdx.x & 31;
        }

        // One lane of a converged warp (always the same one). Keeps the surrounding control flow warp-uniform so the
        // compiler holds descriptors / addresses in uniform registers instead of emitting per-lane "waterfall" loops
        // around tcgen05 and TMA instructions (measured: ~1000 cycles per 4 MMAs when issued from `if (lane == 0)`).
        __device__ __forceinline__ uint32_t elect_one()
        {
            uint32_t pred;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "elect.sync _|p, 0xffffffff;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(pred));
            return pred;
        }

        // Shared-memory loads by 32-bit shared address: keeps the compiler on LDS with immediate offsets instead of
        // generic 64-bit LD.E address arithmetic (measured: ~8 instructions per element in the converter loops).
        __device__ __forceinline__ uint32_t lds_b32(uint32_t addr)
        {
            uint32_t v;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
            return v;
        }

        __device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v)
        {
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
        }

        // Ampere-style asynchronous 4-byte gather copies (LDGSTS): global -> shared without holding a register while the load
        // is in flight. srcBytes = 0 reads nothing and writes zeros (padding taps, ragged channels); `src` must still be a
        // valid address. Completion is tracked per thread in commit groups.
        __device__ __forceinline__ void cp_async_4(uint32_t smemDst, const void* src, uint32_t srcBytes)
        {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smemDst), "l"(src), "r"(srcBytes) : "memory");
        }
        __device__ __forceinline__ void cp_async_commit()
        {
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        template <int N>
        __device__ __forceinline__ void cp_async_wait()
        {
            asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
        }

        __device__ __forceinline__ void lds_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d)
        {
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
        }

        // fp32 bits -> TF32 operand bits, round-to-nearest (ties away from zero): the tensor core ignores the low 13
        // mantissa bits, so adding half a TF32 ulp to the magnitude is all that is needed (1 integer add per element;
        // cvt.rna.tf32.f32 expands to 2-3 ALU instructions).
        __device__ __forceinline__ uint32_t tf32_round_bits(uint32_t fp32Bits)
        {
            return fp32Bits + 0x1000u;
        }

        // ---------------- mbarrier ----------------
        __device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
        }

        __device__ __forceinline__ void fence_mbar_init()
        {
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }

        __device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        }

        __device__ __forceinline__ void mbar_arrive(uint64_t* bar)
        {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
        }

        // ---------------- thread-block clusters (CTA pairs) ----------------
        __device__ __forceinline__ uint32_t cluster_ctarank()
        {
            uint32_t r;
            asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
            return r;
        }

        __device__ __forceinline__ void cluster_sync()
        {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }

        // shared::cluster address of `p` (a shared-memory object of this CTA) as seen in CTA `rank` of the cluster
        __device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank)
        {
            uint32_t r;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
            return r;
        }

        // arrive on an mbarrier that may live in another CTA of the cluster (address from mapa_u32)
        __device__ __forceinline__ void mbar_arrive_cluster(uint32_t clusterAddr)
        {
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(clusterAddr) : "memory");
        }

        // try_wait with a suspend-time hint: a waiting warp sleeps in hardware (up to `hintNs`) instead of spinning, so the
        // ~20 mostly-waiting warps of a CTA do not eat the issue slots the MMA-issuing warp needs.
        __device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity, uint32_t hintNs = 20000)
        {
            uint32_t done;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(parity), "r"(hintNs)
                : "memory");
            return done;
        }

        // Bounded wait: a protocol bug must surface as a trapped kernel (an error the host sees), never as a hung GPU.
        // The bound is a poll count (no clock reads in the loop): each poll sleeps up to the hint, or returns within
        // ~100 cycles if the hardware ignores it, so 2^22 polls is >= ~0.2 s of pure spinning and, as measured on
        // B200 where a poll sleeps ~3.7 us, ~15 s -- far beyond anything a legitimate wait sees even under time-slicing, MPS
        // or profiler replay (a trap poisons the context, so the bound must only ever fire on a genuine deadlock);
        // -DNB200_MBAR_MAX_POLLS=... overrides it.
#ifndef NB200_MBAR_MAX_POLLS
#define NB200_MBAR_MAX_POLLS (1u << 22)
#endif
        constexpr uint32_t kMbarMaxPolls = NB200_MBAR_MAX_POLLS;
        __device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
        {
            uint32_t polls = 0;
            while (!mbar_try_wait(bar, parity))
            {
                if (++polls > kMbarMaxPolls)
                {
                    printf("nb200: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
                    __trap();
                }
            }
        }


        // ---- the same primitives on 32-bit shared-window addresses ----
        // Kernels compute smem_u32() of each barrier array ONCE: passing generic pointers makes ptxas rebuild the generic
        // address (S2UR CgaCtaId/SWINHI, ULEA ...) and convert it back around every call, ~15 instructions in loops whose
        // whole body is ~100 (profiles/: the converter warps are instruction-issue-bound).
        __device__ __forceinline__ void mbar_arrive(uint32_t bar)
        {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
        __device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        }
        __device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity, uint32_t hintNs = 20000)
        {
            uint32_t done;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar), "r"(parity), "r"(hintNs)
                : "memory");
            return done;
        }
        __device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
        {
            uint32_t polls = 0;
            while (!mbar_try_wait(bar, parity))
            {
                if (++polls > kMbarMaxPolls)
                {
                    printf("nb200: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
                    __trap();
                }
            }
        }

        // ---------------- TMA ----------------
        __device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m)
        {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
        }

        __device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2)
        {
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                : "memory");
        }

        __device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3)
        {
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                : "memory");
        }

        // CTA-pair form: data lands in this CTA's shared memory, the byte count is signalled on an mbarrier given by its
        // shared::cluster address (the pair leader's barrier).
        __device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint32_t barClusterAddr, int c0, int c1, int c2)
        {
            asm volatile(
                "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(dst)), "l"((uint64_t)m), "r"(barClusterAddr), "r"(c0), "r"(c1), "r"(c2)
                : "memory");
        }

        // ---------------- tcgen05: TMEM allocation ----------------
        // Whole warp, converged. Writes the TMEM base address (lane 0, column base) to *slot in shared memory.
        __device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t columns)
        {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(columns) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }

        __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t columns)
        {
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(columns) : "memory");
        }

        // CTA-pair allocation: the same warp index of BOTH CTAs of the pair executes these.
        __device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot, uint32_t columns)
        {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(columns) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }

        __device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t columns)
        {
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(columns) : "memory");
        }

        __device__ __forceinline__ void tc_fence_before_sync()
        {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }

        __device__ __forceinline__ void tc_fence_after_sync()
        {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }

        // ---------------- tcgen05: MMA ----------------
        // D[tmem] (+)= A[smem] * B[smem], TF32 inputs, fp32 accumulate. One thread issues for the CTA.
        __device__ __forceinline__ void mma_tf32_ss(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
        {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate)
                : "memory");
        }

        // Same with the A operand read from tensor memory (128 lanes = M rows, 8 consecutive 32-bit columns = K).
        __device__ __forceinline__ void mma_tf32_ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
        {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmemD), "r"(tmemA), "l"(descB), "r"(idesc), "r"(accumulate)
                : "memory");
        }

        // CTA-pair MMA (issued by the pair leader only): D[256 x N] with rows 0-127 in the leader's TMEM and 128-255 in the
        // peer's; A likewise from each CTA's own TMEM; B's N rows are split, N/2 in each CTA's shared memory at `descB`.
        __device__ __forceinline__ void mma_tf32_ts_2sm(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
        {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmemD), "r"(tmemA), "l"(descB), "r"(idesc), "r"(accumulate)
                : "memory");
        }

        // Pair commit: arrives on the mbarrier at the same shared-memory offset in every CTA of `ctaMask`.
        __device__ __forceinline__ void mma_commit_2sm(uint64_t* bar, uint16_t ctaMask)
        {
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(bar)), "h"(ctaMask) : "memory");
        }

        // Arrive on an mbarrier once every MMA issued so far by this thread has completed
        // (implies tcgen05.fence::before_thread_sync).
        __device__ __forceinline__ void mma_commit(uint64_t* bar)
        {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        __device__ __forceinline__ void mma_commit(uint32_t bar)
        {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }

        // ---------------- tcgen05: TMEM -> registers ----------------
        // 32 lanes x 32 consecutive columns: thread t of the warp receives columns [col, col+32) of TMEM lane (base lane + t).
        __device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32])
        {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
        }

        // registers -> TMEM, same shape: thread t writes columns [col, col+32) of TMEM lane (base lane + t).
        // 32 lanes x 16 consecutive columns
        __device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16])
        {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
        }

        __device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32])
        {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                  "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                  "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                : "memory");
        }

        __device__ __forceinline__ void tmem_st_wait()
        {
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }

        __device__ __forceinline__ void tmem_ld_wait()
        {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }

        // ---------------- UMMA descriptors ----------------
        // Shared-memory matrix descriptor (sm_100 "version 1"): start address, leading / stride byte offsets
        // (all >> 4), swizzle mode in bits [61,64). 128-byte swizzle = 2.
        __device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smemAddr, uint32_t leadingBytes, uint32_t strideBytes)
        {
            uint64_t d = 0;
            d |= (uint64_t)((smemAddr >> 4) & 0x3FFF);
            d |= (uint64_t)((leadingBytes >> 4) & 0x3FFF) << 16;
            d |= (uint64_t)((strideBytes >> 4) & 0x3FFF) << 32;
            d |= (uint64_t)1 << 46; // descriptor version (Blackwell)
            d |= (uint64_t)2 << 61; // SWIZZLE_128B
            return d;
        }

        // Same for the narrower K-major swizzle modes (rows of 64 / 32 bytes): layout type 4 = SWIZZLE_64B, 6 = SWIZZLE_32B.
        __device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t smemAddr, uint32_t strideBytes, uint32_t layoutType)
        {
            uint64_t d = 0;
            d |= (uint64_t)((smemAddr >> 4) & 0x3FFF);
            d |= (uint64_t)1 << 16;
            d |= (uint64_t)((strideBytes >> 4) & 0x3FFF) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)layoutType << 61;
            return d;
        }

        // Instruction descriptor for kind::tf32 with fp32 accumulation.
        // aMnMajor/bMnMajor: 1 when the operand's M (resp. N) dimension is the contiguous one in shared memory.
        __host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int aMnMajor, int bMnMajor)
        {
            return (1u << 4)                    // D format: F32
                   | (2u << 7)                  // A format: TF32
                   | (2u << 10)                 // B format: TF32
                   | ((uint32_t)aMnMajor << 15) // A major
                   | ((uint32_t)bMnMajor << 16) // B major
                   | ((uint32_t)(N >> 3) << 17)
                   | ((uint32_t)(M >> 4) << 24);
        }
    }
}
