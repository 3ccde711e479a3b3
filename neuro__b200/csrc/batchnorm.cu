// Batch normalisation (training forward, gradient, inference) for the conv stacks of the GAN / pix2pix configs, with the
// statistics exposed so that batch-sharded replicas can normalise over the GLOBAL batch.
//
// Reference: TensorOpCpu::BatchNormalization / BatchNormalizationTrain / BatchNormalizationGradient
// (Neuro/src/Tensors/TensorOpCpu.cpp:1371-1480), layer Neuro/src/Layers/BatchNormalization.cpp. The three modes
// (EBatchNormMode, Types.h) are one layout here: the tensor is Nn x G x S with element (n, g, s) at (n*G + g)*S + s and
// statistics per group g over (n, s):   Spatial       Nn = N, G = C,     S = H*W   (one mean per channel; the conv case)
//                                       PerActivation Nn = N, G = C*H*W, S = 1
//                                       Instance      Nn = 1, G = N*C,   S = H*W
//
// Bound: HBM. Algorithmic bytes per element: forward 12 (x read for the moments, x read + y written by the apply pass),
// gradient 20 (x, dy read for the sums; x, dy read + dx written). The reference makes ~10 passes (mean, x - mean, sqr,
// mean, ..., each a full tensor). Moments: a block reads its <= 4096 elements ONCE into registers, takes its own mean and
// then its own sum of squared deviations from the registers (two-pass accuracy, one pass over HBM); blocks are combined
// with Chan's parallel formula in a fixed order -> deterministic, and more accurate than the reference's sequential
// fp32 sums (tests compare against the reference within a stated tolerance and against float64).
// Replicas: (mean, M2) of every replica are gathered (2*G floats each) and combined in rank order by the same formula, so
// every replica derives bit-identical statistics; the gradient needs three per-group sums, all-reduced as 3*G floats.
#include "common.cuh"

namespace nb200
{
    namespace
    {
        constexpr int kBnThreads = 256;
        constexpr int kBnPerThread = 16;
        constexpr int kBnChunk = kBnThreads * kBnPerThread; // elements of one group handled by one block

        struct BnLayout
        {
            int Nn, G, S;
            long long m;      // elements per group = Nn * S
            int chunks;       // blocks per group
            bool vec;         // S % 4 == 0: 16-byte accesses never straddle an (n, g) run
        };

        BnLayout bn_layout(const nb200_bn_desc& d)
        {
            BnLayout l{};
            const long long hw = (long long)d.H * d.W;
            if (d.mode == NB200_BN_SPATIAL) { l.Nn = d.N; l.G = d.C; l.S = (int)hw; }
            else if (d.mode == NB200_BN_PER_ACTIVATION) { l.Nn = d.N; l.G = (int)(d.C * hw); l.S = 1; }
            else { l.Nn = 1; l.G = d.N * d.C; l.S = (int)hw; }
            l.m = (long long)l.Nn * l.S;
            l.chunks = (int)((l.m + kBnChunk - 1) / kBnChunk);
            l.vec = (l.S & 3) == 0;
            return l;
        }

        __device__ __forceinline__ float warp_sum(float v)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                v += __shfl_xor_sync(0xffffffffu, v, o);
            return v;
        }

        // sum over the block, result in every thread; fixed tree => deterministic
        __device__ __forceinline__ float block_sum(float v, float* red)
        {
            v = warp_sum(v);
            __syncthreads(); // red may still be read from a previous call
            if ((threadIdx.x & 31) == 0)
                red[threadIdx.x >> 5] = v;
            __syncthreads();
            float t = red[0];
#pragma unroll
            for (int i = 1; i < kBnThreads / 32; ++i)
                t += red[i];
            return t;
        }

        // flattened index j in [0, Nn*S) of group g -> element offset
        __device__ __forceinline__ long long bn_offset(long long j, int g, int G, int S)
        {
            const long long n = j / S;
            return (n * G + g) * (long long)S + (j - n * S);
        }

        // Load this block's chunk of group g into registers. VEC: thread owns 4 float4 (16 consecutive... 4 x 4 values).
        template <bool VEC>
        __device__ __forceinline__ int bn_load_chunk(const float* __restrict__ x, int g, int G, int S, long long m, long long j0, float (&v)[kBnPerThread])
        {
            int cnt = 0;
            if (VEC)
            {
#pragma unroll
                for (int i = 0; i < kBnPerThread / 4; ++i)
                {
                    const long long j = j0 + ((long long)i * kBnThreads + threadIdx.x) * 4;
                    if (j < m)
                    {
                        const float4 q = __ldg(reinterpret_cast<const float4*>(x + bn_offset(j, g, G, S)));
                        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
                        cnt += 4;
                    }
                    else
                        v[4 * i] = v[4 * i + 1] = v[4 * i + 2] = v[4 * i + 3] = 0.f;
                }
            }
            else
            {
#pragma unroll
                for (int i = 0; i < kBnPerThread; ++i)
                {
                    const long long j = j0 + (long long)i * kBnThreads + threadIdx.x;
                    if (j < m) { v[i] = __ldg(x + bn_offset(j, g, G, S)); ++cnt; }
                    else v[i] = 0.f;
                }
            }
            return cnt;
        }

        // partial[(g * chunks + b) * 2 + {0,1}] = (mean, M2) of block b's elements of group g
        template <bool VEC>
        __global__ void __launch_bounds__(kBnThreads)
        bn_moments_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int G, int S, long long m, int chunks)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            __shared__ float red[kBnThreads / 32];
            const int g = blockIdx.x, b = blockIdx.y;
            const long long j0 = (long long)b * kBnChunk;
            float v[kBnPerThread];
            bn_load_chunk<VEC>(x, g, G, S, m, j0, v);
            const long long left = m - j0;
            const float cnt = (float)(left < kBnChunk ? left : kBnChunk);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kBnPerThread; ++i)
                s += v[i]; // slots past the end hold 0
            const float mean = block_sum(s, red) / cnt;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < kBnPerThread; ++i)
            {
                const long long j = VEC ? j0 + ((long long)(i / 4) * kBnThreads + threadIdx.x) * 4 + (i & 3) : j0 + (long long)i * kBnThreads + threadIdx.x;
                const float dlt = v[i] - mean;
                q += j < m ? dlt * dlt : 0.f;
            }
            const float M2 = block_sum(q, red);
            if (threadIdx.x == 0)
            {
                partial[((long long)g * chunks + b) * 2] = mean;
                partial[((long long)g * chunks + b) * 2 + 1] = M2;
            }
        }

        // Chan et al.: merge (na, meanA, M2a) with (nb, meanB, M2b)
        __device__ __forceinline__ void chan_merge(float& na, float& meanA, float& M2a, float nb, float meanB, float M2b)
        {
            if (nb == 0.f) return;
            if (na == 0.f) { na = nb; meanA = meanB; M2a = M2b; return; }
            const float n = na + nb;
            const float dlt = meanB - meanA;
            meanA = meanA + dlt * (nb / n);
            M2a = M2a + M2b + dlt * dlt * (na * nb / n);
            na = n;
        }

        // One warp per group: lanes take interleaved partials sequentially, then a fixed shuffle tree. Every part has
        // `cntFull` elements except the last of each source (cntLast). moments[g] = (mean, M2).
        __global__ void bn_moments_combine_kernel(const float* __restrict__ partial, int G, int parts, float cntFull, float cntLast,
                                                  float* __restrict__ moments)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
            const int lane = threadIdx.x & 31;
            if (g >= G) return;
            float n = 0.f, mean = 0.f, M2 = 0.f;
            for (int b = lane; b < parts; b += 32)
                chan_merge(n, mean, M2, b == parts - 1 ? cntLast : cntFull, partial[((long long)g * parts + b) * 2], partial[((long long)g * parts + b) * 2 + 1]);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const float n2 = __shfl_xor_sync(0xffffffffu, n, o), mean2 = __shfl_xor_sync(0xffffffffu, mean, o), M22 = __shfl_xor_sync(0xffffffffu, M2, o);
                // the lower lane of each pair merges (a, b) in lane order so that both lanes hold the same value
                if (lane & o)
                {
                    float na = n2, ma = mean2, qa = M22;
                    chan_merge(na, ma, qa, n, mean, M2);
                    n = na; mean = ma; M2 = qa;
                }
                else
                    chan_merge(n, mean, M2, n2, mean2, M22);
            }
            if (lane == 0)
            {
                moments[2 * g] = mean;
                moments[2 * g + 1] = M2;
            }
        }

        // Statistics of the global batch from the gathered per-replica moments (rank order), running statistics update:
        //   saveMean = mean; saveInvVar = 1 / sqrt(M2/m + eps); running = (1-momentum)*running + momentum*{mean, var*m/(m-1)}
        // (TensorOpCpu.cpp:1417-1431). all[r][g] = (mean, M2) with mLocal elements each.
        __global__ void bn_finalize_kernel(const float* __restrict__ all, int replicas, int G, float mLocal, float momentum, float epsilon,
                                           float* __restrict__ runningMean, float* __restrict__ runningVar, float* __restrict__ saveMean,
                                           float* __restrict__ saveInvVar)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int g = blockIdx.x * blockDim.x + threadIdx.x;
            if (g >= G) return;
            float n = 0.f, mean = 0.f, M2 = 0.f;
            for (int r = 0; r < replicas; ++r)
                chan_merge(n, mean, M2, mLocal, all[((long long)r * G + g) * 2], all[((long long)r * G + g) * 2 + 1]);
            const float var = M2 / n;
            saveMean[g] = mean;
            saveInvVar[g] = 1.f / sqrtf(var + epsilon);
            if (runningMean)
                runningMean[g] = (1.f - momentum) * runningMean[g] + momentum * mean;
            if (runningVar)
                runningVar[g] = (1.f - momentum) * runningVar[g] + momentum * (var * (n / (n - 1.f)));
        }

        // y = ((x - mean) * inv) * gamma + beta, each step rounded as the reference's separate passes round it
        // (TensorOpCpu.cpp:1422-1424 and :1381-1388 for inference, where inv is derived from the running variance).
        template <bool VEC, bool INFER>
        __global__ void __launch_bounds__(kBnThreads)
        bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                        const float* __restrict__ invOrVar, float epsilon, float* __restrict__ y, long long total, int G, int S)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const long long stride = (long long)gridDim.x * blockDim.x;
            if (VEC)
            {
                for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < total; i += stride)
                {
                    const int g = (int)(((i * 4) / S) % G);
                    const float mu = __ldg(mean + g), ga = __ldg(gamma + g), be = __ldg(beta + g);
                    const float inv = INFER ? 1.f / sqrtf(__ldg(invOrVar + g) + epsilon) : __ldg(invOrVar + g);
                    const float4 q = __ldcs(reinterpret_cast<const float4*>(x) + i);
                    float4 o;
                    o.x = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(q.x, mu), inv), ga), be);
                    o.y = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(q.y, mu), inv), ga), be);
                    o.z = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(q.z, mu), inv), ga), be);
                    o.w = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(q.w, mu), inv), ga), be);
                    reinterpret_cast<float4*>(y)[i] = o;
                }
            }
            else
            {
                for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
                {
                    const int g = (int)((i / S) % G);
                    const float inv = INFER ? 1.f / sqrtf(__ldg(invOrVar + g) + epsilon) : __ldg(invOrVar + g);
                    y[i] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__ldg(x + i), __ldg(mean + g)), inv), __ldg(gamma + g)), __ldg(beta + g));
                }
            }
        }

        // partial[(g * chunks + b) * 3 + {0,1,2}] = sum dy, sum dy*(x - mean), sum (x - mean) over block b's elements
        template <bool VEC>
        __global__ void __launch_bounds__(kBnThreads)
        bn_gradient_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ mean, float* __restrict__ partial,
                                   int G, int S, long long m, int chunks)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            __shared__ float red[kBnThreads / 32];
            const int g = blockIdx.x, b = blockIdx.y;
            const long long j0 = (long long)b * kBnChunk;
            const float mu = __ldg(mean + g);
            float sd = 0.f, sdx = 0.f, sx = 0.f;
            if (VEC)
            {
#pragma unroll
                for (int i = 0; i < kBnPerThread / 4; ++i)
                {
                    const long long j = j0 + ((long long)i * kBnThreads + threadIdx.x) * 4;
                    if (j < m)
                    {
                        const long long o = bn_offset(j, g, G, S);
                        const float4 q = __ldg(reinterpret_cast<const float4*>(x + o)), e = __ldg(reinterpret_cast<const float4*>(dy + o));
                        const float a0 = q.x - mu, a1 = q.y - mu, a2 = q.z - mu, a3 = q.w - mu;
                        sd += (e.x + e.y) + (e.z + e.w);
                        sdx += (e.x * a0 + e.y * a1) + (e.z * a2 + e.w * a3);
                        sx += (a0 + a1) + (a2 + a3);
                    }
                }
            }
            else
            {
#pragma unroll
                for (int i = 0; i < kBnPerThread; ++i)
                {
                    const long long j = j0 + (long long)i * kBnThreads + threadIdx.x;
                    if (j < m)
                    {
                        const long long o = bn_offset(j, g, G, S);
                        const float a = __ldg(x + o) - mu, e = __ldg(dy + o);
                        sd += e; sdx += e * a; sx += a;
                    }
                }
            }
            const float t0 = block_sum(sd, red), t1 = block_sum(sdx, red), t2 = block_sum(sx, red);
            if (threadIdx.x == 0)
            {
                float* p = partial + ((long long)g * chunks + b) * 3;
                p[0] = t0; p[1] = t1; p[2] = t2;
            }
        }

        // sums[g*3 + {0,1,2}] = the three sums of group g (warp per group, interleaved sequential adds + fixed tree)
        __global__ void bn_gradient_combine_kernel(const float* __restrict__ partial, int G, int parts, float* __restrict__ sums)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
            const int lane = threadIdx.x & 31;
            if (g >= G) return;
            float a = 0.f, b = 0.f, c = 0.f;
            for (int i = lane; i < parts; i += 32)
            {
                const float* p = partial + ((long long)g * parts + i) * 3;
                a += p[0]; b += p[1]; c += p[2];
            }
            a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
            if (lane == 0)
            {
                sums[3 * g] = a; sums[3 * g + 1] = b; sums[3 * g + 2] = c;
            }
        }

        // BatchNormalizationGradient (TensorOpCpu.cpp:1465-1476) from the (globally summed) per-group sums; m = elements per
        // group over ALL replicas:
        //   dxNorm = dy*gamma;  dVar = sum(dxNorm*xMu) * -0.5 * inv^3;  dMu = sum(dxNorm * -inv) + dVar * mean(xMu * -2)
        //   dx = dxNorm*inv + dVar*xMu*2/m + dMu/m;   dgamma = sum(dy * xNorm) (local sums);  dbeta = sum(dy) (local sums)
        template <bool VEC>
        __global__ void __launch_bounds__(kBnThreads)
        bn_gradient_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma, const float* __restrict__ mean,
                                 const float* __restrict__ inv, const float* __restrict__ sums, float m, float* __restrict__ dx, long long total, int G, int S)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const long long stride = (long long)gridDim.x * blockDim.x;
            const float invm = 1.f / m;
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i * (VEC ? 4 : 1) < total; i += stride)
            {
                const int g = (int)(((i * (VEC ? 4 : 1)) / S) % G);
                const float mu = __ldg(mean + g), iv = __ldg(inv + g), ga = __ldg(gamma + g);
                const float sumA = ga * __ldg(sums + 3 * g + 1), sumB = -iv * ga * __ldg(sums + 3 * g), sumC = -2.f * __ldg(sums + 3 * g + 2);
                const float dVar = sumA * -.5f * (iv * iv * iv);
                const float dMuM = (sumB + dVar * (sumC * invm)) * invm;
                if (VEC)
                {
                    const float4 q = __ldcs(reinterpret_cast<const float4*>(x) + i), e = __ldcs(reinterpret_cast<const float4*>(dy) + i);
                    float4 o;
                    o.x = ((e.x * ga) * iv + ((dVar * (q.x - mu)) * 2.f) * invm) + dMuM;
                    o.y = ((e.y * ga) * iv + ((dVar * (q.y - mu)) * 2.f) * invm) + dMuM;
                    o.z = ((e.z * ga) * iv + ((dVar * (q.z - mu)) * 2.f) * invm) + dMuM;
                    o.w = ((e.w * ga) * iv + ((dVar * (q.w - mu)) * 2.f) * invm) + dMuM;
                    reinterpret_cast<float4*>(dx)[i] = o;
                }
                else
                    dx[i] = ((__ldg(dy + i) * ga) * iv + ((dVar * (__ldg(x + i) - mu)) * 2.f) * invm) + dMuM;
            }
        }

        // dgamma[g] = inv * local sum dy*xMu, dbeta[g] = local sum dy
        __global__ void bn_param_gradient_kernel(const float* __restrict__ localSums, const float* __restrict__ inv, int G, float* __restrict__ dgamma,
                                                 float* __restrict__ dbeta)
        {
            ptx::pdl_launch_dependents(); // programmatic dependent launch, see common.cuh
            ptx::pdl_wait();
            const int g = blockIdx.x * blockDim.x + threadIdx.x;
            if (g >= G) return;
            if (dgamma) dgamma[g] = __ldg(inv + g) * localSums[3 * g + 1];
            if (dbeta) dbeta[g] = localSums[3 * g];
        }

        inline unsigned grid_for(long long work, int threads)
        {
            long long blocks = (work + threads - 1) / threads;
            const long long cap = 148ll * 16;
            return (unsigned)(blocks < 1 ? 1 : blocks > cap ? cap : blocks);
        }

        int check_desc(const nb200_bn_desc* d)
        {
            if (!d || d->N < 0 || d->C < 0 || d->H < 0 || d->W < 0)
                return fail(NB200_E_INVALID, "bad batch-norm descriptor");
            if (d->mode < NB200_BN_PER_ACTIVATION || d->mode > NB200_BN_INSTANCE)
                return fail(NB200_E_INVALID, "unknown batch-norm mode %d", d->mode);
            const long long lim = 0xFFFFFFFFll;
            long long v = d->N;
            for (long long f : {(long long)d->C, (long long)d->H, (long long)d->W})
            {
                if (f != 0 && v > lim / f)
                    return fail(NB200_E_INVALID, "tensor exceeds 2^32-1 elements");
                v *= f;
            }
            return NB200_OK;
        }
    }

    size_t bn_workspace_bytes(const nb200_bn_desc& d)
    {
        const BnLayout l = bn_layout(d);
        // per-block partials (3 floats per block: the gradient's need covers the moments' 2) + one row of per-group sums
        return ((size_t)l.G * (l.chunks > 0 ? l.chunks : 1) * 3 + (size_t)l.G * 3) * sizeof(float);
    }

    int bn_moments(const nb200_bn_desc& d, const float* x, float* moments, void* ws, size_t wsBytes, cudaStream_t st)
    {
        const BnLayout l = bn_layout(d);
        if (l.G == 0) return NB200_OK;
        if (wsBytes < bn_workspace_bytes(d) || !ws)
            return fail(NB200_E_WORKSPACE, "batch norm needs %zu workspace bytes, got %zu", bn_workspace_bytes(d), wsBytes);
        float* partial = (float*)ws;
        const dim3 grid((unsigned)l.G, (unsigned)l.chunks);
        if (l.chunks > 65535)
            return fail(NB200_E_UNSUPPORTED, "more than 65535 x 4096 elements per normalisation group");
        if (l.vec && !((uintptr_t)x & 15))
            NB200_CUDA_TRY(launch_kernel(bn_moments_partial_kernel<true>, dim3(grid), dim3(kBnThreads), 0, st, x, partial, l.G, l.S, l.m, l.chunks));
        else
            NB200_CUDA_TRY(launch_kernel(bn_moments_partial_kernel<false>, dim3(grid), dim3(kBnThreads), 0, st, x, partial, l.G, l.S, l.m, l.chunks));
        NB200_CUDA_TRY(cudaGetLastError());
        const long long last = l.m - (long long)(l.chunks - 1) * kBnChunk;
        NB200_CUDA_TRY(launch_kernel(bn_moments_combine_kernel, dim3((unsigned)((l.G + 7) / 8)), dim3(256), 0, st, partial, l.G, l.chunks, (float)kBnChunk, (float)last, moments));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch(2);
        return NB200_OK;
    }

    int bn_finalize(const nb200_bn_desc& d, const float* allMoments, int replicas, float momentum, float epsilon, float* runningMean,
                    float* runningVar, float* saveMean, float* saveInvVar, cudaStream_t st)
    {
        const BnLayout l = bn_layout(d);
        if (l.G == 0) return NB200_OK;
        NB200_CUDA_TRY(launch_kernel(bn_finalize_kernel, dim3((unsigned)((l.G + 127) / 128)), dim3(128), 0, st, allMoments, replicas, l.G, (float)l.m, momentum, epsilon, runningMean, runningVar, saveMean, saveInvVar));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int bn_apply(const nb200_bn_desc& d, bool inference, const float* x, const float* gamma, const float* beta, const float* mean,
                 const float* invOrVar, float epsilon, float* y, cudaStream_t st)
    {
        const BnLayout l = bn_layout(d);
        const long long total = (long long)l.G * l.m;
        if (total == 0) return NB200_OK;
        const bool vec = l.vec && !(((uintptr_t)x | (uintptr_t)y) & 15);
        const unsigned grid = grid_for(vec ? total / 4 : total, kBnThreads);
        if (vec)
        {
            if (inference) NB200_CUDA_TRY(launch_kernel(bn_apply_kernel<true, true>, dim3(grid), dim3(kBnThreads), 0, st, x, gamma, beta, mean, invOrVar, epsilon, y, total, l.G, l.S));
            else NB200_CUDA_TRY(launch_kernel(bn_apply_kernel<true, false>, dim3(grid), dim3(kBnThreads), 0, st, x, gamma, beta, mean, invOrVar, epsilon, y, total, l.G, l.S));
        }
        else
        {
            if (inference) NB200_CUDA_TRY(launch_kernel(bn_apply_kernel<false, true>, dim3(grid), dim3(kBnThreads), 0, st, x, gamma, beta, mean, invOrVar, epsilon, y, total, l.G, l.S));
            else NB200_CUDA_TRY(launch_kernel(bn_apply_kernel<false, false>, dim3(grid), dim3(kBnThreads), 0, st, x, gamma, beta, mean, invOrVar, epsilon, y, total, l.G, l.S));
        }
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        return NB200_OK;
    }

    int bn_gradient_sums(const nb200_bn_desc& d, const float* x, const float* dy, const float* saveMean, float* sums, void* ws, size_t wsBytes,
                         cudaStream_t st)
    {
        const BnLayout l = bn_layout(d);
        if (l.G == 0) return NB200_OK;
        if (wsBytes < bn_workspace_bytes(d) || !ws)
            return fail(NB200_E_WORKSPACE, "batch norm needs %zu workspace bytes, got %zu", bn_workspace_bytes(d), wsBytes);
        if (l.chunks > 65535)
            return fail(NB200_E_UNSUPPORTED, "more than 65535 x 4096 elements per normalisation group");
        float* partial = (float*)ws;
        const dim3 grid((unsigned)l.G, (unsigned)l.chunks);
        if (l.vec && !(((uintptr_t)x | (uintptr_t)dy) & 15))
            NB200_CUDA_TRY(launch_kernel(bn_gradient_partial_kernel<true>, dim3(grid), dim3(kBnThreads), 0, st, x, dy, saveMean, partial, l.G, l.S, l.m, l.chunks));
        else
            NB200_CUDA_TRY(launch_kernel(bn_gradient_partial_kernel<false>, dim3(grid), dim3(kBnThreads), 0, st, x, dy, saveMean, partial, l.G, l.S, l.m, l.chunks));
        NB200_CUDA_TRY(cudaGetLastError());
        NB200_CUDA_TRY(launch_kernel(bn_gradient_combine_kernel, dim3((unsigned)((l.G + 7) / 8)), dim3(256), 0, st, partial, l.G, l.chunks, sums));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch(2);
        return NB200_OK;
    }

    int bn_gradient_apply(const nb200_bn_desc& d, int replicas, const float* x, const float* gamma, const float* dy, const float* saveMean,
                          const float* saveInvVar, const float* globalSums, const float* localSums, float* dgamma, float* dbeta, float* dx,
                          cudaStream_t st)
    {
        const BnLayout l = bn_layout(d);
        const long long total = (long long)l.G * l.m;
        if (total == 0) return NB200_OK;
        const float m = (float)l.m * (float)replicas;
        const bool vec = l.vec && !(((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15);
        const unsigned grid = grid_for(vec ? total / 4 : total, kBnThreads);
        if (vec)
            NB200_CUDA_TRY(launch_kernel(bn_gradient_apply_kernel<true>, dim3(grid), dim3(kBnThreads), 0, st, x, dy, gamma, saveMean, saveInvVar, globalSums, m, dx, total, l.G, l.S));
        else
            NB200_CUDA_TRY(launch_kernel(bn_gradient_apply_kernel<false>, dim3(grid), dim3(kBnThreads), 0, st, x, dy, gamma, saveMean, saveInvVar, globalSums, m, dx, total, l.G, l.S));
        NB200_CUDA_TRY(cudaGetLastError());
        count_launch();
        if (dgamma || dbeta)
        {
            NB200_CUDA_TRY(launch_kernel(bn_param_gradient_kernel, dim3((unsigned)((l.G + 127) / 128)), dim3(128), 0, st, localSums, saveInvVar, l.G, dgamma, dbeta));
            NB200_CUDA_TRY(cudaGetLastError());
            count_launch();
        }
        return NB200_OK;
    }

    int bn_check(const nb200_bn_desc* d) { return check_desc(d); }
    long long bn_group_elements(const nb200_bn_desc& d) { return bn_layout(d).m; }
    int bn_groups(const nb200_bn_desc& d) { return bn_layout(d).G; }
}
