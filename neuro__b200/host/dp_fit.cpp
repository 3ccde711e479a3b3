// dp_fit -- batch-sharded data-parallel Fit() of a convolution stack, driven by C++ host code over the C ABI.
//
// The north-star shape of the drop-in (BASELINE.json): "batch-sharded data parallelism in ModelBase::Fit across the 8 GPUs of
// one box ... driven by C++ host code through a thin extern "C" shim". The reference's Fit is single-device
// (Neuro/src/Models/ModelBase.cpp:685-892; TrainStep :1035-1043; Adam::MinimizationOperation::ComputeInternal
// Neuro/src/Optimizers/Adam.cpp:66-111). Here ONE process drives N devices, one host thread per device (the C ABI keeps its
// per-device state per device ordinal and its error slot per thread), each thread holding a full replica:
//
//   forward   Conv2DBiasActivation per layer (+ Pool2D after each VGG block)          nb200_conv2d_forward, nb200_pool2d
//   loss      MSE against a synthetic target; its gradient (y - t) is formed with nb200_sgd_step used as an axpy, the
//             2 / (global count) factor is folded into the optimiser's grad_scale
//   backward  activation + bias gradient (one pass), kernel gradient, input gradient, pool gradient, last layer first
//   exchange  ncclAllReduce(sum) per bucket of the flat gradient buffer on a second stream, issued as soon as the bucket's
//             first layer has produced its kernel gradient (event dependency), i.e. underneath the remaining backward kernels
//   update    nb200_adam_step over the flat parameter buffer (grad_scale = 2 / global element count), every replica identical
//
// No Python, no ctypes, no per-call allocation: this is the number to compare with bench.py's when judging host overhead.
// The same step in Python (neuro__b200/fit.py) is what the parity and loss-curve tests exercise; this driver checks itself by
// comparing the replicas' parameters bit for bit after the run and the loss trajectory against a single-device run (--check).
//
//   dp_fit [--gpus N] [--steps K] [--warmup W] [--batch B] [--model vgg16|dcgan_d] [--res R] [--check]
// prints one JSON line.
#include <cuda_runtime.h>
#include <nccl.h>
#include <pthread.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "neuro_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
#define NCK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { fprintf(stderr, "NCCL error %s at %s:%d\n", ncclGetErrorString(r_), __FILE__, __LINE__); exit(2); } } while (0)
#define NB(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "nb200 error %d (%s) at %s:%d\n", r_, nb200_last_error(), __FILE__, __LINE__); exit(2); } } while (0)

struct LayerSpec { bool pool; int filters, filter, stride, pad, act; float alpha; };

struct Layer
{
    LayerSpec s;
    int C, H, W, K, Ho, Wo;
    size_t wOff = 0, bOff = 0;           // into the flat parameter / gradient buffers
    float *y = nullptr, *dy = nullptr, *dz = nullptr;
    nb200_conv_desc cd{};
    nb200_pool_desc pd{};
};

struct Bucket { size_t lo, hi; int layer; cudaEvent_t ready; };

static std::vector<LayerSpec> model(const std::string& name)
{
    std::vector<LayerSpec> l;
    if (name == "vgg16")
    {   // Neuro/src/Applications/VGG16.cpp:73-91
        const int blocks[5] = {2, 2, 3, 3, 3}, width[5] = {64, 128, 256, 512, 512};
        for (int b = 0; b < 5; ++b)
        {
            for (int i = 0; i < blocks[b]; ++i) l.push_back({false, width[b], 3, 1, 1, NB200_ACT_RELU, 0.f});
            l.push_back({true, 0, 2, 2, 0, 0, 0.f});
        }
    }
    else
    {   // DCGAN discriminator on CIFAR shapes, Neuro.Examples/src/CifarGAN.cpp:11-36
        l.push_back({false, 64, 3, 2, 1, NB200_ACT_LEAKY_RELU, 0.2f});
        l.push_back({false, 128, 3, 2, 1, NB200_ACT_LEAKY_RELU, 0.2f});
        l.push_back({false, 128, 3, 2, 1, NB200_ACT_LEAKY_RELU, 0.2f});
        l.push_back({false, 256, 3, 1, 1, NB200_ACT_LEAKY_RELU, 0.2f});
    }
    return l;
}

// counter-based uniform in [-1, 1): the same sequence on every replica and in every run
static inline float urand(uint64_t seed, uint64_t i)
{
    uint64_t z = (seed << 32) + i + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (float)((z >> 40) * (1.0 / 8388608.0) - 1.0);
}

struct Replica
{
    int dev, rank, world, batch;
    std::vector<Layer> layers;
    std::vector<Bucket> buckets;
    size_t nparams = 0, wsBytes = 0;
    float *params, *grads, *m, *v, *x, *dx, *target, *ws;
    cudaStream_t st, comm;
    ncclComm_t nccl;
    cudaEvent_t e0, e1, exchanged;
    double flopsPerStep = 0;
    int iteration = 0;
    long long outCount = 0;

    void build(const std::string& name, int res, int inC, size_t bucketBytes)
    {
        CK(cudaSetDevice(dev));
        CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&comm, cudaStreamNonBlocking));
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&exchanged, cudaEventDisableTiming));
        int C = inC, H = res, W = res;
        for (const LayerSpec& s : model(name))
        {
            Layer L; L.s = s; L.C = C; L.H = H; L.W = W;
            L.K = s.pool ? C : s.filters;
            L.Ho = (H + 2 * s.pad - s.filter) / s.stride + 1; L.Wo = (W + 2 * s.pad - s.filter) / s.stride + 1;
            if (!s.pool)
            {
                L.wOff = nparams; nparams += (size_t)L.K * C * s.filter * s.filter;
                L.bOff = nparams; nparams += L.K;
                L.cd = nb200_conv_desc{batch, C, H, W, L.K, s.filter, s.filter, L.Ho, L.Wo, s.stride, s.pad, s.pad, NB200_NCHW, NB200_MATH_TF32};
                for (int op = 0; op < 3; ++op) { const size_t b = nb200_conv2d_workspace_bytes(op, &L.cd); if (b > wsBytes) wsBytes = b; }
                const size_t b = nb200_conv2d_bias_activation_gradient_workspace_bytes(&L.cd); if (b > wsBytes) wsBytes = b;
                flopsPerStep += 3 * 2.0 * batch * L.K * L.Ho * L.Wo * C * s.filter * s.filter;
            }
            else
            {
                L.pd = nb200_pool_desc{batch, C, H, W, L.Ho, L.Wo, s.filter, s.stride, s.pad, s.pad, NB200_POOL_MAX, NB200_NCHW};
                const size_t b = nb200_pool2d_gradient_activation_workspace_bytes(&L.pd); if (b > wsBytes) wsBytes = b;
            }
            const size_t n = (size_t)batch * L.K * L.Ho * L.Wo;
            CK(cudaMalloc(&L.y, n * 4)); CK(cudaMalloc(&L.dy, n * 4));
            if (!s.pool) CK(cudaMalloc(&L.dz, n * 4));
            layers.push_back(L);
            C = L.K; H = L.Ho; W = L.Wo;
        }
        outCount = (long long)C * H * W;
        for (float** p : {&params, &grads, &m, &v}) { CK(cudaMalloc(p, nparams * 4)); CK(cudaMemsetAsync(*p, 0, nparams * 4, st)); }
        CK(cudaMalloc(&x, (size_t)batch * inC * res * res * 4)); CK(cudaMalloc(&dx, (size_t)batch * inC * res * res * 4));
        CK(cudaMalloc(&target, (size_t)batch * outCount * 4));
        CK(cudaMalloc(&ws, wsBytes ? wsBytes : 16));
        // Glorot-uniform kernels (VarianceScaling.cpp:59-65), zero biases (Conv2D.h:46): identical on every replica
        std::vector<float> h(nparams, 0.f);
        uint64_t seed = 1337;
        for (const Layer& L : layers)
            if (!L.s.pool)
            {
                const int f2 = L.s.filter * L.s.filter;
                const float limit = std::sqrt(6.0f / (float)(L.C * f2 + L.K * f2));
                const size_t n = (size_t)L.K * L.C * f2;
                for (size_t i = 0; i < n; ++i) h[L.wOff + i] = urand(seed, i) * limit;
                ++seed;
            }
        CK(cudaMemcpyAsync(params, h.data(), nparams * 4, cudaMemcpyHostToDevice, st));
        // this replica's shard of the synthetic batch (global sample index = rank * batch + i)
        const size_t per = (size_t)inC * res * res;
        std::vector<float> hx((size_t)batch * per), ht((size_t)batch * outCount);
        for (int i = 0; i < batch; ++i)
        {
            const uint64_t g = (uint64_t)rank * batch + i;
            for (size_t j = 0; j < per; ++j) hx[i * per + j] = urand(11 + g, j);
            for (long long j = 0; j < outCount; ++j) ht[i * outCount + j] = 0.5f * urand(5000 + g, j);
        }
        CK(cudaMemcpyAsync(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(target, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        // buckets over the flat gradient buffer, filled from the end (backward order)
        size_t hi = nparams, lo = nparams; int first = -1;
        for (int i = (int)layers.size() - 1; i >= 0; --i)
        {
            if (layers[i].s.pool) continue;
            lo = layers[i].wOff; first = i;
            if ((hi - lo) * 4 >= bucketBytes) { buckets.push_back({lo, hi, i, nullptr}); hi = lo; first = -1; }
        }
        if (first >= 0 && hi > lo) buckets.push_back({lo, hi, first, nullptr});
        // keep the bucket that ends the backward pass small (<= 1 MB): its all-reduce is the exposed tail of the exchange (see fit.py)
        if (!buckets.empty() && (buckets.back().hi - buckets.back().lo) * 4 > (1u << 20))
        {
            const size_t blo = buckets.back().lo, bhi = buckets.back().hi;
            size_t cut = 0; int next = -1;
            for (size_t i = 0; i < layers.size(); ++i)
            {
                if (layers[i].s.pool) continue;
                const size_t end = layers[i].bOff + layers[i].K;
                if (layers[i].wOff >= bhi || (end - blo) * 4 > (1u << 20)) break;
                cut = end;
            }
            if (cut > blo && cut < bhi)
            {
                for (size_t i = 0; i < layers.size(); ++i)
                    if (!layers[i].s.pool && layers[i].wOff >= cut) { next = (int)i; break; }
                int firstLayer = 0;
                for (size_t i = 0; i < layers.size(); ++i) if (!layers[i].s.pool) { firstLayer = (int)i; break; }
                buckets.back() = {cut, bhi, next, nullptr};
                buckets.push_back({blo, cut, firstLayer, nullptr});
            }
        }
        for (Bucket& b : buckets) CK(cudaEventCreateWithFlags(&b.ready, cudaEventDisableTiming));
    }

    void step(float lr)
    {
        const float* in = x;
        for (Layer& L : layers)
        {
            if (L.s.pool) NB(nb200_pool2d(&L.pd, in, L.y, st));
            else NB(nb200_conv2d_forward(&L.cd, in, params + L.wOff, params + L.bOff, L.s.act, L.s.alpha, L.y, ws, wsBytes, st));
            in = L.y;
        }
        // gradient of the summed squared error w.r.t. the output: dy = y - target (the 2 / count factor rides on grad_scale)
        Layer& last = layers.back();
        const size_t nOut = (size_t)batch * outCount;
        CK(cudaMemcpyAsync(last.dy, last.y, nOut * 4, cudaMemcpyDeviceToDevice, st));
        NB(nb200_sgd_step(last.dy, target, nOut, 1.f, 1.f, st));
        std::vector<char> fused(layers.size(), 0);
        for (int i = (int)layers.size() - 1; i >= 0; --i)
        {
            Layer& L = layers[i];
            const float* xin = i ? layers[i - 1].y : x;
            float* dxo = i ? layers[i - 1].dy : dx;
            if (L.s.pool)
            {
                if (i > 0 && !layers[i - 1].s.pool && nb200_pool2d_gradient_activation_supported(&L.pd))
                {   // "fused conv layer -> max pooling": pooling, activation and bias gradients in one pass
                    Layer& P = layers[i - 1];
                    NB(nb200_pool2d_gradient_activation(&L.pd, P.s.act, P.s.alpha, L.y, xin, L.dy, P.dz, grads + P.bOff, ws, wsBytes, st));
                    fused[i - 1] = true;
                }
                else
                    NB(nb200_pool2d_gradient(&L.pd, L.y, xin, L.dy, dxo, st));
                continue;
            }
            if (!fused[i])
                NB(nb200_conv2d_bias_activation_gradient(&L.cd, L.s.act, L.s.alpha, L.y, L.dy, L.dz, grads + L.bOff, ws, wsBytes, st));
            NB(nb200_conv2d_kernels_gradient(&L.cd, xin, L.dz, grads + L.wOff, nullptr, ws, wsBytes, st));
            if (world > 1)
                for (Bucket& b : buckets)
                    if (b.layer == i)
                    {
                        CK(cudaEventRecord(b.ready, st));
                        CK(cudaStreamWaitEvent(comm, b.ready, 0));
                        NCK(ncclAllReduce(grads + b.lo, grads + b.lo, b.hi - b.lo, ncclFloat, ncclSum, nccl, comm));
                    }
            NB(nb200_conv2d_input_gradient(&L.cd, L.dz, params + L.wOff, dxo, ws, wsBytes, st));
        }
        if (world > 1) { CK(cudaEventRecord(exchanged, comm)); CK(cudaStreamWaitEvent(st, exchanged, 0)); }
        ++iteration;
        const float lr_t = lr * std::sqrt(1.f - std::pow(0.999f, (float)iteration)) / (1.f - std::pow(0.9f, (float)iteration)); // Adam.cpp:90
        const float scale = 2.f / (float)((double)batch * world * outCount);
        NB(nb200_adam_step(params, grads, m, v, nparams, scale, lr_t, 0.9f, 0.999f, 1e-8f, st));
    }

    double loss() // sum over this shard of (y - t)^2 / global count; host-side, outside any timed region
    {
        Layer& last = layers.back();
        const size_t n = (size_t)batch * outCount;
        std::vector<float> y(n), t(n);
        CK(cudaMemcpyAsync(y.data(), last.y, n * 4, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(t.data(), target, n * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        double s = 0;
        for (size_t i = 0; i < n; ++i) s += ((double)y[i] - t[i]) * ((double)y[i] - t[i]);
        return s / ((double)batch * world * outCount);
    }
};

int main(int argc, char** argv)
{
    int gpus = 1, steps = 20, warmup = 3, batch = 8, res = 512; bool check = false; std::string name = "vgg16";
    for (int i = 1; i < argc; ++i)
    {
        auto val = [&](int& v) { if (i + 1 < argc) v = atoi(argv[++i]); };
        if (!strcmp(argv[i], "--gpus")) val(gpus);
        else if (!strcmp(argv[i], "--steps")) val(steps);
        else if (!strcmp(argv[i], "--warmup")) val(warmup);
        else if (!strcmp(argv[i], "--batch")) val(batch);
        else if (!strcmp(argv[i], "--res")) val(res);
        else if (!strcmp(argv[i], "--model") && i + 1 < argc) name = argv[++i];
        else if (!strcmp(argv[i], "--check")) check = true;
    }
    if (name == "dcgan_d" && res == 512) res = 32;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < gpus) { fprintf(stderr, "dp_fit needs %d CUDA devices, found %d (there is no CPU fallback)\n", gpus, ndev); return 2; }
    if (!getenv("NCCL_MAX_CTAS")) setenv("NCCL_MAX_CTAS", "4", 1);   // see bench.py: NCCL CTAs compete with one-CTA-per-SM conv grids

    auto run = [&](int world, std::vector<double>* lossOut, std::vector<std::vector<float>>* paramsOut, double* msOut, double* flopsOut) {
        std::vector<Replica> reps(world);
        std::vector<ncclComm_t> comms(world);
        if (world > 1) { std::vector<int> devs(world); for (int i = 0; i < world; ++i) devs[i] = i; NCK(ncclCommInitAll(comms.data(), world, devs.data())); }
        pthread_barrier_t bar; pthread_barrier_init(&bar, nullptr, world);
        std::vector<double> ms(world, 0), losses(world * (check ? steps : 1), 0);
        std::vector<std::thread> th;
        for (int r = 0; r < world; ++r)
            th.emplace_back([&, r] {
                Replica& R = reps[r];
                R.dev = r; R.rank = r; R.world = world; R.batch = batch; R.nccl = world > 1 ? comms[r] : nullptr;
                R.build(name, res, 3, 24u << 20);
                for (int i = 0; i < warmup; ++i) R.step(1e-4f);
                CK(cudaStreamSynchronize(R.st));
                pthread_barrier_wait(&bar);
                CK(cudaEventRecord(R.e0, R.st));
                for (int i = 0; i < steps; ++i)
                {
                    R.step(1e-4f);
                    if (check) losses[r * steps + i] = R.loss();
                }
                CK(cudaEventRecord(R.e1, R.st));
                CK(cudaStreamSynchronize(R.st));
                float t = 0; CK(cudaEventElapsedTime(&t, R.e0, R.e1)); ms[r] = t / steps;
                pthread_barrier_wait(&bar);
                if (paramsOut) { (*paramsOut)[r].resize(R.nparams); CK(cudaMemcpy((*paramsOut)[r].data(), R.params, R.nparams * 4, cudaMemcpyDeviceToHost)); }
            });
        for (auto& t : th) t.join();
        double worst = 0; for (double t : ms) worst = t > worst ? t : worst;
        *msOut = worst; *flopsOut = reps[0].flopsPerStep;
        if (lossOut && check) { lossOut->assign(steps, 0); for (int r = 0; r < world; ++r) for (int i = 0; i < steps; ++i) (*lossOut)[i] += losses[r * steps + i]; }
        if (world > 1) for (auto c : comms) ncclCommDestroy(c);
    };

    std::vector<std::vector<float>> params(gpus);
    std::vector<double> loss;
    double ms = 0, flops = 0;
    run(gpus, &loss, &params, &ms, &flops);
    bool identical = true;
    for (int r = 1; r < gpus; ++r) identical = identical && params[r] == params[0];
    printf("{\"driver\": \"dp_fit (C++ host threads over the C ABI, ncclCommInitAll)\", \"model\": \"%s\", \"res\": %d, \"n_gpus\": %d, \"per_gpu_batch\": %d, "
           "\"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.4f, \"samples_per_s\": %.2f, \"conv_tflops_per_gpu\": %.2f, \"replicas_identical\": %s",
           name.c_str(), res, gpus, batch, steps, warmup, ms, batch * gpus / (ms * 1e-3), flops / (ms * 1e-3) / 1e12, identical ? "true" : "false");
    if (check)
    {
        // the same GLOBAL batch on one device: batch * gpus samples; the loss trajectory must be the sharded run's
        std::vector<double> ref; std::vector<std::vector<float>> p1(1); double ms1 = 0, fl1 = 0;
        const int b0 = batch; batch = b0 * gpus;
        run(1, &ref, &p1, &ms1, &fl1);
        batch = b0;
        double worst = 0;
        for (int i = 0; i < steps; ++i) worst = std::fmax(worst, std::fabs(loss[i] - ref[i]) / std::fabs(ref[i]));
        printf(", \"loss_first\": %.6g, \"loss_last\": %.6g, \"max_rel_loss_diff_vs_single_device\": %.3g", loss[0], loss[steps - 1], worst);
    }
    printf("}\n");
    return identical ? 0 : 1;
}
