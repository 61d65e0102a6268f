// sampler.cu -- device fast path for the reference sampler: sample() of src/sampler.c:186-201
// (temperature, softmax, nucleus / top-p, inverse CDF with a host-supplied coin) on the logits the decode step left
// in HBM, returning ONE token id instead of vocab_size floats. The reference copies 608 KB of logits to the host,
// scales, exponentiates and qsorts all 151 936 entries per token (~11 ms); only the head of the sorted order can ever
// be chosen, so the device sorts just that head.
//
// One CTA, 1024 threads, four passes over the logits in L2:
//   1. m = max(l_i / T)                                   (sampler.c:188-190, forward.c:36-50)
//   2. s = sum expf(l_i / T - m)                          (forward.c:53-69; tree order instead of serial)
//   3. p_i = expf(l_i / T - m) / s; candidates = { p_i >= cut } (tested in the log domain) with cut = 0.5 * (1 - top_p) / (V - 1): an entry
//      below (1 - top_p) / (V - 1) cannot lie in the smallest prefix whose mass exceeds top_p, so the candidates are
//      a prefix of the reference's sorted array that contains the whole nucleus; if more than kMaxCand entries pass,
//      the cut is raised (x4 per round) until they fit -- still a prefix, and step 4 checks that the nucleus closes in it
//   4. bitonic sort of the candidates (descending p, ascending index among equals -- the reference's order among
//      equals is whatever qsort does), then ONE thread walks them exactly like sampler_mass_index (sampler.c:88-113,
//      including the "heal" step) and sampler_cdf_index (sampler.c:126-136, inclusive bound and fallback).
// status 1 = not handled (more than kMaxCand candidates: a nearly flat distribution or top_p == 1): the caller then
// uses the reference's sample() on the host logits. Probabilities differ from the host's in the last bits (expf,
// summation order), so a decision within ~1e-6 of a boundary may fall on the neighbouring token; tests bound that.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 1024;
constexpr int kMaxCand = 4096;

__device__ __forceinline__ bool before(float pa, int ia, float pb, int ib) { return pa > pb || (pa == pb && ia < ib); }

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < kThreads / 32; ++i) r = is_max ? fmaxf(r, red[i]) : __fadd_rn(r, red[i]);
    return r;
}

__device__ __forceinline__ void sort_and_walk(float* sp, int* si, int n, float top_p, float coin, int* __restrict__ out, int* token_dev);

__global__ void __launch_bounds__(kThreads, 1)
k_sample(const float* __restrict__ logits, int V, float temperature, float top_p, float coin, int* __restrict__ out, int* token_dev) {
    __shared__ float red[32];
    __shared__ float sp[kMaxCand];
    __shared__ int si[kMaxCand];
    __shared__ int count;
    const int tid = threadIdx.x;
    if (tid == 0) count = 0;
    // 1. max of the scaled logits: division by T > 0 is monotone, so max(l_i / T) = max(l_i) / T exactly
    float m = -INFINITY;
    for (int i = tid; i < V; i += kThreads) m = fmaxf(m, logits[i]);
    m = __fdiv_rn(block_reduce(m, red, true), temperature);
    // 2. sum of exponentials
    float s = 0.0f;
    for (int i = tid; i < V; i += kThreads) s = __fadd_rn(s, expf(__fsub_rn(__fdiv_rn(logits[i], temperature), m)));
    s = block_reduce(s, red, false);
    // 3. candidates. The membership test runs on z_i = l_i / T - m against theta = log(cut * s): monotone in p_i, so
    // { z_i >= theta } is a prefix of the sorted order for ANY theta (up to the order among equal probabilities, which
    // the reference leaves to qsort), and the walk below verifies that the nucleus closes inside it. theta starts at
    // the safe cut (nothing below (1 - top_p) / (V - 1) can be in the nucleus; half of that here) and is raised until
    // at most kMaxCand entries pass (long flat tails). Only the candidates' probabilities are computed again.
    const float cut = V > 1 ? 0.5f * __fdiv_rn(__fsub_rn(1.0f, top_p), (float) (V - 1)) : 0.0f;
    float theta = cut > 0.0f ? logf(cut * s) : -INFINITY;
    theta = fmaxf(theta, -87.0f); // expf underflows to 0 below: such entries can never be chosen
    for (int round = 0; round < 64; ++round) {
        int mine = 0;
        for (int i = tid; i < V; i += kThreads) mine += __fsub_rn(__fdiv_rn(logits[i], temperature), m) >= theta ? 1 : 0;
        const int total = (int) block_reduce((float) mine, red, false); // counts < 2^24: exact in fp32
        if (total <= kMaxCand) break;
        theta += 1.0f; // e times fewer probability mass per step
    }
    for (int i = tid; i < V; i += kThreads) {
        const float z = __fsub_rn(__fdiv_rn(logits[i], temperature), m);
        if (z >= theta) {
            const float p = __fdiv_rn(expf(z), s);
            const int k = atomicAdd(&count, 1);
            if (k < kMaxCand && p > 0.0f) {
                sp[k] = p;
                si[k] = i;
            } else if (k < kMaxCand) { // keep the slot dense: a zero-probability entry sorts last and is never chosen
                sp[k] = 0.0f;
                si[k] = i;
            }
        }
    }
    __syncthreads();
    const int n = count;
    if (n > kMaxCand || n == 0) {
        if (tid == 0) {
            out[0] = -1;
            out[1] = 1; // not handled
        }
        return;
    }
    sort_and_walk(sp, si, n, top_p, coin, out, token_dev);
}

// Steps 4 of the header comment: sort the n candidates in shared memory (sp / si), then one thread walks them.
__device__ __forceinline__ void sort_and_walk(float* sp, int* si, int n, float top_p, float coin, int* __restrict__ out, int* token_dev) {
    const int tid = threadIdx.x;
    // 4. bitonic sort of the n candidates padded to a power of two
    int N = 2;
    while (N < n) N <<= 1;
    for (int i = n + tid; i < N; i += kThreads) {
        sp[i] = -1.0f; // sorts behind everything
        si[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < N; i += kThreads) {
                const int x = i ^ j;
                if (x > i) {
                    const float pa = sp[i], pb = sp[x];
                    const int ia = si[i], ib = si[x];
                    const bool up = (i & k) == 0;
                    if (up ? before(pb, ib, pa, ia) : before(pa, ia, pb, ib)) {
                        sp[i] = pb; si[i] = ib;
                        sp[x] = pa; si[x] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid != 0) return;
    // sampler_mass_index over the sorted head
    float mass = 0.0f;
    int id = -1;
    for (int i = 0; i < n; ++i) {
        mass = __fadd_rn(mass, sp[i]);
        if (mass > top_p) {
            id = i;
            break;
        }
    }
    if (id < 0) { // the nucleus reaches past the candidates (top_p ~ 1): let the host path decide
        out[0] = -1;
        out[1] = 1;
        return;
    }
    if (mass < 1e-3f)
        for (int i = 0; i <= id; ++i) mass = __fadd_rn(mass, sp[i]);
    // sampler_cdf_index(dist, n = id, coin, mass)
    float cdf = 0.0f;
    const float r = __fmul_rn(coin, mass);
    int tok = -2;
    for (int i = 0; i <= id; ++i) {
        cdf = __fadd_rn(cdf, sp[i]);
        if (r < cdf) {
            tok = si[i];
            break;
        }
    }
    if (tok == -2) tok = si[id > 0 ? id - 1 : 0];
    out[0] = tok;
    out[1] = 0;
    if (token_dev) *token_dev = tok;
}


// The same computation spread over the whole chip (cooperative launch, one CTA per SM): every pass touches one or two
// logits per thread and the passes meet at grid barriers. Opt-in (QWEN_SAMPLE_GRID=1): measured SLOWER than the
// single CTA at vocabulary 151 936 (see run_sample); kept as the cross-check of the single-CTA kernel in the tests. Workspace (ints / floats, zeroed by block 0 at entry):
//   [0 .. G) per-CTA maxima   [G .. 2G) per-CTA sums   [2G .. 2G+64) per-round counts   [2G+64] candidate counter
//   then kMaxCand probabilities and kMaxCand indices. Sums are added in CTA order: deterministic.
__global__ void __launch_bounds__(kThreads, 1)
k_sample_grid(const float* __restrict__ logits, int V, float temperature, float top_p, float coin, int* __restrict__ out, int* token_dev,
              float* __restrict__ ws) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float red[32];
    __shared__ float sp[kMaxCand];
    __shared__ int si[kMaxCand];
    const int tid = threadIdx.x, G = gridDim.x, b = blockIdx.x;
    const int stride = G * kThreads, i0 = b * kThreads + tid;
    float* pmax = ws;
    float* psum = ws + G;
    int* counts = reinterpret_cast<int*>(ws + 2 * G);
    int* ncand = counts + 64;
    float* cand_p = ws + 2 * G + 65;
    int* cand_i = reinterpret_cast<int*>(cand_p + kMaxCand);
    if (b == 0 && tid < 65) counts[tid] = 0;
    float m = -INFINITY;
    for (int i = i0; i < V; i += stride) m = fmaxf(m, logits[i]);
    m = block_reduce(m, red, true);
    if (tid == 0) pmax[b] = m;
    grid.sync();
    m = -INFINITY;
    for (int i = 0; i < G; ++i) m = fmaxf(m, __ldcg(pmax + i));
    m = __fdiv_rn(m, temperature);
    float s = 0.0f;
    for (int i = i0; i < V; i += stride) s = __fadd_rn(s, expf(__fsub_rn(__fdiv_rn(logits[i], temperature), m)));
    s = block_reduce(s, red, false);
    if (tid == 0) psum[b] = s;
    grid.sync();
    s = 0.0f;
    for (int i = 0; i < G; ++i) s = __fadd_rn(s, __ldcg(psum + i));
    const float cut = V > 1 ? 0.5f * __fdiv_rn(__fsub_rn(1.0f, top_p), (float) (V - 1)) : 0.0f;
    float theta = cut > 0.0f ? logf(cut * s) : -INFINITY;
    theta = fmaxf(theta, -87.0f);
    int n = 0;
    for (int round = 0; round < 64; ++round) {
        int mine = 0;
        for (int i = i0; i < V; i += stride) mine += __fsub_rn(__fdiv_rn(logits[i], temperature), m) >= theta ? 1 : 0;
        const int blk = (int) block_reduce((float) mine, red, false);
        if (tid == 0 && blk) atomicAdd(&counts[round], blk);
        grid.sync();
        n = *reinterpret_cast<volatile int*>(&counts[round]);
        if (n <= kMaxCand) break;
        theta += 1.0f;
    }
    if (n > kMaxCand || n == 0) { // uniform over the grid
        if (b == 0 && tid == 0) {
            out[0] = -1;
            out[1] = 1;
        }
        return;
    }
    for (int i = i0; i < V; i += stride) {
        const float z = __fsub_rn(__fdiv_rn(logits[i], temperature), m);
        if (z >= theta) {
            const int k = atomicAdd(ncand, 1);
            cand_p[k] = __fdiv_rn(expf(z), s);
            cand_i[k] = i;
        }
    }
    grid.sync();
    if (b != 0) return;
    for (int i = tid; i < n; i += kThreads) { // plain loads after the grid barrier: written by other CTAs
        sp[i] = __ldcg(cand_p + i);
        si[i] = __ldcg(cand_i + i);
    }
    __syncthreads();
    sort_and_walk(sp, si, n, top_p, coin, out, token_dev);
}

void clamp_like_sampler_create(float& temperature, float& top_p) { // reference src/sampler.c:34-52
    const float epsilon = 1e-6f;
    if (top_p > 1.0f || isnan(top_p) || (isinf(top_p) && top_p > 0)) top_p = 1.0f;
    else if (top_p < epsilon || (isinf(top_p) && top_p < 0)) top_p = epsilon;
    if (isnan(temperature) || (isinf(temperature) && temperature > 0)) temperature = 1.0f;
    else if (temperature < epsilon || (isinf(temperature) && temperature < 0)) temperature = epsilon;
}

constexpr size_t kWsFloats = 2 * 256 + 65 + 2 * kMaxCand;

int run_sample(const float* logits_dev, int V, float temperature, float top_p, float coin, int* out_dev, int* token_dev,
               float* ws, int num_sms, cudaStream_t st, int* token_out) {
    clamp_like_sampler_create(temperature, top_p);
    // QWEN_SAMPLE_GRID=1 selects the cooperative multi-CTA kernel. Measured on B200 (4B shape, bench.py sampled_generation):
    // single CTA 500 tok/s (the kernel costs ~0.12 ms), cooperative grid 447 tok/s (~0.36 ms: seven grid barriers of
    // 148 x 1024 threads and the cooperative launch cost more than the passes save) -- so one CTA is the default.
    const char* e = getenv("QWEN_SAMPLE_GRID");
    int G = std::min(std::min(num_sms, 256), (V + kThreads - 1) / kThreads);
    if (ws && G > 1 && e && atoi(e) == 1) {
        void* args[] = {(void*) &logits_dev, (void*) &V, (void*) &temperature, (void*) &top_p, (void*) &coin, (void*) &out_dev,
                        (void*) &token_dev, (void*) &ws};
        QW_CUDA(cudaLaunchCooperativeKernel((const void*) k_sample_grid, dim3(G), dim3(kThreads), args, 0, st));
    } else {
        k_sample<<<1, kThreads, 0, st>>>(logits_dev, V, temperature, top_p, coin, out_dev, token_dev);
    }
    QW_CUDA(cudaGetLastError());
    int res[2] = {-1, -1};
    QW_CUDA(cudaMemcpyAsync(res, out_dev, sizeof res, cudaMemcpyDeviceToHost, st));
    QW_CUDA(cudaStreamSynchronize(st));
    if (res[1] == 0) {
        *token_out = res[0];
        return 0;
    }
    *token_out = -1;
    return 1;
}

} // namespace

// Stands behind sample() (reference: src/sampler.c:186-201) for the logits of the context's last step.
extern "C" int qwen_cuda_sample(QwenCudaCtx* c, float temperature, float top_p, float coin, int* token_out) {
    if (!c || !token_out) return -2;
    QW_CUDA(cudaSetDevice(c->device));
    const float* src = c->tp_size > 1 ? c->logits_all : c->logits;
    if (!c->sample_ws && cudaMalloc((void**) &c->sample_ws, kWsFloats * 4) != cudaSuccess) {
        cudaGetLastError();
        c->sample_ws = nullptr; // the single-CTA kernel needs no workspace
    }
    // argmax_out doubles as the 2-int result slot (its first entries are rewritten by every greedy chain anyway)
    return run_sample(src, c->V, temperature, top_p, coin, c->argmax_out, c->token_dev, c->sample_ws, c->num_sms, c->stream, token_out);
}

// Context-free variant for tests: host logits in, token out.
extern "C" int qwen_cuda_sample_host(const float* logits_host, int vocab_size, float temperature, float top_p, float coin,
                                     int* token_out) {
    if (!logits_host || !token_out || vocab_size <= 0) return -2;
    if (qwen_cuda_device_count() <= 0) {
        qw_set_error("no CUDA device: this library has no CPU path");
        return -1;
    }
    float *d = nullptr, *ws = nullptr;
    int* o = nullptr;
    int rc = -1, sms = 1;
    do {
        if (cudaMalloc((void**) &d, (size_t) vocab_size * 4) || cudaMalloc((void**) &o, 16) || cudaMalloc((void**) &ws, kWsFloats * 4)) break;
        if (cudaMemcpy(d, logits_host, (size_t) vocab_size * 4, cudaMemcpyHostToDevice)) break;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        rc = run_sample(d, vocab_size, temperature, top_p, coin, o, nullptr, ws, sms, 0, token_out);
    } while (0);
    if (rc < 0) qw_set_error("sample_host: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
    cudaFree(o);
    cudaFree(ws);
    return rc;
}
