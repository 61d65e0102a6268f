// Micro-benchmark: is a weight stream that every SM reads ONLY from lines homed on its own die's half of L2 as fast as
// L2-near data (17-18 TB/s), even when an L2 prefetch (which fills the home half only) brought it in?
//  1. probe: L2-hit latency of one line per 2 KB grain of a pool, from every SM -> near / far per (SM, grain)
//  2. SMs are split into two dies by agreement of their near/far labels with SM 0's on 512 common grains
//  3. die(grain) for the whole pool (each grain probed by two SMs of different dies: consistency check)
//  4. rate: every CTA prefetches (cp.async.bulk.prefetch.L2) a list of grains, waits, then bulk-copies them (2 KB copies, 15 warps
//     x 2 in flight): list = grains of its own die / of the other die / any grains
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dieprobe dieprobe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
constexpr int kGrain = 2048;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ld_cg(const void* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
// one warp per block, lane 0 measures: lat[b * n + i] = min over reps of the L2-hit latency of grain list[i] (cycles)
__global__ void k_probe(const uint8_t* pool, const unsigned* list, int n, int reps, unsigned short* lat, unsigned* smid, unsigned* sink) {
    if (threadIdx.x != 0) return;
    unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    smid[blockIdx.x] = s;
    unsigned acc = 0, off = 0;
    __shared__ unsigned scratch[32];
    const uint32_t sa = s_u32(&scratch[0]);
    for (int i = 0; i < n; ++i) {
        const uint8_t* a = pool + (size_t) list[i] * kGrain;
        for (int j = 0; j < 4; ++j) acc += ld_cg(a + 128 * j); // bring four lines of the grain into L2
        unsigned best = 0xffffu;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned v = ld_cg(a + 128 * j + off);
                off = v - 0x01010101u; // 0 at run time (the pool is filled with 1s): the next load depends on this one
            }
            if (off != 0) break;       // a branch on the last value: the clock below is read after it has arrived
            const long long t1 = clock64();
            best = min(best, (unsigned) (t1 - t0) / 4u);
        }
        lat[(size_t) blockIdx.x * n + i] = (unsigned short) best;
    }
    if (acc == 0x12345u) sink[0] = acc;
}
// whole-pool map: block b probes grains b, b + G, ...; out[g] = latency from this block's SM
__global__ void k_map(const uint8_t* pool, size_t ngrains, int reps, unsigned short* out, int shift, unsigned* sink) {
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) != 0) return;
    unsigned acc = 0, off = 0;
    __shared__ unsigned scratch[32];
    const uint32_t sa = s_u32(&scratch[warp]);
    const size_t G = gridDim.x;
    for (size_t g = (blockIdx.x + shift) % G + (size_t) warp * G; g < ngrains; g += G * nw) {
        const uint8_t* a = pool + g * kGrain;
        for (int j = 0; j < 4; ++j) acc += ld_cg(a + 128 * j); // bring four lines of the grain into L2
        unsigned best = 0xffffu;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned v = ld_cg(a + 128 * j + off);
                off = v - 0x01010101u; // 0 at run time (the pool is filled with 1s): the next load depends on this one
            }
            if (off != 0) break;       // a branch on the last value: the clock below is read after it has arrived
            const long long t1 = clock64();
            best = min(best, (unsigned) (t1 - t0) / 4u);
        }
        out[g] = (unsigned short) best;
    }
    if (acc == 0x12345u) sink[0] = acc;
}
// one-shot probe: the window [g0, g1) has been brought to its HOME half by an L2 prefetch and not been read since; block b times
// the FIRST load of grains g with (g + shift) % G == b (a second load would find the copy the first one left in the near half)
__global__ void k_pfw(const uint8_t* pool, size_t g0, size_t g1) {
    const size_t g = g0 + (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (g < g1) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pool + g * kGrain), "r"(kGrain) : "memory");
}
__global__ void k_map1(const uint8_t* pool, size_t g0, size_t g1, unsigned short* out, int shift, unsigned* sink) {
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) != 0) return;
    unsigned off = 0;
    const size_t G = gridDim.x;
    size_t g = g0 + ((blockIdx.x + G - (g0 + shift) % G) % G) + (size_t) warp * G; // first grain >= g0 with (g + shift) % G == b, then warps interleave
    for (; g < g1; g += G * nw) {
        const uint8_t* a = pool + g * kGrain + 1024; // a line in the middle of the grain
        const long long t0 = clock64();
        const unsigned v = ld_cg(a + off);
        off = v - 0x01010101u;
        if (off != 0) break;
        const long long t1 = clock64();
        out[g] = (unsigned short) min((unsigned) (t1 - t0), 0xffffu);
    }
    if (off == 0x12345u) sink[0] = off;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
// every CTA: list[b * per + i] = grain index; prefetch all, wait, copy all (15 warps x 2 stages of 2 KB)
__global__ void __launch_bounds__(512, 1) k_rate(const uint8_t* pool, const unsigned* list, int per, long long wait_ns, unsigned long long* out, int cb = kGrain, int unit = kGrain) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned* my = list + (size_t) blockIdx.x * per;
    if (wait_ns >= 0) {
        const unsigned long long t = gtime();
        for (int i = threadIdx.x; i < per; i += blockDim.x)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pool + (size_t) my[i] * unit), "r"(cb) : "memory");
        while ((long long) (gtime() - t) < wait_ns) {}
        __syncthreads();
    }
    if (lane != 0 || warp >= 15) return;
    uint8_t* region = smem + (size_t) warp * 2 * cb;
    const unsigned long long t0 = gtime();
    int issued = 0, done = 0;
    const int total = (per - warp + 14) / 15;
    auto issue = [&](int i) {
        const uint32_t bar = s_u32(&bars[warp * 2 + (i & 1)]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(region + (i & 1) * cb)),
                     "l"(pool + (size_t) my[warp + i * 15] * unit), "r"(cb), "r"(bar) : "memory");
    };
    for (; issued < total && issued < 2; ++issued) issue(issued);
    for (; done < total; ++done) {
        wait(s_u32(&bars[warp * 2 + (done & 1)]), (unsigned) ((done >> 1) & 1));
        if (issued < total) { issue(issued); ++issued; }
    }
    out[blockIdx.x * 16 + warp] = gtime() - t0;
}
// contiguous regions: CTA b copies region (b + shift) % G (per bytes, 15 warps x 2 x cb in flight), one pass
__global__ void __launch_bounds__(512, 1) k_rate2(const uint8_t* pool, size_t per, int cb, int shift, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bars[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bars[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (lane != 0 || warp >= 15) return;
    const uint8_t* src = pool + per * ((blockIdx.x + shift) % gridDim.x);
    uint8_t* region = smem + (size_t) warp * 2 * cb;
    const unsigned long long t0 = gtime();
    int issued = 0, done = 0;
    const int ncopy = (int) (per / cb), total = (ncopy - warp + 14) / 15;
    auto issue = [&](int i) {
        const uint32_t bar = s_u32(&bars[warp * 2 + (i & 1)]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(region + (i & 1) * cb)),
                     "l"(src + (size_t) (warp + i * 15) * cb), "r"(cb), "r"(bar) : "memory");
    };
    for (; issued < total && issued < 2; ++issued) issue(issued);
    for (; done < total; ++done) {
        wait(s_u32(&bars[warp * 2 + (done & 1)]), (unsigned) ((done >> 1) & 1));
        if (issued < total) { issue(issued); ++issued; }
    }
    out[blockIdx.x * 16 + warp] = gtime() - t0;
}
int main() {
    const int G = 148;
    const size_t pool_bytes = (size_t) 2 << 30, ngrains = pool_bytes / kGrain;
    uint8_t* pool; cudaMalloc(&pool, pool_bytes); cudaMemset(pool, 1, pool_bytes);
    unsigned* sink; cudaMalloc(&sink, 64);
    // 1 + 2: common grains, every SM
    const int nc = 512, reps = 4;
    std::vector<unsigned> common(nc);
    for (int i = 0; i < nc; ++i) common[i] = (unsigned) ((size_t) rand() * 7919 % ngrains);
    unsigned* d_list; cudaMalloc(&d_list, nc * 4); cudaMemcpy(d_list, common.data(), nc * 4, cudaMemcpyHostToDevice);
    unsigned short* d_lat; cudaMalloc(&d_lat, (size_t) G * nc * 2);
    unsigned* d_smid; cudaMalloc(&d_smid, G * 4);
    for (int rep = 0; rep < 2; ++rep) k_probe<<<G, 32>>>(pool, d_list, nc, reps, d_lat, d_smid, sink);
    if (cudaDeviceSynchronize()) { printf("probe failed\n"); return 1; }
    std::vector<unsigned short> lat((size_t) G * nc); std::vector<unsigned> smid(G);
    cudaMemcpy(lat.data(), d_lat, lat.size() * 2, cudaMemcpyDeviceToHost); cudaMemcpy(smid.data(), d_smid, G * 4, cudaMemcpyDeviceToHost);
    std::vector<int> thr(G), die(G);
    for (int b = 0; b < G; ++b) {
        std::vector<unsigned short> v(lat.begin() + (size_t) b * nc, lat.begin() + (size_t) (b + 1) * nc);
        std::sort(v.begin(), v.end());
        thr[b] = (v[nc / 10] + v[nc * 9 / 10]) / 2;
        if (b < 4) printf("block %d smid %u: latency p10 %u p50 %u p90 %u, threshold %d\n", b, smid[b], v[nc / 10], v[nc / 2], v[nc * 9 / 10], thr[b]);
    }
    int n0 = 0;
    for (int b = 0; b < G; ++b) {
        int agree = 0;
        for (int i = 0; i < nc; ++i) agree += (lat[(size_t) b * nc + i] > thr[b]) == (lat[i] > thr[0]);
        die[b] = agree * 2 > nc ? 0 : 1;
        n0 += die[b] == 0;
        if (b < 8 || agree * 10 > nc * 2 && agree * 10 < nc * 8) printf("block %d: agreement with block 0 %d / %d\n", b, agree, nc);
    }
    printf("dies: %d / %d blocks\n", n0, G - n0);
    // 3: the whole pool, twice with shifted block -> grain assignment
    unsigned short *d_m0, *d_m1; cudaMalloc(&d_m0, ngrains * 2); cudaMalloc(&d_m1, ngrains * 2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    uint8_t* flush0; cudaMalloc(&flush0, 512 << 20);
    cudaEventRecord(e0);
    for (int pass = 0; pass < 2; ++pass) {
        cudaMemset(flush0, pass, 512 << 20);
        const size_t win = (size_t) 24 << 20 >> 11; // grains per window: 24 MB
        for (size_t g0 = 0; g0 < ngrains; g0 += win) {
            const size_t g1 = std::min(ngrains, g0 + win);
            k_pfw<<<(unsigned) ((g1 - g0 + 255) / 256), 256>>>(pool, g0, g1);
            k_map1<<<G, 128>>>(pool, g0, g1, pass ? d_m1 : d_m0, pass, sink); // (the launch gap is the wait)
        }
        if (pass == 0) cudaEventRecord(e1);
    }
    if (cudaDeviceSynchronize()) { printf("map failed\n"); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned short> m0(ngrains), m1(ngrains);
    cudaMemcpy(m0.data(), d_m0, ngrains * 2, cudaMemcpyDeviceToHost); cudaMemcpy(m1.data(), d_m1, ngrains * 2, cudaMemcpyDeviceToHost);
    std::vector<int> thr1(G);
    for (int pass = 0; pass < 2; ++pass) { // the one-shot latencies have their own scale: threshold per block from its own sample
        for (int b = 0; b < G; ++b) {
            std::vector<unsigned short> v;
            for (size_t g = (size_t) ((b + G - pass) % G); g < ngrains && v.size() < 4000; g += G) v.push_back((pass ? m1 : m0)[g]);
            std::sort(v.begin(), v.end());
            const int t = (v[v.size() / 10] + v[v.size() * 9 / 10]) / 2;
            if (pass == 0) thr[b] = t; else thr1[b] = t;
            if (b < 3) printf("one-shot pass %d block %d: p05 %u p25 %u p50 %u p75 %u p95 %u threshold %d\n", pass, b, v[v.size() / 20], v[v.size() / 4], v[v.size() / 2], v[v.size() * 3 / 4], v[v.size() * 19 / 20], t);
        }
    }
    { // dies of the blocks from the one-shot passes: grain g was timed by block g % G (pass 0) and block (g + 1) % G (pass 1); the two
      // see it the same way (both near or both far) iff they sit on the same die
        std::vector<long long> same(G, 0), tot(G, 0);
        for (size_t g = 0; g < ngrains; ++g) {
            const int b0 = (int) (g % G), b1 = (int) ((g + 1) % G);
            same[b0] += (m0[g] > thr[b0]) == (m1[g] > thr1[b1]);
            ++tot[b0];
        }
        die[0] = 0;
        int n0b = 1, unsure = 0;
        for (int b = 0; b + 1 < G; ++b) {
            const double f = (double) same[b] / tot[b];
            die[b + 1] = f > 0.5 ? die[b] : 1 - die[b];
            n0b += die[b + 1] == 0;
            unsure += f > 0.2 && f < 0.8;
        }
        printf("dies from the one-shot passes: %d / %d blocks (%d adjacent pairs unsure); block 147 vs block 0: same-label share %.3f (expect %s)\n", n0b, G - n0b, unsure,
               (double) same[G - 1] / tot[G - 1], die[G - 1] == die[0] ? "high" : "low");
    }
    std::vector<unsigned char> gdie(ngrains);
    size_t cmp = 0, bad = 0, ones = 0;
    for (size_t g = 0; g < ngrains; ++g) {
        const int b0 = (int) ((g % G + G - 0) % G), b1 = (int) ((g + 1) % G); // block that probed g in pass 0 / 1
        const int d0 = (m0[g] > thr[b0]) ? 1 - die[b0] : die[b0], d1 = (m1[g] > thr1[b1]) ? 1 - die[b1] : die[b1];
        gdie[g] = (unsigned char) d0;
        ones += d0;
        ++cmp; bad += d0 != d1;
    }
    printf("map of %zu grains in %.2f ms per pass (8 warps per SM probing); die 1 share %.3f; the two passes disagree on %.3f %%\n", ngrains, ms, (double) ones / ngrains,
           100.0 * bad / cmp);
    size_t runs[8] = {0};
    for (size_t g = 0, len = 1; g + 1 <= ngrains; ++g) {
        if (g + 1 < ngrains && gdie[g + 1] == gdie[g]) { ++len; continue; }
        runs[std::min<size_t>(len, 7)] += len; len = 1;
    }
    printf("share of grains in same-die runs of length 1..6, >= 7:");
    for (int i = 1; i < 8; ++i) printf(" %.3f", (double) runs[i] / ngrains);
    printf("\n");
    // 4: rate
    const int per = 200; // 400 KB per SM
    std::vector<unsigned> own((size_t) G * per), other((size_t) G * per), any((size_t) G * per);
    size_t cur[3] = {0, ngrains / 3, 2 * ngrains / 3};
    for (int b = 0; b < G; ++b)
        for (int i = 0; i < per; ++i) {
            while (gdie[cur[0]] != die[b]) ++cur[0];
            own[(size_t) b * per + i] = (unsigned) cur[0]++;
            while (gdie[cur[1]] == die[b]) ++cur[1];
            other[(size_t) b * per + i] = (unsigned) cur[1]++;
            any[(size_t) b * per + i] = (unsigned) cur[2]++;
        }
    unsigned* d_l; cudaMalloc(&d_l, (size_t) G * per * 4);
    unsigned long long* d_out; cudaMalloc(&d_out, G * 16 * 8);
    uint8_t* flush; cudaMalloc(&flush, 512 << 20);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[3] = {"grains of the SM's own die", "grains of the other die", "any grains"};
    std::vector<unsigned>* lists[3] = {&own, &other, &any};
    for (int rep = 0; rep < 2; ++rep)
        for (int w = 0; w < 3; ++w)
            for (long long T : {-1LL, 10000LL}) {
                cudaMemset(flush, rep, 512 << 20);
                cudaMemcpy(d_l, lists[w]->data(), (size_t) G * per * 4, cudaMemcpyHostToDevice);
                cudaMemset(d_out, 0, G * 16 * 8);
                k_rate<<<G, 512, 15 * 2 * kGrain>>>(pool, d_l, per, T, d_out);
                if (cudaDeviceSynchronize()) { printf("rate failed\n"); return 1; }
                unsigned long long t[148 * 16], mx = 0;
                cudaMemcpy(t, d_out, sizeof t, cudaMemcpyDeviceToHost);
                for (int i = 0; i < G * 16; ++i) mx = std::max(mx, t[i]);
                printf("%-28s %s: %6.2f us  %6.2f TB/s\n", names[w], T < 0 ? "cold (HBM)          " : "L2 prefetch, 10 us, copy", mx / 1e3, (double) per * kGrain * G / mx / 1e3);
            }
    // 4b: 5440-byte items (a row pair of a 2560-column matrix) dealt to the die that holds most of their bytes
    {
        const int cb = 5440, per = 75;
        cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        std::vector<unsigned> maj((size_t) G * per), mino((size_t) G * per), anyi((size_t) G * per);
        auto item_die_bytes = [&](size_t item, int d) { // bytes of item (at item * cb) that live on die d
            size_t a = item * cb, e = a + cb, n = 0;
            for (size_t g = a / kGrain; g * kGrain < e; ++g) {
                const size_t lo = std::max(a, g * kGrain), hi = std::min(e, (g + 1) * kGrain);
                if (gdie[g] == d) n += hi - lo;
            }
            return n;
        };
        size_t c0 = 0, c1 = ngrains * kGrain / cb / 3, c2 = 2 * (ngrains * kGrain / cb / 3);
        double near_share = 0;
        for (int b = 0; b < G; ++b)
            for (int i = 0; i < per; ++i) {
                while (item_die_bytes(c0, die[b]) * 2 < cb) ++c0;
                near_share += (double) item_die_bytes(c0, die[b]) / cb;
                maj[(size_t) b * per + i] = (unsigned) (c0++ * (cb / 16));
                while (item_die_bytes(c1, die[b]) * 2 >= cb) ++c1;
                mino[(size_t) b * per + i] = (unsigned) (c1++ * (cb / 16));
                anyi[(size_t) b * per + i] = (unsigned) (c2++ * (cb / 16));
            }
        printf("5440-byte items by majority die: %.1f %% of their bytes classified near\n", 100.0 * near_share / (G * per));
        const char* nm[3] = {"items mostly on the SM's own die", "items mostly on the other die", "any items"};
        std::vector<unsigned>* ls[3] = {&maj, &mino, &anyi};
        for (int rep = 0; rep < 2; ++rep)
            for (int w = 0; w < 3; ++w)
                for (long long T : {-1LL, 10000LL}) {
                    cudaMemset(flush, rep, 512 << 20);
                    cudaMemcpy(d_l, ls[w]->data(), (size_t) G * per * 4, cudaMemcpyHostToDevice);
                    cudaMemset(d_out, 0, G * 16 * 8);
                    k_rate<<<G, 512, 15 * 2 * cb>>>(pool, d_l, per, T, d_out, cb, 16);
                    if (cudaDeviceSynchronize()) { printf("rate failed\n"); return 1; }
                    unsigned long long t[148 * 16], mx = 0;
                    cudaMemcpy(t, d_out, sizeof t, cudaMemcpyDeviceToHost);
                    for (int i = 0; i < G * 16; ++i) mx = std::max(mx, t[i]);
                    printf("%-34s %s: %6.2f us  %6.2f TB/s\n", nm[w], T < 0 ? "cold (HBM)          " : "L2 prefetch, 10 us, copy", mx / 1e3, (double) per * cb * G / mx / 1e3);
                }
    }
    // 5: who has to have read the data before for the copies to be fast? warm pass by CTA b + shift, timed pass by CTA b
    cudaFuncSetAttribute(k_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cb = 5440;
    const size_t per2 = (size_t) 400 * 1024 / cb * cb;
    size_t region = 0;
    for (int shift : {0, 74}) {
        int same = 0;
        for (int b = 0; b < G; ++b) same += die[b] == die[(b + G - shift) % G];
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(flush, rep, 512 << 20);
            const uint8_t* base = pool + region;
            region = (region + per2 * G + 4095) / 4096 * 4096;
            if (region + per2 * G > pool_bytes) region = 0;
            // warm: CTA b reads region b + shift  <=>  region r is read by CTA r - shift
            k_rate2<<<G, 512, 15 * 2 * cb>>>(base, per2, cb, shift, d_out);
            cudaMemset(d_out, 0, G * 16 * 8);
            k_rate2<<<G, 512, 15 * 2 * cb>>>(base, per2, cb, 0, d_out);
            if (cudaDeviceSynchronize()) { printf("rate2 failed\n"); return 1; }
            unsigned long long t[148 * 16], mx = 0, tb[148] = {0};
            cudaMemcpy(t, d_out, sizeof t, cudaMemcpyDeviceToHost);
            for (int i = 0; i < G * 16; ++i) { mx = std::max(mx, t[i]); tb[i / 16] = std::max(tb[i / 16], t[i]); }
            double ts = 0, td = 0; int ns = 0, nd = 0;
            for (int b = 0; b < G; ++b) { if (die[b] == die[(b + G - shift) % G]) { ts += tb[b]; ++ns; } else { td += tb[b]; ++nd; } }
            printf("warm pass by CTA b - %2d (%3d of 148 on the same die): %6.2f us %6.2f TB/s; mean CTA time same-die warm %.2f us (%d), other-die warm %.2f us (%d)\n", shift, same,
                   mx / 1e3, (double) per2 * G / mx / 1e3, ns ? ts / ns / 1e3 : 0.0, ns, nd ? td / nd / 1e3 : 0.0, nd);
        }
    }
    return 0;
}
