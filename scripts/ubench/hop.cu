// Micro-benchmark: one hand-off hop of the persistent decode kernel in isolation, by publish method. 148 CTAs x 480 threads;
// per hop every CTA publishes its ~17 floats of a D-float vector and then polls the whole vector (sentinel protocol), as
// prologue_quant does (10 warps x 1 record of 256 floats for D = 2560), optionally followed by the RMSNorm reduction and the
// quantiser-sized arithmetic. 64 hops chained, idle chip otherwise.
//   PUB 0: every warp stores its own rows (st.relaxed.gpu) + __threadfence()          [round-1 first version]
//   PUB 1: rows staged in shared memory, CTA barrier, ONE TMA bulk store by thread 0   [what the kernel does]
//   PUB 2: staged, CTA barrier, warp 0 stores the slice coalesced + one __threadfence()
//   PUB 3: staged, CTA barrier, warp 0 stores the slice with st.release.gpu (no fence)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop hop.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr uint32_t kSent = 0x7F808080u;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 load_u4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool unset(const uint4& v) { return v.x == kSent || v.y == kSent || v.z == kSent || v.w == kSent; }
template <int PUB, int WORK, int SKEW = 0, int POLL = 0>
__global__ void __launch_bounds__(480, 1) k_hop(uint32_t* arena, int D, int hops, unsigned long long* out, float* sink) {
    __shared__ __align__(16) float stage[64];
    __shared__ float red[16];
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = (int) ((long long) (D / 4) * b / G) * 4, r1 = (int) ((long long) (D / 4) * (b + 1) / G) * 4; // 16-byte aligned slices
    float acc = 0.f;
    const unsigned long long t0 = gtime();
    for (int h = 0; h < hops; ++h) {
        uint32_t* vec = arena + (size_t) h * D;
        if (SKEW) { // emulate load imbalance: every CTA is late by a pseudo-random 0 .. SKEW cycles
            const unsigned r = ((unsigned) b * 2654435761u + (unsigned) h * 40503u) >> 7;
            const long long until = clock64() + (long long) (r % (unsigned) SKEW);
            while (clock64() < until) {}
        }
        // produce
        if (PUB == 0 || PUB >= 4) {
            // PUB 4: no fence at all (every word validates itself) | 5: red.xor against the sentinel, no fence | 6: fence.cta only
            // PUB 7: as 0 plus a CTA barrier before the poll | 8: only the storing warps fence | 9: no fence, nanosleep(400) before the poll
            bool stored = false;
            for (int u = warp; 2 * u < r1 - r0; u += 15) {
                const int row = r0 + 2 * u + lane;
                stored = true;
                if (lane < 2 && row < r1) {
                    const uint32_t val = __float_as_uint(acc + (float) row);
                    if (PUB == 5) asm volatile("red.relaxed.gpu.global.xor.b32 [%0], %1;" ::"l"(vec + row), "r"(val ^ kSent) : "memory");
                    else asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(vec + row), "r"(val) : "memory");
                }
            }
            if (PUB == 0 || PUB == 7) __threadfence();
            if (PUB == 6) __threadfence_block();
            if (PUB == 8 && stored) __threadfence();
            if (PUB == 7) asm volatile("bar.sync 1, 480;" ::: "memory");
            if (PUB == 9) __nanosleep(400);
        } else {
            for (int u = warp; 2 * u < r1 - r0; u += 15) {
                const int row = r0 + 2 * u + lane;
                if (lane < 2 && row < r1) stage[row - r0] = acc + (float) row;
            }
            if (PUB == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 480;" ::: "memory");
            if (PUB == 1) {
                if (tid == 0 && r1 > r0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(vec + r0), "r"(s_u32(stage)), "r"((r1 - r0) * 4) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
            } else if (warp == 0) {
                for (int i = lane; i < r1 - r0; i += 32) {
                    if (PUB == 2) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(vec + r0 + i), "r"(__float_as_uint(stage[i])) : "memory");
                    else asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(vec + r0 + i), "r"(__float_as_uint(stage[i])) : "memory");
                }
                if (PUB == 2) __threadfence();
            }
        }
        // consume: poll the whole vector (lane owns 8 consecutive values of record warp), all pending pieces re-issued together
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (warp * 256 < D) {
            const uint32_t* q = vec + warp * 256 + lane * 8;
            uint4 a = load_u4(q), c = load_u4(q + 4);
            while (unset(a) || unset(c)) {
                if (POLL) __nanosleep(POLL);
                if (unset(a)) a = load_u4(q);
                if (unset(c)) c = load_u4(q + 4);
            }
            v[0] = __uint_as_float(a.x); v[1] = __uint_as_float(a.y); v[2] = __uint_as_float(a.z); v[3] = __uint_as_float(a.w);
            v[4] = __uint_as_float(c.x); v[5] = __uint_as_float(c.y); v[6] = __uint_as_float(c.z); v[7] = __uint_as_float(c.w);
        }
        if (WORK) { // RMSNorm reduction + a quantiser's worth of arithmetic
            float ss = 0.f;
            for (int i = 0; i < 8; ++i) ss += v[i] * v[i];
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (lane == 0) red[warp] = ss;
            asm volatile("bar.sync 1, 480;" ::: "memory");
            float tot = 0.f;
            for (int i = 0; i < 15; ++i) tot += red[i];
            const float r = 1.0f / sqrtf(tot / D + 1e-6f);
            float amax = 0.f;
            for (int i = 0; i < 8; ++i) { v[i] *= r; amax = fmaxf(amax, fabsf(v[i])); }
            for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            const float sc = amax / 127.0f, ri = 1.0f / (sc + 1e-30f);
            for (int i = 0; i < 8; ++i) acc += rintf(v[i] * ri) * 1e-9f;
        } else {
            acc += v[0] * 1e-9f + v[7] * 1e-9f;
        }
        asm volatile("bar.sync 1, 480;" ::: "memory");
    }
    if (tid == 0) out[b] = gtime() - t0;
    sink[b * 480 + tid] = acc;
}
__global__ void k_fill(uint32_t* p, size_t n, uint32_t v) {
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) p[i] = v;
}
template <int PUB, int WORK, int SKEW = 0, int POLL = 0>
void run(const char* name, uint32_t* arena, unsigned long long* out, float* sink, int D) {
    const int hops = 64, G = 148;
    for (int rep = 0; rep < 2; ++rep) {
        k_fill<<<148, 256>>>(arena, (size_t) hops * D, kSent);
        void* args[] = {&arena, (void*) &D, (void*) &hops, &out, &sink};
        cudaLaunchCooperativeKernel((const void*) k_hop<PUB, WORK, SKEW, POLL>, dim3(G), dim3(480), args, 0, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    }
    unsigned long long t[148], mx = 0;
    cudaMemcpy(t, out, sizeof t, cudaMemcpyDeviceToHost);
    for (int i = 0; i < G; ++i) mx = t[i] > mx ? t[i] : mx;
    printf("hop D=%5d %-58s %6.2f us per hop\n", D, name, (double) mx / hops / 1e3);
}
int main() {
    uint32_t* arena; unsigned long long* out; float* sink;
    cudaMalloc(&arena, 64 << 20); cudaMalloc(&out, 148 * 8); cudaMalloc(&sink, 148 * 480 * 4);
    const int D = 2560;
    run<0, 0>("per-warp st.relaxed + __threadfence, poll only", arena, out, sink, D);
    run<1, 0>("staged, barrier, ONE TMA bulk store, poll only", arena, out, sink, D);
    run<2, 0>("staged, barrier, warp 0 st.relaxed + fence, poll only", arena, out, sink, D);
    run<3, 0>("staged, barrier, warp 0 st.release, poll only", arena, out, sink, D);
    run<4, 0>("per-warp st.relaxed, NO fence, poll only", arena, out, sink, D);
    run<5, 0>("per-warp red.xor vs sentinel, no fence, poll only", arena, out, sink, D);
    run<6, 0>("per-warp st.relaxed + fence.cta, poll only", arena, out, sink, D);
    run<7, 0>("per-warp st.relaxed + __threadfence + CTA barrier, poll only", arena, out, sink, D);
    run<8, 0>("per-warp st.relaxed, only storing warps fence, poll only", arena, out, sink, D);
    run<9, 0>("per-warp st.relaxed, no fence, nanosleep 400 before poll", arena, out, sink, D);
    run<0, 1>("per-warp st.relaxed + __threadfence, + norm + quant work", arena, out, sink, D);
    run<4, 1>("per-warp st.relaxed, NO fence, + norm + quant work", arena, out, sink, D);
    run<5, 1>("per-warp red.xor, no fence, + norm + quant work", arena, out, sink, D);
    run<1, 1>("staged, barrier, ONE TMA bulk store, + norm + quant work", arena, out, sink, D);
    run<2, 1>("staged, barrier, warp 0 st.relaxed + fence, + norm + quant work", arena, out, sink, D);
    // load imbalance (every CTA late by 0 .. 3000 cycles = 0 .. 1.5 us per hop) and back-off between poll rounds
    run<0, 1, 3000, 0>("skew 3000: per-warp st + fence, tight poll", arena, out, sink, D);
    run<0, 1, 3000, 100>("skew 3000: per-warp st + fence, nanosleep 100", arena, out, sink, D);
    run<0, 1, 3000, 300>("skew 3000: per-warp st + fence, nanosleep 300", arena, out, sink, D);
    run<4, 1, 3000, 0>("skew 3000: per-warp st no fence, tight poll", arena, out, sink, D);
    run<4, 1, 3000, 100>("skew 3000: per-warp st no fence, nanosleep 100", arena, out, sink, D);
    run<4, 1, 3000, 200>("skew 3000: per-warp st no fence, nanosleep 200", arena, out, sink, D);
    run<4, 1, 3000, 400>("skew 3000: per-warp st no fence, nanosleep 400", arena, out, sink, D);
    run<4, 1, 3000, 800>("skew 3000: per-warp st no fence, nanosleep 800", arena, out, sink, D);
    run<1, 1, 3000, 0>("skew 3000: staged bulk store, tight poll", arena, out, sink, D);
    run<1, 1, 3000, 100>("skew 3000: staged bulk store, nanosleep 100", arena, out, sink, D);
    run<1, 1, 3000, 200>("skew 3000: staged bulk store, nanosleep 200", arena, out, sink, D);
    run<1, 1, 3000, 400>("skew 3000: staged bulk store, nanosleep 400", arena, out, sink, D);
    run<1, 1, 0, 200>("no skew: staged bulk store, nanosleep 200", arena, out, sink, D);
    run<1, 1, 0, 400>("no skew: staged bulk store, nanosleep 400", arena, out, sink, D);
    run<4, 1, 0, 200>("no skew: per-warp st no fence, nanosleep 200", arena, out, sink, D);
    run<4, 1, 0, 400>("no skew: per-warp st no fence, nanosleep 400", arena, out, sink, D);
    run<5, 1, 3000, 200>("skew 3000: per-warp red.xor no fence, nanosleep 200", arena, out, sink, D);
    return 0;
}
