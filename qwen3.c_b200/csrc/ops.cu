// ops.cu -- one sm_100a kernel per forward.h / q8.h op, plus their host-in/host-out
// C-ABI wrappers (include/qwen_cuda.h). These are the op-level parity boundary and the
// building blocks of the per-op decode path (decode_ops.cu). The fast path is the
// persistent kernel in decode_mega.cu.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

// ---------------------------------------------------------------------------
// error text
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void qw_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char* qwen_cuda_last_error(void) { return g_err; }
extern "C" void qwen_cuda_clear_error(void) { g_err[0] = 0; }

extern "C" int qwen_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ---------------------------------------------------------------------------
// K3  q8_quantize   (reference: src/q8.c:5-30)
// One warp per group of 64; lane owns elements lane and lane+32. absmax by
// shuffle (max is order-independent, so the scale is bit-identical), IEEE
// division, roundf half-away-from-zero, clamp.
// ---------------------------------------------------------------------------
__global__ void k_quantize(const float* __restrict__ x, int8_t* __restrict__ q, float* __restrict__ s,
                           int groups) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= groups) return;
    const float a = x[(size_t) warp * 64 + lane];
    const float b = x[(size_t) warp * 64 + 32 + lane];
    const float amax = warp_max(fmaxf(fabsf(a), fabsf(b)));
    const float scale = q8_scale(amax);
    q[(size_t) warp * 64 + lane] = (int8_t) q8_code(a, scale);
    q[(size_t) warp * 64 + 32 + lane] = (int8_t) q8_code(b, scale);
    if (lane == 0) s[warp] = scale;
}

void launch_quantize(const float* x, int8_t* q, float* s, int n, cudaStream_t st) {
    const int groups = n / 64;
    if (groups <= 0) return;
    const int threads = 256;
    const int blocks = (groups * 32 + threads - 1) / threads;
    k_quantize<<<blocks, threads, 0, st>>>(x, q, s, groups);
}

// generic-group variant (any group size, one thread block per group) for the op ABI
__global__ void k_quantize_generic(const float* __restrict__ x, int8_t* __restrict__ q, float* __restrict__ s,
                                   int group) {
    __shared__ float red[32];
    const float* xg = x + (size_t) blockIdx.x * group;
    float amax = 0.0f;
    for (int i = threadIdx.x; i < group; i += blockDim.x) amax = fmaxf(amax, fabsf(xg[i]));
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
        v = warp_max(v);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const float scale = q8_scale(red[0]);
    if (threadIdx.x == 0) s[blockIdx.x] = scale;
    for (int i = threadIdx.x; i < group; i += blockDim.x)
        q[(size_t) blockIdx.x * group + i] = (int8_t) q8_code(xg[i], scale);
}

// K1/a3  q8_dequantize  (reference: src/q8.c:32-37)
__global__ void k_dequantize(const int8_t* __restrict__ q, const float* __restrict__ s, float* __restrict__ x,
                             int n, int group) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __fmul_rn((float) q[i], s[i / group]);
}

// ---------------------------------------------------------------------------
// K2  rmsnorm  (reference: src/forward.c:12-28). One block; fixed-order tree sum.
// ---------------------------------------------------------------------------
// out may alias x (the final norm runs in place, forward.c:344), so no __restrict__ on them.
__global__ void k_rmsnorm(float* out, const float* x, const float* __restrict__ w, int size) {
    __shared__ float red[32];
    float ss = 0.0f;
    for (int i = threadIdx.x; i < size; i += blockDim.x) ss = __fadd_rn(ss, __fmul_rn(x[i], x[i]));
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const float r = rms_rscale(red[0], size);
    for (int i = threadIdx.x; i < size; i += blockDim.x) out[i] = __fmul_rn(w[i], __fmul_rn(r, x[i]));
}

void launch_rmsnorm(float* out, const float* x, const float* w, int size, cudaStream_t st) {
    const int threads = size >= 1024 ? 1024 : (size >= 256 ? 256 : 128);
    k_rmsnorm<<<1, threads, 0, st>>>(out, x, w, size);
}

// ---------------------------------------------------------------------------
// a8  softmax  (reference: src/forward.c:34-77). One block, in place.
// ---------------------------------------------------------------------------
__global__ void k_softmax(float* __restrict__ x, int size) {
    __shared__ float red[32];
    __shared__ float bcast;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < size; i += blockDim.x) m = fmaxf(m, x[i]);
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
        v = warp_max(v);
        if (threadIdx.x == 0) bcast = v;
    }
    __syncthreads();
    m = bcast;
    float sum = 0.0f;
    for (int i = threadIdx.x; i < size; i += blockDim.x) {
        const float e = expf(__fsub_rn(x[i], m));
        x[i] = e;
        sum = __fadd_rn(sum, e);
    }
    sum = warp_sum(sum);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) bcast = v;
    }
    __syncthreads();
    sum = bcast;
    for (int i = threadIdx.x; i < size; i += blockDim.x) x[i] = __fdiv_rn(x[i], sum);
}

// ---------------------------------------------------------------------------
// a6  rotary  (reference: src/forward.c:104-118). cos/sin are host libm values.
// ---------------------------------------------------------------------------
__global__ void k_rotary(float* __restrict__ x, const float* __restrict__ c, const float* __restrict__ s, int half) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const float a = x[i], b = x[i + half];
    x[i] = __fsub_rn(__fmul_rn(a, c[i]), __fmul_rn(b, s[i]));
    x[i + half] = __fadd_rn(__fmul_rn(a, s[i]), __fmul_rn(b, c[i]));
}

// a9  swiglu / silu / sigmoid  (reference: src/forward.c:122-139)
__global__ void k_swiglu(float* __restrict__ x1, const float* __restrict__ x3, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x1[i] = __fmul_rn(silu_ref(x1[i]), x3[i]);
}
__global__ void k_silu(float* __restrict__ y, const float* __restrict__ x, int n, int sigmoid_only) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = sigmoid_only ? sigmoid_ref(x[i]) : silu_ref(x[i]);
}

// ---------------------------------------------------------------------------
// repack: checkpoint layout (int8[rows*src_n] + fp32[rows*src_n/64]) -> SG layout.
// Takes a column window [col0, col0+n) of each source row (tensor-parallel row
// splits) and scatters destination rows as dst_row0 + r*dst_row_step (QKV
// concatenation, w1/w3 interleave). One thread per 16 destination bytes.
// ---------------------------------------------------------------------------
__global__ void k_repack(const int8_t* __restrict__ src_q, const float* __restrict__ src_s, int src_n, int col0,
                         int n, int rows, uint8_t* __restrict__ dst, int dst_row0, int dst_row_step) {
    const int sgpr = qw_sg_per_row(n);
    const long long pieces_per_row = (long long) sgpr * 17;
    const long long total = pieces_per_row * rows;
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long) gridDim.x * blockDim.x) {
        const int r = (int) (t / pieces_per_row);
        const int rem = (int) (t % pieces_per_row);
        const int sg = rem / 17, piece = rem % 17;
        uint8_t* drow = dst + (size_t) (dst_row0 + (size_t) r * dst_row_step) * ((size_t) sgpr * QW_SG_BYTES);
        uint8_t* d = drow + (size_t) sg * QW_SG_BYTES + piece * 16;
        const size_t srow = (size_t) r * src_n + col0;
        if (piece < 16) {
            const int c = sg * 256 + piece * 16;
            int4 v = make_int4(0, 0, 0, 0);
            if (c + 16 <= n) {
                // source rows are only guaranteed 64-byte aligned relative to the tensor start
                const int8_t* sp = src_q + srow + c;
                if ((((uintptr_t) sp) & 15) == 0) {
                    v = *reinterpret_cast<const int4*>(sp);
                } else {
                    int8_t tmp[16];
                    for (int k = 0; k < 16; ++k) tmp[k] = sp[k];
                    v = *reinterpret_cast<int4*>(tmp);
                }
            } else if (c < n) {
                int8_t tmp[16];
                for (int k = 0; k < 16; ++k) tmp[k] = (c + k < n) ? src_q[srow + c + k] : (int8_t) 0;
                v = *reinterpret_cast<int4*>(tmp);
            }
            *reinterpret_cast<int4*>(d) = v;
        } else {
            float sc[4];
            for (int g = 0; g < 4; ++g) {
                const int c = sg * 256 + g * 64;
                sc[g] = (c < n) ? src_s[(srow + c) / 64] : 0.0f;
            }
            *reinterpret_cast<float4*>(d) = make_float4(sc[0], sc[1], sc[2], sc[3]);
        }
    }
}

void launch_repack(const int8_t* src_q, const float* src_s, int src_n, int col0, int n, int rows, uint8_t* dst,
                   int dst_row0, int dst_row_step, cudaStream_t st) {
    const long long total = (long long) qw_sg_per_row(n) * 17 * rows;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    k_repack<<<(int) blocks, 256, 0, st>>>(src_q, src_s, src_n, col0, n, rows, dst, dst_row0, dst_row_step);
}

// ---------------------------------------------------------------------------
// K4  Q8_0 GEMV over the SG layout  (reference: src/forward.c:79-101)
// One warp per row. A half-warp takes one 272-byte record per step: each lane
// loads 16 codes of W and the matching 16 codes of x, 4 x dp4a, and the four lanes
// of a group add their int32 partials by shuffle -> the group's exact int32 dot.
// The group leader applies ((float) dot * ws) * xs and folds it into its fp32
// accumulator; the 8 leaders are combined by a fixed shuffle tree.
// x codes and scales are staged in shared memory once per block.
// ---------------------------------------------------------------------------
#define GEMV_THREADS 256

__global__ void __launch_bounds__(GEMV_THREADS)
k_gemv_sg(const uint8_t* __restrict__ w, const int8_t* __restrict__ xq, const float* __restrict__ xs,
          float* __restrict__ out, int rows, int n, int32_t* __restrict__ dots) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int sgpr = qw_sg_per_row(n);
    const int ncols = sgpr * 256;
    int8_t* sxq = reinterpret_cast<int8_t*>(smem);
    float* sxs = reinterpret_cast<float*>(smem + ncols);
    for (int i = threadIdx.x; i < ncols / 16; i += blockDim.x)
        reinterpret_cast<int4*>(sxq)[i] = reinterpret_cast<const int4*>(xq)[i];
    for (int i = threadIdx.x; i < sgpr * 4; i += blockDim.x) sxs[i] = xs[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int l16 = lane & 15, half = lane >> 4;
    const int grp = l16 >> 2;
    // the two half-warps run different trip counts when sgpr is odd: shuffle within the half only
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int warps_per_block = blockDim.x >> 5;
    const int n_groups = n / 64;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows;
         row += gridDim.x * warps_per_block) {
        const uint8_t* wr = w + (size_t) row * sgpr * QW_SG_BYTES;
        float acc = 0.0f;
        for (int sg = half; sg < sgpr; sg += 2) {
            const uint8_t* rec = wr + (size_t) sg * QW_SG_BYTES;
            const int4 wv = __ldg(reinterpret_cast<const int4*>(rec) + l16);
            const int4 xv = reinterpret_cast<const int4*>(sxq + sg * 256)[l16];
            int dot = dot16(wv, xv);
            dot += __shfl_xor_sync(hmask, dot, 1);
            dot += __shfl_xor_sync(hmask, dot, 2);
            if ((l16 & 3) == 0) {
                const float wsc = __ldg(reinterpret_cast<const float*>(rec + 256) + grp);
                acc = __fadd_rn(acc, q8_term(dot, wsc, sxs[sg * 4 + grp]));
                if (dots && sg * 4 + grp < n_groups) dots[(size_t) row * n_groups + sg * 4 + grp] = dot;
            }
        }
        __syncwarp();
        // leaders are lanes 0,4,...,28; everyone else holds 0
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 4));
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));
        acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));
        if (lane == 0) out[row] = acc;
    }
}

void launch_gemv_sg(const uint8_t* w, const int8_t* xq, const float* xs, float* out, int rows, int n,
                    int32_t* dots, cudaStream_t st) {
    const int sgpr = qw_sg_per_row(n);
    const size_t smem = (size_t) sgpr * 256 + (size_t) sgpr * 16;
    const int wpb = GEMV_THREADS / 32;
    int blocks = (rows + wpb - 1) / wpb;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (smem > 48 * 1024) {
        static bool once = false;
        if (!once) {
            cudaFuncSetAttribute(k_gemv_sg, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            once = true;
        }
    }
    k_gemv_sg<<<blocks, GEMV_THREADS, smem, st>>>(w, xq, xs, out, rows, n, dots);
}

// argmax with lowest-index tie break (greedy chain); writes *out and, if given, *also
__global__ void k_argmax(const float* __restrict__ v, int n, int* __restrict__ out, int* __restrict__ also) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float f = v[i];
        if (f > best || (f == best && i < idx)) {
            best = f;
            idx = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) {
            best = ob;
            idx = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        bv[threadIdx.x >> 5] = best;
        bi[threadIdx.x >> 5] = idx;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? bv[threadIdx.x] : -INFINITY;
        idx = threadIdx.x < (blockDim.x >> 5) ? bi[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ob > best || (ob == best && oi < idx)) {
                best = ob;
                idx = oi;
            }
        }
        if (threadIdx.x == 0) {
            *out = idx;
            if (also) *also = idx;
        }
    }
}

void launch_argmax(const float* v, int n, int* out, int* also, cudaStream_t st) {
    k_argmax<<<1, 1024, 0, st>>>(v, n, out, also);
}

// ---------------------------------------------------------------------------
// host-in / host-out wrappers (context-free): allocate, copy, run, copy back.
// ---------------------------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        QW_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
        return 0;
    }
    template <typename T>
    T* as() { return reinterpret_cast<T*>(p); }
};
int finish() {
    QW_CUDA(cudaGetLastError());
    QW_CUDA(cudaDeviceSynchronize());
    return 0;
}
int need_device() {
    if (qwen_cuda_device_count() <= 0) {
        qw_set_error("no CUDA device: this library has no CPU path");
        return -1;
    }
    return 0;
}
} // namespace

extern "C" int qwen_cuda_q8_quantize(int8_t* q, float* s, const float* x, int n, int group) {
    if (need_device()) return -1;
    if (group <= 0 || n < 0) {
        qw_set_error("q8_quantize: bad n/group");
        return -2;
    }
    const int groups = n / group; // the tail n % group is ignored like the reference
    if (groups == 0) return 0;
    const size_t m = (size_t) groups * group;
    DevBuf dx, dq, ds;
    if (dx.alloc(m * 4) || dq.alloc(m) || ds.alloc((size_t) groups * 4)) return -1;
    QW_CUDA(cudaMemcpy(dx.p, x, m * 4, cudaMemcpyHostToDevice));
    if (group == 64)
        launch_quantize(dx.as<float>(), dq.as<int8_t>(), ds.as<float>(), (int) m, 0);
    else
        k_quantize_generic<<<groups, 128>>>(dx.as<float>(), dq.as<int8_t>(), ds.as<float>(), group);
    if (finish()) return -1;
    QW_CUDA(cudaMemcpy(q, dq.p, m, cudaMemcpyDeviceToHost));
    QW_CUDA(cudaMemcpy(s, ds.p, (size_t) groups * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int qwen_cuda_q8_dequantize(float* x, const int8_t* q, const float* s, int n, int group) {
    if (need_device()) return -1;
    if (group <= 0 || n < 0) {
        qw_set_error("q8_dequantize: bad n/group");
        return -2;
    }
    if (n == 0) return 0;
    const size_t ns = ((size_t) n + group - 1) / group;
    DevBuf dx, dq, ds;
    if (dx.alloc((size_t) n * 4) || dq.alloc(n) || ds.alloc(ns * 4)) return -1;
    QW_CUDA(cudaMemcpy(dq.p, q, n, cudaMemcpyHostToDevice));
    QW_CUDA(cudaMemcpy(ds.p, s, ns * 4, cudaMemcpyHostToDevice));
    k_dequantize<<<(n + 255) / 256, 256>>>(dq.as<int8_t>(), ds.as<float>(), dx.as<float>(), n, group);
    if (finish()) return -1;
    QW_CUDA(cudaMemcpy(x, dx.p, (size_t) n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

static int matmul_impl(float* out, int32_t* dots, const int8_t* xq, const float* xs, const int8_t* wq,
                       const float* ws, int n, int d, int group) {
    if (need_device()) return -1;
    if (group != 64) {
        qw_set_error("matmul: only block_size 64 is supported (every qwen3.c export uses 64)");
        return -2;
    }
    if (n <= 0 || d <= 0 || n % 64) {
        qw_set_error("matmul: n must be a positive multiple of 64");
        return -2;
    }
    const int ncols = qw_pad_cols(n), groups = n / 64;
    DevBuf dwq, dws, dw, dxq, dxs, dout, ddots;
    if (dwq.alloc((size_t) d * n) || dws.alloc((size_t) d * groups * 4) || dw.alloc(qw_row_bytes(n) * d)
        || dxq.alloc(ncols) || dxs.alloc((size_t) ncols / 64 * 4) || dout.alloc((size_t) d * 4))
        return -1;
    if (dots && ddots.alloc((size_t) d * groups * 4)) return -1;
    QW_CUDA(cudaMemcpy(dwq.p, wq, (size_t) d * n, cudaMemcpyHostToDevice));
    if (ws) {
        QW_CUDA(cudaMemcpy(dws.p, ws, (size_t) d * groups * 4, cudaMemcpyHostToDevice));
    } else {
        QW_CUDA(cudaMemset(dws.p, 0, (size_t) d * groups * 4));
    }
    QW_CUDA(cudaMemset(dxq.p, 0, ncols));
    QW_CUDA(cudaMemset(dxs.p, 0, (size_t) ncols / 64 * 4));
    QW_CUDA(cudaMemcpy(dxq.p, xq, n, cudaMemcpyHostToDevice));
    if (xs) QW_CUDA(cudaMemcpy(dxs.p, xs, (size_t) groups * 4, cudaMemcpyHostToDevice));
    launch_repack(dwq.as<int8_t>(), dws.as<float>(), n, 0, n, d, dw.as<uint8_t>(), 0, 1, 0);
    launch_gemv_sg(dw.as<uint8_t>(), dxq.as<int8_t>(), dxs.as<float>(), dout.as<float>(), d, n,
                   dots ? ddots.as<int32_t>() : nullptr, 0);
    if (finish()) return -1;
    if (out) QW_CUDA(cudaMemcpy(out, dout.p, (size_t) d * 4, cudaMemcpyDeviceToHost));
    if (dots) QW_CUDA(cudaMemcpy(dots, ddots.p, (size_t) d * groups * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int qwen_cuda_matmul(float* out, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                                int n, int d, int group) {
    return matmul_impl(out, nullptr, xq, xs, wq, ws, n, d, group);
}
extern "C" int qwen_cuda_matmul_group_dots(int32_t* dots, const int8_t* xq, const int8_t* wq, int n, int d,
                                           int group) {
    return matmul_impl(nullptr, dots, xq, nullptr, wq, nullptr, n, d, group);
}

extern "C" int qwen_cuda_rmsnorm(float* out, const float* x, const float* w, int size) {
    if (need_device()) return -1;
    if (size <= 0) return 0;
    DevBuf dx, dw, dout;
    if (dx.alloc((size_t) size * 4) || dw.alloc((size_t) size * 4) || dout.alloc((size_t) size * 4)) return -1;
    QW_CUDA(cudaMemcpy(dx.p, x, (size_t) size * 4, cudaMemcpyHostToDevice));
    QW_CUDA(cudaMemcpy(dw.p, w, (size_t) size * 4, cudaMemcpyHostToDevice));
    launch_rmsnorm(dout.as<float>(), dx.as<float>(), dw.as<float>(), size, 0);
    if (finish()) return -1;
    QW_CUDA(cudaMemcpy(out, dout.p, (size_t) size * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int qwen_cuda_softmax(float* x, int size) {
    if (need_device()) return -1;
    if (size <= 0) return 0;
    DevBuf dx;
    if (dx.alloc((size_t) size * 4)) return -1;
    QW_CUDA(cudaMemcpy(dx.p, x, (size_t) size * 4, cudaMemcpyHostToDevice));
    k_softmax<<<1, 1024>>>(dx.as<float>(), size);
    if (finish()) return -1;
    QW_CUDA(cudaMemcpy(x, dx.p, (size_t) size * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int qwen_cuda_rotary(float* x, int head_dim, const float* cos_host, const float* sin_host) {
    if (need_device()) return -1;
    const int half = head_dim / 2;
    if (half <= 0) return 0;
    DevBuf dx, dc, ds;
    if (dx.alloc((size_t) head_dim * 4) || dc.alloc((size_t) half * 4) || ds.alloc((size_t) half * 4)) return -1;
    QW_CUDA(cudaMemcpy(dx.p, x, (size_t) head_dim * 4, cudaMemcpyHostToDevice));
    QW_CUDA(cudaMemcpy(dc.p, cos_host, (size_t) half * 4, cudaMemcpyHostToDevice));
    QW_CUDA(cudaMemcpy(ds.p, sin_host, (size_t) half * 4, cudaMemcpyHostToDevice));
    k_rotary<<<(half + 127) / 128, 128>>>(dx.as<float>(), dc.as<float>(), ds.as<float>(), half);
    if (finish()) return -1;
    QW_CUDA(cudaMemcpy(x, dx.p, (size_t) head_dim * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int qwen_cuda_swiglu(float* x1, const float* x3, int size) {
    if (need_device()) return -1;
    if (size <= 0) return 0;
    DevBuf d1, d3;
    if (d1.alloc((size_t) size * 4) || d3.alloc((size_t) size * 4)) return -1;
    QW_CUDA(cudaMemcpy(d1.p, x1, (size_t) size * 4, cudaMemcpyHostToDevice));
    QW_CUDA(cudaMemcpy(d3.p, x3, (size_t) size * 4, cudaMemcpyHostToDevice));
    k_swiglu<<<(size + 255) / 256, 256>>>(d1.as<float>(), d3.as<float>(), size);
    if (finish()) return -1;
    QW_CUDA(cudaMemcpy(x1, d1.p, (size_t) size * 4, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int qwen_cuda_silu(float* y, const float* x, int size, int sigmoid_only) {
    if (need_device()) return -1;
    if (size <= 0) return 0;
    DevBuf dx, dy;
    if (dx.alloc((size_t) size * 4) || dy.alloc((size_t) size * 4)) return -1;
    QW_CUDA(cudaMemcpy(dx.p, x, (size_t) size * 4, cudaMemcpyHostToDevice));
    k_silu<<<(size + 255) / 256, 256>>>(dy.as<float>(), dx.as<float>(), size, sigmoid_only);
    if (finish()) return -1;
    QW_CUDA(cudaMemcpy(y, dy.p, (size_t) size * 4, cudaMemcpyDeviceToHost));
    return 0;
}
