// common.cuh -- shared device helpers, the HBM weight layout and the device context.
//
// sm_100a only. No fast-math anywhere: the parity contract needs IEEE division,
// round-half-away quantisation and accurate expf.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qwen_cuda.h"

// ---------------------------------------------------------------------------
// HBM layout of a Q8_0 weight matrix W[d][n]  ("SG layout")
//
// The checkpoint stores int8[d*n] followed by fp32[d*n/64]. A GEMV over that
// layout needs two address streams per row. In HBM we interleave them at
// "super-group" granularity: 4 consecutive Q8_0 groups (256 columns) become one
// 272-byte record
//        [256 x int8 codes][4 x fp32 scales]
// and a row is ceil(n/256) records back to back. 272 = 17*16, so every record
// (and every row) is 16-byte aligned, a row segment is ONE contiguous byte range
// (one bulk copy), and a half-warp reads a record's codes with one conflict-free
// LDS.128 / LDG.128 per lane. Rows whose n is not a multiple of 256 are padded with
// zero codes and zero scales (they contribute exactly +0 to the fp32 fold).
// ---------------------------------------------------------------------------
#define QW_GROUP 64
#define QW_SG_COLS 256
#define QW_SG_BYTES 272
#define QW_HEAD_DIM 128

static inline __host__ __device__ int qw_sg_per_row(int n) { return (n + QW_SG_COLS - 1) / QW_SG_COLS; }
static inline __host__ __device__ size_t qw_row_bytes(int n) { return (size_t) qw_sg_per_row(n) * QW_SG_BYTES; }
static inline __host__ __device__ int qw_pad_cols(int n) { return qw_sg_per_row(n) * QW_SG_COLS; }

// ---------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------
void qw_set_error(const char* fmt, ...);

#define QW_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            qw_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

#define QW_CUDA_NULL(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            qw_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return nullptr;                                                             \
        }                                                                               \
    } while (0)

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact int32 dot of 16 int8 pairs
__device__ __forceinline__ int dot16(const int4& a, const int4& b) {
    int acc = __dp4a(a.x, b.x, 0);
    acc = __dp4a(a.y, b.y, acc);
    acc = __dp4a(a.z, b.z, acc);
    acc = __dp4a(a.w, b.w, acc);
    return acc;
}

// Q8_0 activation code: clamp(roundf(x / scale), -127, 127)   (reference q8.c:27-28)
__device__ __forceinline__ int q8_code(float x, float scale) {
    float r = roundf(__fdiv_rn(x, scale)); // IEEE division, half-away-from-zero
    r = fminf(fmaxf(r, -127.0f), 127.0f);
    return (int) r;
}
// Q8_0 scale from a group's absmax (reference q8.c:19-21)
__device__ __forceinline__ float q8_scale(float amax) {
    return (amax == 0.0f) ? 1e-6f : __fdiv_rn(amax, 127.0f);
}
// One Q8_0 matmul term: ((float) dot * w_scale) * x_scale   (reference forward.c:94-96)
__device__ __forceinline__ float q8_term(int dot, float ws, float xs) {
    return __fmul_rn(__fmul_rn((float) dot, ws), xs);
}
// rmsnorm scale: 1 / sqrt(ss / size + 1e-6)   (reference forward.c:21)
__device__ __forceinline__ float rms_rscale(float ss, int size) {
    return __fdiv_rn(1.0f, sqrtf(__fadd_rn(__fdiv_rn(ss, (float) size), 1e-6f)));
}
// silu(x) = x * (1 / (1 + expf(-x)))   (reference forward.c:122-129)
__device__ __forceinline__ float sigmoid_ref(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}
__device__ __forceinline__ float silu_ref(float x) { return __fmul_rn(x, sigmoid_ref(x)); }

// ---------------------------------------------------------------------------
// device context
// ---------------------------------------------------------------------------
struct QwenCudaCtx {
    int device;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    int num_sms;
    int path; // 0 = persistent kernel, 1 = per-op kernels

    // global shape
    int D, Hd, L, H, KVH, V, S, hd, G;
    // this rank's share (== global when tp_size == 1)
    int tp_rank, tp_size;
    int Hl, KVHl, Pl, Kl, Hdl, Vl;

    // weights, SG layout, one allocation each, per-layer stride in bytes
    uint8_t* w_qkv;  size_t w_qkv_stride;   // rows Pl + 2*Kl, n = D
    uint8_t* w_o;    size_t w_o_stride;     // rows D, n = Pl
    uint8_t* w_13;   size_t w_13_stride;    // rows 2*Hdl (w1/w3 interleaved), n = D
    uint8_t* w_2;    size_t w_2_stride;     // rows D, n = Hdl
    uint8_t* w_cls;                         // rows Vl, n = D
    uint8_t* w_emb;                         // rows V, n = D (aliases w_cls when tied and tp_size == 1)
    float *att_norm, *ffn_norm, *out_norm, *q_norm, *k_norm;
    float *rope_cos, *rope_sin;             // [S][hd/2]

    // KV cache [L][KVHl][S][hd] fp32
    float *k_cache, *v_cache;

    // activations
    float* x;        // [D] residual stream
    float* xb;       // [max(D, Pl, Hdl) padded] scratch
    float* qkv;      // [Pl + 2*Kl] raw projections
    float* q;        // [Pl] normalised + rotated queries
    float* att;      // [Pl] attention output
    float* h;        // [Hdl] swiglu output
    float* h13;      // [2*Hdl] raw w1/w3 outputs (per-op path)
    float* logits;   // [Vl]
    int8_t* aq;      // activation codes, padded to 256 columns, pad stays zero
    float* as;       // activation scales, padded
    int* token_dev;  // next token for the greedy chain
    int* argmax_out; // [max chain]
    int argmax_cap;

    // split-KV attention partials
    float *att_m, *att_l, *att_acc;
    int att_max_splits;

    // persistent kernel bookkeeping
    unsigned long long* bar_counter;
    unsigned long long bar_epoch;
    int* err_flag;

    void* nccl_comm;  // tensor-parallel communicator (tp_nccl.cu), NULL when tp_size == 1
    float* logits_all; // [V] gathered logits on tensor-parallel contexts
    void* mega;       // persistent-kernel state (decode_mega.cu)
    void* prefill;    // prefill activation buffers (prefill.cu), allocated on first use
    int layers_run;   // debug: run only the first n layers (-1 = all)
    int layer_begin;  // debug: first layer to run (0)
    float* x_inject;  // debug: [D] residual stream entering layer_begin instead of the embedding row
    int x_inject_on;
    uint8_t* dbg_codes; // debug: [4 * L + 1][dbg_codes_stride] Q8_0 activation vectors of the last persistent-kernel step (SG layout)
    size_t dbg_codes_stride;
    float* logits_pinned; // optional pinned bounce buffer
    float* sample_ws;     // workspace of the device sampler (sampler.cu), allocated on first use
    size_t bytes_weights, bytes_kv, bytes_read; // HBM bytes of this rank's weights / KV cache; checkpoint bytes it read at create
};

// ops.cu -- launchers used by both the op-level ABI and the per-op decode path
void launch_quantize(const float* x, int8_t* q, float* s, int n, cudaStream_t st);
void launch_rmsnorm(float* out, const float* x, const float* w, int size, cudaStream_t st);
void launch_gemv_sg(const uint8_t* w, const int8_t* xq, const float* xs, float* out, int rows, int n,
                    int32_t* dots, cudaStream_t st);
void launch_repack(const int8_t* src_q, const float* src_s, int src_n, int col0, int n, int rows,
                   uint8_t* dst, int dst_row0, int dst_row_step, cudaStream_t st);

// decode paths
int qw_decode_ops(QwenCudaCtx* c, int token, const int* token_dev, int pos);
int qw_decode_mega(QwenCudaCtx* c, int token, const int* token_dev, int pos);
int qw_mega_init(QwenCudaCtx* c);
int qw_prefill(QwenCudaCtx* c, const int* tokens_host, int n, int pos0);
void qw_prefill_free(QwenCudaCtx* c);
void qw_mega_free(QwenCudaCtx* c);
int qw_tp_allreduce(QwenCudaCtx* c, float* buf, size_t n);
int qw_tp_allgather(QwenCudaCtx* c, const float* src, float* dst, size_t n_per_rank);
void qw_tp_free(QwenCudaCtx* c);
void launch_argmax(const float* v, int n, int* out, int* also, cudaStream_t st);
