/*
 * qwen_cuda.h -- the C-ABI shim between qwen3.c's host code and the sm_100a kernels.
 *
 * Plain C types only (pointers, ints, sizes); no C++/torch types cross this line.
 * The reference has no FFI of its own: its forward path is a set of C functions
 * declared in include/forward.h, include/q8.h and include/model.h. libqwen3.so
 * re-exports exactly those symbols (host C in qwen3.c_b200/csrc/host_*.c) and each
 * of them is a thin call into one qwen_cuda_* entry point declared here. Each
 * declaration names the reference interface it stands behind.
 *
 * Conventions: every function that can fail returns 0 on success and a negative
 * code otherwise; qwen_cuda_last_error() returns a static, thread-local text for
 * the last failure. Nothing falls back to the CPU: with no CUDA device,
 * qwen_cuda_create() fails and the host layer reports it.
 *
 * "host" pointers are ordinary (or pinned) host memory; the shim stages them.
 */
#ifndef QWEN_CUDA_H
#define QWEN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct QwenCudaCtx QwenCudaCtx; /* opaque device context */

/* A Q8_0 tensor as it sits in the checkpoint mapping: numel int8 codes and
 * numel/group fp32 scales (reference: include/q8.h:16-19, src/model.c:120-142). */
typedef struct QwenCudaQ8 {
    const int8_t* q;
    const float* s;
} QwenCudaQ8;

/* Everything qwen_cuda_create needs, as host pointers into the checkpoint mapping.
 * Mirrors ModelParams + ModelWeights (reference: include/model.h:30-82) without
 * depending on those headers. Per-layer arrays have n_layers entries. */
typedef struct QwenCudaModelDesc {
    int dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len, head_dim,
        shared_classifier, group_size;
    const float* att_rms_norm; /* [L][dim] */
    const float* ffn_rms_norm; /* [L][dim] */
    const float* out_rms_norm; /* [dim] */
    const float* q_rms_norm;   /* [L][head_dim] */
    const float* k_rms_norm;   /* [L][head_dim] */
    QwenCudaQ8 emb;            /* [vocab][dim] */
    QwenCudaQ8 cls;            /* [vocab][dim]; same pointers as emb when tied */
    const QwenCudaQ8* wq;      /* [L] of [n_heads*head_dim][dim] */
    const QwenCudaQ8* wk;      /* [L] of [n_kv_heads*head_dim][dim] */
    const QwenCudaQ8* wv;
    const QwenCudaQ8* wo;      /* [L] of [dim][n_heads*head_dim] */
    const QwenCudaQ8* w1;      /* [L] of [hidden_dim][dim] */
    const QwenCudaQ8* w2;      /* [L] of [dim][hidden_dim] */
    const QwenCudaQ8* w3;
    /* RoPE tables computed by the HOST with libm so they are the reference's values
     * bit for bit (reference: src/forward.c:109-110): [seq_len][head_dim/2] each. */
    const float* rope_cos;
    const float* rope_sin;
} QwenCudaModelDesc;

/* Tensor-parallel placement of this context. tp_size 1 = whole model. */
typedef struct QwenCudaTp {
    int rank;
    int size;
} QwenCudaTp;

/* ---- lifecycle ------------------------------------------------------------- */
int qwen_cuda_device_count(void);
const char* qwen_cuda_last_error(void);
void qwen_cuda_clear_error(void);

/* Stands behind model_create's weight + state setup (reference: src/model.c:162-282,
 * 321-406): uploads every tensor of `desc` to `device`, repacks it into the kernels'
 * layout, allocates the device KV cache and scratch. NULL on failure. */
QwenCudaCtx* qwen_cuda_create(const QwenCudaModelDesc* desc, int device, QwenCudaTp tp);
/* Stands behind model_free (reference: src/model.c:491-500). NULL-safe. */
void qwen_cuda_destroy(QwenCudaCtx* ctx);

/* Tensor parallelism, one process per GPU (no reference counterpart: the reference is single
 * process). Rank 0 makes an id, the caller's own rendezvous hands the same 128
 * bytes to every rank, each rank then joins. Contexts with tp.size == 1 need neither.
 * qwen_cuda_tp_init also exchanges cudaIpc handles of the persistent kernel's flow arenas through the
 * new communicator and maps every peer's arena (NVLink peer access): the all-reduce after wo / w2 then
 * happens INSIDE the persistent kernel (each rank's GEMV epilogue stores its partial sums into every
 * rank's arena, readers add them in rank order). Without peer access the context stays on the per-op
 * kernels + ncclAllReduce path. */
int qwen_cuda_tp_unique_id(void* out128);
int qwen_cuda_tp_init(QwenCudaCtx* ctx, const void* id128);

/* Pinned host memory for state.logits (reference: src/model.c:342 callocs it). */
void* qwen_cuda_host_alloc(size_t bytes);
void qwen_cuda_host_free(void* p);

/* ---- the decode step --------------------------------------------------------
 * Stands behind forward() (reference: src/forward.c:225-350). Runs the step for
 * (token, pos), then copies vocab_size logits to `logits_host` (pinned or not) and
 * waits for them. */
int qwen_cuda_forward(QwenCudaCtx* ctx, int token, int pos, float* logits_host);
/* Same step, enqueue only: logits stay on the device (no copy, no wait). */
int qwen_cuda_forward_async(QwenCudaCtx* ctx, int token, int pos);
/* Copy the device logits of the last step to the host and wait. */
int qwen_cuda_logits_to_host(QwenCudaCtx* ctx, float* logits_host);
/* Prompt prefill: n tokens at positions pos0 .. pos0 + n - 1 in one pass (tcgen05 int8 GEMMs + batched
 * small ops + causal attention), KV cache rows written for all of them; logits_host (may be NULL) receives
 * the LAST token's logits. Same result as n forward() calls (reference: src/completion.c:57-66, which
 * discards all but the last logits). */
int qwen_cuda_prefill(QwenCudaCtx* ctx, const int* tokens, int n, int pos0, float* logits_host);
/* Greedy chain entirely on the device: n steps starting at (first_token, pos0); step i
 * feeds the argmax (lowest index on ties) of step i-1. out_tokens_host[n] receives the
 * n argmax tokens. Stands behind the completion loop's forward+sample pair when the
 * sampler is greedy (reference: src/completion.c:57-66). */
int qwen_cuda_decode_greedy(QwenCudaCtx* ctx, int first_token, int pos0, int n, int* out_tokens_host);
/* Stands behind sample() (reference: src/sampler.c:186-201: logits /= temperature, softmax, top-p, inverse CDF)
 * for the logits the LAST step of `ctx` left on the device: returns one token id instead of vocab_size floats.
 * temperature / top_p as stored in the reference's Sampler (sampler_create's clamps are re-applied); `coin` is the
 * caller's xorshift_float(&sampler->seed) (reference: src/xorshift.c:14-16), so the RNG stream stays the host's.
 * Returns 0 and *token_out; 1 when the fast path does not apply (more than 4096 tokens could be in the nucleus:
 * nearly flat distribution or top_p == 1) -- then copy the logits with qwen_cuda_logits_to_host and call the
 * reference's sample(); negative on errors. The token is also left on the device for qwen_cuda_decode_greedy-style
 * chaining. */
int qwen_cuda_sample(QwenCudaCtx* ctx, float temperature, float top_p, float coin, int* token_out);
/* Test hook: the same kernel on host logits. */
int qwen_cuda_sample_host(const float* logits_host, int vocab_size, float temperature, float top_p, float coin,
                          int* token_out);
/* Device-timed decode for bench.py: `warmup` untimed then `steps` timed steps at
 * positions pos0, pos0+1, ... (token fixed), CUDA events on the context's stream.
 * *ms_total = device milliseconds for the timed steps; *launches = kernels launched. */
int qwen_cuda_time_decode(QwenCudaCtx* ctx, int token, int pos0, int steps, int warmup,
                          float* ms_total, int* launches);
int qwen_cuda_sync(QwenCudaCtx* ctx);
/* 0 = persistent decode kernel (default), 1 = one kernel per op (debug / cross-check). */
int qwen_cuda_set_path(QwenCudaCtx* ctx, int path);
/* Which path the next step takes (0 / 1 as above; negative on a NULL context). Tensor-parallel contexts
 * report 0 once qwen_cuda_tp_init has mapped the peers' flow arenas (the persistent kernel then does the
 * all-reduce itself with NVLink peer stores), 1 if they fell back to per-op kernels + NCCL. */
int qwen_cuda_get_path(const QwenCudaCtx* ctx);

/* Bytes this context (= this tensor-parallel rank) holds in HBM -- weights in the kernels' layout, KV cache -- and the
 * checkpoint bytes it read at create (a rank reads only the rows / column windows it owns). Any pointer may be NULL. */
int qwen_cuda_memory(const QwenCudaCtx* ctx, size_t* weight_bytes, size_t* kv_bytes, size_t* read_bytes);

/* ---- KV cache access (test + long-context parity hooks) ----------------------
 * Host side uses the reference's order: [npos][n_kv_heads*head_dim] for one layer
 * (reference: src/model.c:353-361, src/forward.c:244-248). */
int qwen_cuda_kv_write(QwenCudaCtx* ctx, int layer, int pos0, int npos, const float* k_host, const float* v_host);
int qwen_cuda_kv_read(QwenCudaCtx* ctx, int layer, int pos0, int npos, float* k_host, float* v_host);
/* Read back a device activation by name ("x", "q", "att", "h", "aq", "as") for layer-by-layer
 * parity tests. Returns the number of elements copied or a negative code. */
int qwen_cuda_debug_read(QwenCudaCtx* ctx, const char* what, void* host, size_t max_bytes);
/* Test hook: the quantiser fused into the persistent kernel's prologues (reciprocal candidate + exact
 * fallback), host in / host out; must equal q8_quantize bit for bit. n % 64 == 0. */
int qwen_cuda_debug_quantize_fused(int8_t* q, float* s, const float* x, int n);
/* Debug: run only the first n layers of the step (then final norm + classifier); -1 = all. */
int qwen_cuda_debug_set_layers(QwenCudaCtx* ctx, int n);
/* Debug (layer-by-layer parity with teacher forcing): the next steps run layers [l0, l1) only (l1 = -1: to the last),
 * then the final norm and the classifier; x_host (dim floats, may be NULL) replaces the embedding row as the residual
 * stream entering layer l0. (0, -1, NULL) restores the normal step. Both decode paths honour it. */
int qwen_cuda_debug_set_window(QwenCudaCtx* ctx, int l0, int l1, const float* x_host);
/* Debug (flip audit): record the Q8_0 activation vector (codes + scales) the persistent kernel feeds to every GEMV of the
 * following steps. which = 4 * layer + {0: wq|wk|wv input, 1: wo input, 2: w1/w3 input, 3: w2 input}; 4 * n_layers = the
 * classifier's input. read unpacks n codes and n / 64 scales of the last step (n = the GEMV's column count). */
int qwen_cuda_debug_codes_enable(QwenCudaCtx* ctx, int on);
int qwen_cuda_debug_codes_read(QwenCudaCtx* ctx, int which, int8_t* q, float* s, int n);
/* Debug: per-CTA phase timestamps (globaltimer ns) of the persistent kernel, [grid][L+1][16] (+ per-warp cycle counters in
 * -DQW_UNITPROF builds). enable returns the element count; read returns the grid size. */
int qwen_cuda_debug_profile_enable(QwenCudaCtx* ctx);
int qwen_cuda_debug_profile_read(QwenCudaCtx* ctx, unsigned long long* host, size_t max_elems);

/* ---- single ops, host in / host out (context-free) ---------------------------
 * Each stands behind the same-named function of include/forward.h / q8.h. */
int qwen_cuda_q8_quantize(int8_t* q, float* s, const float* x, int n, int group);          /* q8.c:5-30 */
int qwen_cuda_q8_dequantize(float* x, const int8_t* q, const float* s, int n, int group);  /* q8.c:32-37 */
int qwen_cuda_matmul(float* out, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                     int n, int d, int group);                                           /* forward.c:79-101 */
/* Test hook: the exact int32 dot of every (row, group), dots[d][n/group]. */
int qwen_cuda_matmul_group_dots(int32_t* dots, const int8_t* xq, const int8_t* wq, int n, int d, int group);
/* Batched matmul for prefill: T tokens at once on the tcgen05 tensor cores (csrc/prefill_gemm.cu).
 * Same arithmetic as matmul() per token -- outputs are bit-identical to the reference's. xq [T][n],
 * xs [T][n/64], out [T][d]; dots (optional) receives the exact int32 group dots [T][d][n/64];
 * reps >= 1 runs of the kernel, *ms (optional) = best device time of one run. */
int qwen_cuda_matmul_batch(float* out, int32_t* dots, const int8_t* xq, const float* xs, const int8_t* wq,
                           const float* ws, int n, int d, int T, int reps, float* ms);
/* Measured dense int8 tensor throughput of the device in tera-ops/s (2 ops per MAC): `iters` back-to-back
 * tcgen05.mma.kind::i8 M128 N256 K32 per SM on resident operands, best of `reps` runs. bench.py's prefill fraction is
 * quoted against this number, not the nominal 4.5 POPS. */
int qwen_cuda_int8_peak(int iters, int reps, float* tops);

/* Test hooks of the prefill kernels (csrc/prefill.cu, csrc/prefill_gemm.cu).
 * debug_attn_prefill: the chunk attention kernels on host data -- q [T][n_heads][128], k / v [pos0 + T][n_kv_heads * 128]
 * (the reference's cache layout, src/forward.c:141-195), out [T][n_heads][128]; variant 1 = per-warp kernel, 2 = tiled kernel.
 * debug_attn_plan / debug_gemm_plan: host-only (no GPU): the fixed key blocks of the chunk attention and the GEMM tile shape. */
int qwen_cuda_debug_attn_prefill(float* out, const float* q, const float* k, const float* v, int n_heads, int n_kv_heads, int pos0,
                                 int T, int variant);
int qwen_cuda_debug_attn_plan(int pos0, int T, int* parts, int* k01, int max_parts);
int qwen_cuda_debug_gemm_plan(int d, int T, int sms, int* tiles);
int qwen_cuda_rmsnorm(float* out, const float* x, const float* w, int size);                /* forward.c:12-28 */
int qwen_cuda_softmax(float* x, int size);                                                  /* forward.c:34-77 */
/* cos/sin: head_dim/2 host-computed values for this position (forward.c:109-110). */
int qwen_cuda_rotary(float* x, int head_dim, const float* cos_host, const float* sin_host); /* forward.c:104-118 */
int qwen_cuda_swiglu(float* x1, const float* x3, int size);                                 /* forward.c:134-139 */
int qwen_cuda_silu(float* y, const float* x, int size, int sigmoid_only);                   /* forward.c:122-129 */
/* Stands behind attention() (reference: src/forward.c:141-195): q_host is n_heads*head_dim
 * floats (normalised + rotated), out_host receives n_heads*head_dim floats; K/V come from
 * the device cache of `layer`, slots 0..pos. */
int qwen_cuda_attention(QwenCudaCtx* ctx, int layer, int pos, const float* q_host, float* out_host);

#ifdef __cplusplus
}
#endif
#endif /* QWEN_CUDA_H */
