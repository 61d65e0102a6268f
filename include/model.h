/*
 * model.h -- checkpoint header, weight views, forward scratch, Model handle.
 *
 * ABI mirror of the reference header (reference: include/model.h:30-168): the
 * four public structs keep the reference's field order and types so that
 * unchanged callers (src/qwen.c:27, src/completion.c:265, examples/model.c:16)
 * read the same offsets.
 *
 * What differs in the B200 build is where the bytes live:
 *   - Weight Q8Tensor views still point into the read-only mmap of the .bin
 *     (they are the upload source); the copies the kernels read are repacked
 *     in HBM and owned by the private device context behind the Model.
 *   - weights.fe (the reference's host fp32 embedding table, model.c:201-206)
 *     is NULL: the embedding row is dequantised on the device per token.
 *   - state.logits is PINNED host memory of vocab_size floats; forward()
 *     returns it and callers may scribble on it (src/sampler.c:189-196 does).
 *   - state.x, x_rms_norm, q, mlp_in, mlp_gate, qx, qh are small host staging
 *     buffers used only by the op-level wrappers in forward.h.
 *   - state.k / v / scores / k_cache / v_cache are NULL: the KV cache is device
 *     resident. Use qwen_cuda_kv_read / qwen_cuda_kv_write (qwen_cuda.h) to
 *     move slices across in the reference's [layer][pos][kv_dim] order.
 */
#ifndef QWEN_MODEL_H
#define QWEN_MODEL_H

#include "q8.h"
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QWEN_MAGIC 0x7177656E /* "qwen" */
#define QWEN_VERSION 1

/* First 48 bytes of the .bin, twelve little-endian int32 (model.h:30-43). */
typedef struct ModelParams {
    int magic;
    int version;
    int dim;               /* residual width D */
    int hidden_dim;        /* FFN width Hd */
    int n_layers;
    int n_heads;           /* query heads */
    int n_kv_heads;        /* key/value heads (GQA) */
    int vocab_size;
    int seq_len;           /* context window; may be lowered by model_create */
    int head_dim;
    int shared_classifier; /* 1: classifier aliases the embedding table */
    int block_size;        /* Q8_0 group length, 64 in every export */
} ModelParams;

/* Per-layer arrays of views into the checkpoint (model.h:55-82). */
typedef struct ModelWeights {
    Q8Tensor* wq; /* [L] each [n_heads*head_dim][dim] */
    Q8Tensor* wk; /* [L] each [n_kv_heads*head_dim][dim] */
    Q8Tensor* wv; /* [L] each [n_kv_heads*head_dim][dim] */
    Q8Tensor* wo; /* [L] each [dim][n_heads*head_dim] */

    Q8Tensor* w1; /* [L] gate_proj [hidden_dim][dim] */
    Q8Tensor* w2; /* [L] down_proj [dim][hidden_dim] */
    Q8Tensor* w3; /* [L] up_proj   [hidden_dim][dim] */

    Q8Tensor* cls; /* [vocab][dim]; == qe when shared_classifier */

    Q8Tensor* qe; /* [vocab][dim] token embedding, quantised */
    float* fe;    /* NULL in this build (no host fp32 table) */

    float* att_rms_norm; /* [L][dim] */
    float* ffn_rms_norm; /* [L][dim] */
    float* out_rms_norm; /* [dim] */

    float* q_rms_norm; /* [L][head_dim], one vector per layer for all heads */
    float* k_rms_norm; /* [L][head_dim] */
} ModelWeights;

/* Forward scratch (model.h:92-117); see the file comment for what is NULL. */
typedef struct ForwardState {
    float* x;
    float* x_rms_norm;

    float* q;
    float* k;
    float* v;
    float* scores;

    float* mlp_in;
    float* mlp_gate;

    float* logits;

    float* k_cache;
    float* v_cache;

    Q8Tensor qx;
    Q8Tensor qh;
} ForwardState;

/*
 * The handle callers hold. model_create() really allocates a larger private
 * record whose first member is this struct (callers never allocate a Model
 * themselves: src/qwen.c:27, examples/model.c:16), so the device context rides
 * along without changing the public layout (model.h:123-129).
 */
typedef struct Model {
    ModelParams params;
    ModelWeights weights;
    ForwardState state;
    void* data;   /* base of the checkpoint mmap */
    ssize_t size; /* its length in bytes */
} Model;

/*
 * mmap `path`, validate the header, upload and repack every tensor into HBM,
 * allocate the device KV cache for min(override_seq_len, header seq_len)
 * positions (override ignored unless 0 < override <= header, model.c:74-76).
 * Returns NULL (after a "[Tag] ..." line on stderr) on any failure, including
 * "no CUDA device": there is no CPU fallback. (reference: src/model.c:451-489)
 */
Model* model_create(const char* path, int override_seq_len);

/* Releases device memory, pinned buffers, the mmap and the handle. NULL-safe
 * (reference: src/model.c:491-500). */
void model_free(Model* m);

#ifdef __cplusplus
}
#endif

#endif /* QWEN_MODEL_H */
