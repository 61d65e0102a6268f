/*
 * forward.h -- the transformer step and its building-block ops, B200 build.
 *
 * Same eleven prototypes as the reference (reference: include/forward.h:31-140)
 * so that src/completion.c:59,321 and src/sampler.c:196 link unchanged.
 *
 * forward() runs the whole token step on the device (one persistent sm_100a
 * kernel: csrc/decode_pw.cuh on one GPU, csrc/decode_mega.cu under tensor
 * parallelism) and copies vocab_size logits into the pinned host
 * buffer it returns. The nine small ops and attention() are host-in / host-out
 * wrappers over single-op kernels (csrc/ops.cu); they exist for callers such as
 * the sampler's softmax() and for op-level parity tests, not for speed.
 * Nothing here computes on the CPU.
 */
#ifndef QWEN_FORWARD_H
#define QWEN_FORWARD_H

#include "q8.h"
#include "model.h"

#ifdef __cplusplus
extern "C" {
#endif

/* out[i] = w[i] * (x[i] / sqrt(mean(x^2) + 1e-6)); out may alias x
 * (reference: src/forward.c:12-28). */
void rmsnorm(float* out, float* x, float* w, int size);

/* In-place exp(x - max) / sum over `size` host floats
 * (reference: src/forward.c:34-77). */
void softmax(float* x, int size);

/*
 * Q8_0 matrix-vector product, W[d][n] times x[n]:
 *   out[i] = sum over groups g of  ((float) dot_g * w.s[(i*n)/bs + g]) * x.s[g]
 * with dot_g the exact int32 dot of the group's int8 codes
 * (reference: src/forward.c:79-101).
 */
void matmul(float* out, Q8Tensor* x, Q8Tensor* w, int n, int d, int block_size);

/* Half-split rotary embedding, theta = 1e6, pairs (i, i + head_dim/2), in place
 * (reference: src/forward.c:104-118). */
void rotary(float* x, int head_dim, int pos);

/* 1 / (1 + exp(-x))  (reference: src/forward.c:122-124). */
float sigmoid(float x);

/* x * sigmoid(x)  (reference: src/forward.c:127-129). */
float silu(float x);

/* x1[i] = silu(x1[i]) * x3[i]  (reference: src/forward.c:134-139). */
void swiglu(float* x1, float* x3, int size);

/*
 * Grouped-query attention for one layer at position `pos`
 * (reference: src/forward.c:141-195). B200 build: reads m->state.q (host,
 * already normalised and rotated), attends over the DEVICE KV cache slots
 * 0..pos of `layer`, writes n_heads*head_dim floats to m->state.x_rms_norm.
 */
void attention(Model* m, int layer, int pos);

/*
 * One decode step: embedding row of `token`, every layer, final norm,
 * classifier. Writes KV slot `pos` of every layer on the device and returns
 * m->state.logits (pinned host, vocab_size floats, valid until the next call).
 * Returns NULL if pos is outside [0, seq_len) or the device reports an error
 * (the reference trusts pos; reference: src/forward.c:225-350).
 */
float* forward(Model* m, int token, int pos);

/*
 * EXTENSION of the B200 build (the reference has no such function): the prompt
 * loop of src/completion.c:57-66 -- forward() once per prompt token, all logits
 * but the last discarded -- as one call. Runs n tokens at positions pos .. pos+n-1
 * through the layers together (tensor-core int8 GEMMs), writes their KV slots and
 * returns m->state.logits holding the LAST token's logits; NULL on error.
 * A generation loop switches to it by replacing its prompt loop with one call
 * (INTEGRATION.md); callers that never use it are unaffected.
 */
float* forward_prefill(Model* m, const int* tokens, int n, int pos);

#ifdef __cplusplus
}
#endif

#endif /* QWEN_FORWARD_H */
