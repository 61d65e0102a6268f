/*
 * host_forward.c -- the forward.h / q8.h entry points of the B200 build (host C).
 *
 * Every function keeps the reference's name, arguments and meaning
 * (reference: include/forward.h:31-140, include/q8.h:25-30) and forwards to exactly
 * one qwen_cuda_* call; no arithmetic happens on the CPU except the RoPE angle
 * table, which the parity contract wants from the host libm.
 * Failures print one "[Device] ..." line on stderr, like the reference's own
 * diagnostics, and leave outputs untouched.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "host_priv.h"

static void report(const char* who, int rc) {
    if (rc) {
        fprintf(stderr, "[Device] %s failed (%d): %s\n", who, rc, qwen_cuda_last_error());
    }
}

void q8_quantize(Q8Tensor* qt, float* x, int n, int block_size) {
    report("q8_quantize", qwen_cuda_q8_quantize(qt->q, qt->s, x, n, block_size));
}

void q8_dequantize(Q8Tensor* qt, float* x, int n, int block_size) {
    report("q8_dequantize", qwen_cuda_q8_dequantize(x, qt->q, qt->s, n, block_size));
}

void rmsnorm(float* out, float* x, float* w, int size) {
    report("rmsnorm", qwen_cuda_rmsnorm(out, x, w, size));
}

void softmax(float* x, int size) {
    report("softmax", qwen_cuda_softmax(x, size));
}

void matmul(float* out, Q8Tensor* x, Q8Tensor* w, int n, int d, int block_size) {
    report("matmul", qwen_cuda_matmul(out, x->q, x->s, w->q, w->s, n, d, block_size));
}

void rotary(float* x, int head_dim, int pos) {
    const int half = head_dim / 2;
    if (half <= 0) {
        return;
    }
    float* c = malloc(sizeof(float) * (size_t) half * 2);
    if (!c) {
        return;
    }
    float* s = c + half;
    for (int i = 0; i < half; i++) { /* src/forward.c:109-110, host libm */
        float angle = pos * powf(1e6f, -(float) i / half);
        c[i] = cosf(angle);
        s[i] = sinf(angle);
    }
    report("rotary", qwen_cuda_rotary(x, head_dim, c, s));
    free(c);
}

float sigmoid(float x) {
    float y = 0.0f;
    report("sigmoid", qwen_cuda_silu(&y, &x, 1, 1));
    return y;
}

float silu(float x) {
    float y = 0.0f;
    report("silu", qwen_cuda_silu(&y, &x, 1, 0));
    return y;
}

void swiglu(float* x1, float* x3, int size) {
    report("swiglu", qwen_cuda_swiglu(x1, x3, size));
}

void attention(Model* m, int layer, int pos) {
    report("attention", qwen_cuda_attention(model_cuda_ctx(m), layer, pos, m->state.q, m->state.x_rms_norm));
}

float* forward(Model* m, int token, int pos) {
    const int rc = qwen_cuda_forward(model_cuda_ctx(m), token, pos, m->state.logits);
    if (rc) {
        report("forward", rc);
        return NULL;
    }
    return m->state.logits;
}

/* Extension (no reference counterpart; see include/forward.h): the prompt loop of
 * src/completion.c:57-66 as one call. */
float* forward_prefill(Model* m, const int* tokens, int n, int pos) {
    const int rc = qwen_cuda_prefill(model_cuda_ctx(m), tokens, n, pos, m->state.logits);
    if (rc) {
        report("forward_prefill", rc);
        return NULL;
    }
    return m->state.logits;
}
