/*
 * qwen3_oracle.h -- CPU restatement of qwen3.c's forward hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product library (libqwen3.so, built
 * from qwen3.c_b200/csrc) includes, links or calls this file. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load liboracle.so, and only as the checker or the timed CPU baseline.
 *
 * Parity status: PINNED. tests/test_oracle_pinned.py checks every function here
 *   (a) bit-for-bit against the reference's own compiled sources
 *       (oracle/_ref/libqwen3_ref_strict.so, built by oracle/Makefile from
 *       /root/reference/src/{q8,model,forward}.c) when that library is present,
 *   (b) bit-for-bit against the committed fixtures in tests/golden/, which
 *       were produced by that same compiled reference (tests/golden/make_golden.py).
 * The reference ships no tests, golden vectors or known-answer fixtures of its
 * own, so (a)/(b) are the only pins that exist.
 *
 * Everything is strictly serial, IEEE fp32, compiled -O2 -ffp-contract=off with
 * no -march flag, so it reproduces the reference's strict build
 * (gcc -O2 -DNDEBUG, one thread) operation for operation.
 */
#ifndef QWEN3_ORACLE_H
#define QWEN3_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Q8_0 primitives (reference: src/q8.c) ---- */
void orc_q8_quantize(int8_t* q, float* s, const float* x, int n, int gs);
void orc_q8_dequantize(float* x, const int8_t* q, const float* s, int n, int gs);

/* ---- ops (reference: src/forward.c) ---- */
/* dots[i * (n/gs) + g] = exact int32 dot of group g of row i with x. */
void orc_group_dots(int32_t* dots, const int8_t* xq, const int8_t* wq, int n, int d, int gs);
void orc_matmul(float* out, const int8_t* xq, const float* xs,
                const int8_t* wq, const float* ws, int n, int d, int gs);
void orc_rmsnorm(float* out, const float* x, const float* w, int size);
void orc_softmax(float* x, int size);
void orc_rotary(float* x, int head_dim, int pos);
float orc_sigmoid(float x);
float orc_silu(float x);
void orc_swiglu(float* x1, const float* x3, int size);
/* All query heads of one layer. k_layer / v_layer point at that layer's
 * [seq_len][kv_dim] slab (reference layout). scores is n_heads*seq_len scratch. */
void orc_attention(float* out, const float* q, const float* k_layer, const float* v_layer,
                   float* scores, int n_heads, int n_kv_heads, int head_dim,
                   int seq_len, int pos);

/* ---- checkpoint + whole forward (reference: src/model.c, src/forward.c:225-350) ---- */
typedef struct OrcQ8 {
    const int8_t* q;
    const float* s;
} OrcQ8;

typedef struct OrcModel {
    /* header, reference: include/model.h:30-43 */
    int dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len, head_dim,
        shared_classifier, group_size;
    /* views into the mapping */
    const float *att_norm, *ffn_norm, *out_norm, *q_norm, *k_norm;
    OrcQ8 emb, cls;
    OrcQ8 *wq, *wk, *wv, *wo, *w1, *w2, *w3;
    /* state */
    float *x, *xb, *q, *att, *h1, *h3, *scores, *logits, *k_cache, *v_cache;
    int8_t* aq; /* activation codes, max(P, Hd) */
    float* as;  /* activation scales */
    /* mapping */
    void* map;
    size_t map_len;
    /* optional trace of the LAST forward call (allocated by orc_trace_enable) */
    int trace_on;
    int8_t* tr_qkv_in_q;  /* [L][dim]   codes fed to wq/wk/wv */
    float* tr_qkv_in_s;   /* [L][dim/gs] */
    float* tr_q_rot;      /* [L][P]     q after norm+rope */
    float* tr_att_out;    /* [L][P]     attention output */
    float* tr_x_after_att;/* [L][dim]   residual after wo */
    float* tr_x_after_ffn;/* [L][dim]   residual after w2 */
    float* tr_h;          /* [L][Hd]    swiglu output */
    /* the other three quantised GEMV inputs of every layer, the classifier's, and the fp32 vectors they came from */
    int8_t* tr_wo_in_q;   /* [L][P]     codes fed to wo */
    float* tr_wo_in_s;    /* [L][P/gs] */
    int8_t* tr_ffn_in_q;  /* [L][dim]   codes fed to w1/w3 */
    float* tr_ffn_in_s;   /* [L][dim/gs] */
    int8_t* tr_w2_in_q;   /* [L][Hd]    codes fed to w2 */
    float* tr_w2_in_s;    /* [L][Hd/gs] */
    int8_t* tr_cls_in_q;  /* [dim]      codes fed to the classifier */
    float* tr_cls_in_s;   /* [dim/gs] */
    float* tr_x_final;    /* [dim]      residual stream entering the final norm */
    float* tr_x_normed;   /* [dim]      after the final norm (what the classifier's quantiser sees) */
} OrcModel;

OrcModel* orc_model_open(const char* path, int seq_len_override);
void orc_model_close(OrcModel* m);
int orc_trace_enable(OrcModel* m);
float* orc_forward(OrcModel* m, int token, int pos);
float* orc_forward_ex(OrcModel* m, int token, int pos, int want_logits);

/* argmax with lowest-index tie break; margin = top1 - top2 (H7 bookkeeping). */
int orc_argmax(const float* v, int n, float* margin);

/* sampler restatement (reference src/sampler.c, src/xorshift.c) */
uint32_t orc_xorshift_int32(uint64_t* state);
float orc_xorshift_float(uint64_t* state);
void orc_sampler_clamp(float* temperature, float* top_p);
int orc_sample(float* logits, int vocab_size, float temperature, float top_p, float coin, float* gap);

#ifdef __cplusplus
}
#endif
#endif
