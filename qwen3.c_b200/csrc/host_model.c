/*
 * host_model.c -- model_create / model_free for the B200 build (host C).
 *
 * Same contract as the reference loader (reference: src/model.c:451-500): map the
 * checkpoint, validate the 48-byte header, carve tensor views out of the mapping in
 * export order (src/model.c:162-244) and hand the caller a Model*. Instead of
 * dequantising the embedding table and calloc'ing a host KV cache, the views are
 * passed to qwen_cuda_create(), which uploads and repacks them into HBM.
 *
 * Environment (no new CLI flags, so examples/qwen.c stays untouched):
 *   QWEN_CUDA_DEVICE   device ordinal, default 0
 *   QWEN_CUDA_PATH     "ops" selects the one-kernel-per-op debug path
 *   QWEN_CUDA_TP_RANK / QWEN_CUDA_TP_SIZE   tensor-parallel placement of this process (one process per
 *                      GPU); the caller then hands every rank the same NCCL id via qwen_cuda_tp_init
 */
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "host_priv.h"

static Q8Tensor* carve(const unsigned char** cur, const unsigned char* end, int count, size_t numel, int bs) {
    Q8Tensor* t = calloc((size_t) count, sizeof(Q8Tensor));
    if (!t) {
        return NULL;
    }
    for (int i = 0; i < count; i++) {
        const size_t need = numel + (numel / (size_t) bs) * sizeof(float);
        if ((size_t) (end - *cur) < need) {
            free(t);
            return NULL;
        }
        t[i].q = (int8_t*) *cur;
        t[i].s = (float*) (*cur + numel);
        *cur += need;
    }
    return t;
}

static QwenCudaQ8* as_desc(const Q8Tensor* t, int count) {
    QwenCudaQ8* d = calloc((size_t) count, sizeof(QwenCudaQ8));
    for (int i = 0; d && i < count; i++) {
        d[i].q = t[i].q;
        d[i].s = t[i].s;
    }
    return d;
}

/* RoPE tables with the reference's exact expression (src/forward.c:109-110):
 * angle = pos * powf(1e6f, -(float) i / half); cosf / sinf from the host libm. */
static int rope_tables(int seq_len, int head_dim, float** cos_out, float** sin_out) {
    const int half = head_dim / 2;
    float* c = malloc((size_t) seq_len * half * sizeof(float));
    float* s = malloc((size_t) seq_len * half * sizeof(float));
    if (!c || !s) {
        free(c);
        free(s);
        return -1;
    }
    for (int pos = 0; pos < seq_len; pos++) {
        for (int i = 0; i < half; i++) {
            float angle = pos * powf(1e6f, -(float) i / half);
            c[(size_t) pos * half + i] = cosf(angle);
            s[(size_t) pos * half + i] = sinf(angle);
        }
    }
    *cos_out = c;
    *sin_out = s;
    return 0;
}

static void free_views(ModelWeights* w, int shared) {
    free(w->qe);
    free(w->wq);
    free(w->wk);
    free(w->wv);
    free(w->wo);
    free(w->w1);
    free(w->w2);
    free(w->w3);
    if (!shared) {
        free(w->cls);
    }
}

Model* model_create(const char* path, int override_seq_len) {
    if (!path) {
        return NULL;
    }
    ModelPriv* mp = calloc(1, sizeof(ModelPriv));
    if (!mp) {
        return NULL;
    }
    Model* m = &mp->pub;

    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        fprintf(stderr, "[Model] cannot open %s: %s\n", path, strerror(errno));
        free(mp);
        return NULL;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 256) {
        fprintf(stderr, "[Model] %s is not a checkpoint (too short)\n", path);
        close(fd);
        free(mp);
        return NULL;
    }
    void* map = mmap(NULL, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) {
        fprintf(stderr, "[Model] mmap failed: %s\n", strerror(errno));
        free(mp);
        return NULL;
    }
    m->data = map;
    m->size = (ssize_t) st.st_size;

    ModelParams* p = &m->params;
    memcpy(p, map, sizeof(ModelParams));
    if (p->magic != QWEN_MAGIC || p->version != QWEN_VERSION) {
        fprintf(stderr, "[Params] bad magic/version %x/%d\n", p->magic, p->version);
        goto fail_map;
    }
    if (p->block_size <= 0 || p->dim <= 0 || p->n_layers <= 0 || p->n_heads <= 0 || p->n_kv_heads <= 0
        || p->vocab_size <= 0 || p->seq_len <= 0 || p->head_dim <= 0 || p->hidden_dim <= 0
        || p->dim % p->block_size || p->hidden_dim % p->block_size) {
        fprintf(stderr, "[Params] inconsistent header\n");
        goto fail_map;
    }
    if (override_seq_len > 0 && override_seq_len <= p->seq_len) { /* src/model.c:74-76 */
        p->seq_len = override_seq_len;
    }
    /* same banner as the reference (scripts may grep it; src/model.c:81-94) */
    fprintf(stderr, "[Params] magic=%x\n", p->magic);
    fprintf(stderr, "[Params] version=%d\n", p->version);
    fprintf(stderr, "[Params] hidden_size=%d\n", p->dim);
    fprintf(stderr, "[Params] intermediate_size=%d\n", p->hidden_dim);
    fprintf(stderr, "[Params] num_hidden_layers=%d\n", p->n_layers);
    fprintf(stderr, "[Params] num_attention_heads=%d\n", p->n_heads);
    fprintf(stderr, "[Params] num_kv_heads=%d\n", p->n_kv_heads);
    fprintf(stderr, "[Params] vocab_size=%d\n", p->vocab_size);
    fprintf(stderr, "[Params] seq_len=%d\n", p->seq_len);
    fprintf(stderr, "[Params] head_dim=%d\n", p->head_dim);
    fprintf(stderr, "[Params] shared_classifier=%d\n", p->shared_classifier);
    fprintf(stderr, "[Params] block_size=%d\n", p->block_size);

    const int L = p->n_layers, D = p->dim, Hd = p->hidden_dim, bs = p->block_size;
    const int P = p->n_heads * p->head_dim, K = p->n_kv_heads * p->head_dim;
    ModelWeights* w = &m->weights;
    const unsigned char* end = (const unsigned char*) map + st.st_size;
    const unsigned char* cur = (const unsigned char*) map + 256;
    const size_t n_norm = (size_t) L * D * 2 + D + (size_t) L * p->head_dim * 2;
    if ((size_t) (end - cur) < n_norm * sizeof(float)) {
        fprintf(stderr, "[Weights] checkpoint truncated\n");
        goto fail_map;
    }
    float* f = (float*) cur;
    w->att_rms_norm = f;  f += (size_t) L * D;
    w->ffn_rms_norm = f;  f += (size_t) L * D;
    w->out_rms_norm = f;  f += D;
    w->q_rms_norm = f;    f += (size_t) L * p->head_dim;
    w->k_rms_norm = f;    f += (size_t) L * p->head_dim;
    cur = (const unsigned char*) f;

    w->qe = carve(&cur, end, 1, (size_t) p->vocab_size * D, bs);
    w->fe = NULL; /* no host fp32 embedding table in this build */
    w->wq = carve(&cur, end, L, (size_t) D * P, bs);
    w->wk = carve(&cur, end, L, (size_t) D * K, bs);
    w->wv = carve(&cur, end, L, (size_t) D * K, bs);
    w->wo = carve(&cur, end, L, (size_t) P * D, bs);
    w->w1 = carve(&cur, end, L, (size_t) D * Hd, bs);
    w->w2 = carve(&cur, end, L, (size_t) Hd * D, bs);
    w->w3 = carve(&cur, end, L, (size_t) D * Hd, bs);
    w->cls = p->shared_classifier ? w->qe : carve(&cur, end, 1, (size_t) p->vocab_size * D, bs);
    if (!w->qe || !w->wq || !w->wk || !w->wv || !w->wo || !w->w1 || !w->w2 || !w->w3 || !w->cls) {
        fprintf(stderr, "[Weights] checkpoint truncated or out of memory\n");
        goto fail_views;
    }

    /* host staging for the op-level wrappers + pinned logits */
    ForwardState* s = &m->state;
    s->x = calloc((size_t) D, sizeof(float));
    s->x_rms_norm = calloc((size_t) (P > D ? P : D), sizeof(float));
    s->q = calloc((size_t) P, sizeof(float));
    s->mlp_in = calloc((size_t) Hd, sizeof(float));
    s->mlp_gate = calloc((size_t) Hd, sizeof(float));
    s->qx.q = calloc((size_t) (P > D ? P : D), 1);
    s->qx.s = calloc((size_t) (P > D ? P : D) / bs + 1, sizeof(float));
    s->qh.q = calloc((size_t) Hd, 1);
    s->qh.s = calloc((size_t) Hd / bs + 1, sizeof(float));
    s->logits = qwen_cuda_host_alloc((size_t) p->vocab_size * sizeof(float));
    if (!s->x || !s->x_rms_norm || !s->q || !s->mlp_in || !s->mlp_gate || !s->qx.q || !s->qx.s || !s->qh.q
        || !s->qh.s || !s->logits) {
        fprintf(stderr, "[ForwardState] Allocation failed! %s\n", qwen_cuda_last_error());
        goto fail_state;
    }

    /* device side */
    QwenCudaModelDesc d;
    memset(&d, 0, sizeof d);
    d.dim = D; d.hidden_dim = Hd; d.n_layers = L; d.n_heads = p->n_heads; d.n_kv_heads = p->n_kv_heads;
    d.vocab_size = p->vocab_size; d.seq_len = p->seq_len; d.head_dim = p->head_dim;
    d.shared_classifier = p->shared_classifier; d.group_size = bs;
    d.att_rms_norm = w->att_rms_norm; d.ffn_rms_norm = w->ffn_rms_norm; d.out_rms_norm = w->out_rms_norm;
    d.q_rms_norm = w->q_rms_norm; d.k_rms_norm = w->k_rms_norm;
    d.emb.q = w->qe->q; d.emb.s = w->qe->s;
    d.cls.q = w->cls->q; d.cls.s = w->cls->s;
    QwenCudaQ8* dq = as_desc(w->wq, L), *dk = as_desc(w->wk, L), *dv = as_desc(w->wv, L), *dox = as_desc(w->wo, L);
    QwenCudaQ8* d1 = as_desc(w->w1, L), *d2 = as_desc(w->w2, L), *d3 = as_desc(w->w3, L);
    float *rc = NULL, *rs = NULL;
    QwenCudaCtx* ctx = NULL;
    if (dq && dk && dv && dox && d1 && d2 && d3 && rope_tables(p->seq_len, p->head_dim, &rc, &rs) == 0) {
        d.wq = dq; d.wk = dk; d.wv = dv; d.wo = dox; d.w1 = d1; d.w2 = d2; d.w3 = d3;
        d.rope_cos = rc; d.rope_sin = rs;
        const char* dev = getenv("QWEN_CUDA_DEVICE");
        QwenCudaTp tp = {0, 1};
        const char* tr = getenv("QWEN_CUDA_TP_RANK");
        const char* ts = getenv("QWEN_CUDA_TP_SIZE");
        if (tr && ts) {
            tp.rank = atoi(tr);
            tp.size = atoi(ts);
        }
        ctx = qwen_cuda_create(&d, dev ? atoi(dev) : 0, tp);
    }
    free(dq); free(dk); free(dv); free(dox); free(d1); free(d2); free(d3); free(rc); free(rs);
    if (!ctx) {
        fprintf(stderr, "[Device] %s\n", qwen_cuda_last_error());
        goto fail_state;
    }
    const char* pathsel = getenv("QWEN_CUDA_PATH");
    if (pathsel && strcmp(pathsel, "ops") == 0) {
        qwen_cuda_set_path(ctx, 1);
    }
    mp->ctx = ctx;
    /* the reference's banners (src/model.c:277, 399), with what THIS process put on its device: under tensor parallelism a
     * rank holds 1/tp of the matrices and of the KV cache */
    size_t wb = 0, kvb = 0, rd = 0;
    qwen_cuda_memory(ctx, &wb, &kvb, &rd);
    fprintf(stderr, "[Weights] Uploaded %.2f MB to device\n", (double) wb / (1024.0 * 1024.0));
    fprintf(stderr, "[ForwardState] Allocated %.2f MB on device\n", (double) kvb / (1024.0 * 1024.0));
    (void) K;
    return m;

fail_state:
    free(s->x); free(s->x_rms_norm); free(s->q); free(s->mlp_in); free(s->mlp_gate);
    free(s->qx.q); free(s->qx.s); free(s->qh.q); free(s->qh.s);
    qwen_cuda_host_free(s->logits);
fail_views:
    free_views(w, p->shared_classifier);
fail_map:
    munmap(map, (size_t) st.st_size);
    free(mp);
    return NULL;
}

void model_free(Model* m) {
    if (!m) {
        return;
    }
    ModelPriv* mp = (ModelPriv*) m;
    qwen_cuda_destroy(mp->ctx);
    ForwardState* s = &m->state;
    free(s->x); free(s->x_rms_norm); free(s->q); free(s->mlp_in); free(s->mlp_gate);
    free(s->qx.q); free(s->qx.s); free(s->qh.q); free(s->qh.s);
    qwen_cuda_host_free(s->logits);
    free_views(&m->weights, m->params.shared_classifier);
    munmap(m->data, (size_t) m->size);
    free(mp);
}

/* Extension used by bench.py / tests: the device context behind a Model. */
QwenCudaCtx* model_cuda_ctx(Model* m) {
    return m ? ((ModelPriv*) m)->ctx : NULL;
}
