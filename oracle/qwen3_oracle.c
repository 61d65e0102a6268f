/*
 * qwen3_oracle.c -- CPU restatement of qwen3.c's Q8_0 forward path.
 *
 * TEST INFRASTRUCTURE ONLY (see qwen3_oracle.h). IEEE fp32, every output element computed serially in the reference's
 * order (the independent output rows of orc_matmul may run on several host threads: bit-identical, pinned).
 * Build: gcc -std=gnu17 -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).
 * Every function names the reference lines it restates; arithmetic order and
 * association are kept exactly so results are bit-identical to the reference's
 * strict build (pinned by tests/test_oracle_pinned.py).
 */
#include "qwen3_oracle.h"

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

/* ------------------------------------------------------------------ */
/* Q8_0 primitives                                                      */
/* ------------------------------------------------------------------ */

/* reference: src/q8.c:5-30. Groups of gs; absmax -> scale = absmax/127 (true
 * division), all-zero group -> 1e-6f; code = clamp(roundf(x/scale)). The tail
 * n % gs is left untouched, as in the reference. */
void orc_q8_quantize(int8_t* q, float* s, const float* x, int n, int gs) {
    const int groups = n / gs;
    for (int g = 0; g < groups; ++g) {
        const float* xg = x + (size_t) g * gs;
        float amax = fabsf(xg[0]);
        for (int i = 1; i < gs; ++i) {
            amax = fmaxf(amax, fabsf(xg[i]));
        }
        const float scale = (amax == 0.0f) ? 1e-6f : (amax / 127.0f);
        s[g] = scale;
        for (int i = 0; i < gs; ++i) {
            float r = roundf(xg[i] / scale);
            r = fmaxf(r, -127.0f);
            r = fminf(r, 127.0f);
            q[(size_t) g * gs + i] = (int8_t) r;
        }
    }
}

/* reference: src/q8.c:32-37 */
void orc_q8_dequantize(float* x, const int8_t* q, const float* s, int n, int gs) {
    for (int i = 0; i < n; ++i) {
        x[i] = (float) q[i] * s[i / gs];
    }
}

/* ------------------------------------------------------------------ */
/* ops                                                                  */
/* ------------------------------------------------------------------ */

/* The integer half of reference src/forward.c:89-92, exposed on its own so the
 * GPU's per-group dots can be compared bit for bit. */
void orc_group_dots(int32_t* dots, const int8_t* xq, const int8_t* wq, int n, int d, int gs) {
    const int groups = n / gs;
    /* Output rows are independent: they may run on several host threads (the checker must finish the 1.7B / 4B
     * shapes in seconds per token). Inside a row everything is the reference's serial arithmetic, so the result is
     * bit-identical to the single-thread run whatever the thread count (pinned in tests/test_oracle_pinned.py). */
#pragma omp parallel for schedule(static) if ((size_t) d * (size_t) n > (size_t) 1 << 18)
    for (int i = 0; i < d; ++i) {
        const int8_t* row = wq + (size_t) i * n;
        for (int g = 0; g < groups; ++g) {
            int32_t acc = 0;
            for (int k = 0; k < gs; ++k) {
                acc += (int32_t) xq[g * gs + k] * (int32_t) row[g * gs + k];
            }
            dots[(size_t) i * groups + g] = acc;
        }
    }
}

/* reference: src/forward.c:79-101. Per row, groups are folded left to right
 * into an fp32 accumulator that starts at 0; each term is
 * ((float) dot * w_scale) * x_scale in that association. */
void orc_matmul(float* out, const int8_t* xq, const float* xs,
                const int8_t* wq, const float* ws, int n, int d, int gs) {
    const int groups = n / gs;
    for (int i = 0; i < d; ++i) {
        const int8_t* row = wq + (size_t) i * n;
        const float* rs = ws + (size_t) i * groups;
        float acc = 0.0f;
        for (int g = 0; g < groups; ++g) {
            int32_t dot = 0;
            for (int k = 0; k < gs; ++k) {
                dot += (int32_t) xq[g * gs + k] * (int32_t) row[g * gs + k];
            }
            float term = (float) dot * rs[g];
            term = term * xs[g];
            acc = acc + term;
        }
        out[i] = acc;
    }
}

/* reference: src/forward.c:12-28. Sequential sum of squares, then
 * r = 1/sqrt(ss/size + 1e-6), out = w * (r * x). */
void orc_rmsnorm(float* out, const float* x, const float* w, int size) {
    float ss = 0.0f;
    for (int i = 0; i < size; ++i) {
        ss += x[i] * x[i];
    }
    const float r = 1.0f / sqrtf((ss / (float) size) + 1e-6f);
    for (int i = 0; i < size; ++i) {
        out[i] = w[i] * (r * x[i]);
    }
}

/* reference: src/forward.c:34-77 (diagnostic prints omitted). */
void orc_softmax(float* x, int size) {
    float mx = x[0];
    for (int i = 1; i < size; ++i) {
        if (x[i] > mx) {
            mx = x[i];
        }
    }
    float sum = 0.0f;
    for (int i = 0; i < size; ++i) {
        x[i] = expf(x[i] - mx);
        sum += x[i];
    }
    for (int i = 0; i < size; ++i) {
        x[i] /= sum;
    }
}

/* reference: src/forward.c:104-118. Pairs (i, i+half), theta 1e6, angle formed
 * in fp32 as (float) pos * powf(1e6f, -(float) i / half). */
void orc_rotary(float* x, int head_dim, int pos) {
    const int half = head_dim / 2;
    for (int i = 0; i < half; ++i) {
        const float angle = (float) pos * powf(1e6f, -(float) i / (float) half);
        const float c = cosf(angle), sn = sinf(angle);
        const float a = x[i], b = x[i + half];
        x[i] = a * c - b * sn;
        x[i + half] = a * sn + b * c;
    }
}

/* reference: src/forward.c:122-124 */
float orc_sigmoid(float x) {
    return 1.0f / (1.0f + expf(-x));
}

/* reference: src/forward.c:127-129 */
float orc_silu(float x) {
    return x * orc_sigmoid(x);
}

/* reference: src/forward.c:134-139 */
void orc_swiglu(float* x1, const float* x3, int size) {
    for (int i = 0; i < size; ++i) {
        x1[i] = orc_silu(x1[i]) * x3[i];
    }
}

/* reference: src/forward.c:141-195, single-threaded reading: per head, scores
 * 0..pos = (q . k_i) / sqrtf(head_dim); softmax; out = sum_i p_i * v_i folded
 * in position order from 0. */
void orc_attention(float* out, const float* q, const float* k_layer, const float* v_layer,
                   float* scores, int n_heads, int n_kv_heads, int head_dim,
                   int seq_len, int pos) {
    const int kv_mul = n_heads / n_kv_heads;
    const int kv_dim = n_kv_heads * head_dim;
    const float denom = sqrtf((float) head_dim);
    for (int h = 0; h < n_heads; ++h) {
        const float* qh = q + (size_t) h * head_dim;
        float* sc = scores + (size_t) h * seq_len;
        float* oh = out + (size_t) h * head_dim;
        const size_t hoff = (size_t) (h / kv_mul) * head_dim;
        for (int i = 0; i <= pos; ++i) {
            const float* k = k_layer + (size_t) i * kv_dim + hoff;
            float dot = 0.0f;
            for (int j = 0; j < head_dim; ++j) {
                dot += qh[j] * k[j];
            }
            sc[i] = dot / denom;
        }
        orc_softmax(sc, pos + 1);
        for (int j = 0; j < head_dim; ++j) {
            oh[j] = 0.0f;
        }
        for (int i = 0; i <= pos; ++i) {
            const float* v = v_layer + (size_t) i * kv_dim + hoff;
            for (int j = 0; j < head_dim; ++j) {
                oh[j] += sc[i] * v[j];
            }
        }
        /* the reference adds its (single) thread-local buffer to a zeroed
         * output: 0.0f + t == t, so nothing further to do */
    }
}

/* ------------------------------------------------------------------ */
/* checkpoint                                                           */
/* ------------------------------------------------------------------ */

/* Consume `count` Q8 tensors of `numel` codes each from the cursor:
 * int8[numel] then float[numel/gs] (reference: src/model.c:120-142). */
static OrcQ8* take_q8(const uint8_t** cur, int count, size_t numel, int gs) {
    OrcQ8* t = (OrcQ8*) calloc((size_t) count, sizeof(OrcQ8));
    if (!t) {
        return NULL;
    }
    for (int i = 0; i < count; ++i) {
        t[i].q = (const int8_t*) *cur;
        *cur += numel;
        t[i].s = (const float*) *cur;
        *cur += (numel / (size_t) gs) * sizeof(float);
    }
    return t;
}

/* reference: src/model.c:19-97 (map + header), :162-244 (tensor order),
 * :321-377 (state). The reference's inverted block_size assert is not
 * reproduced (it is compiled out with -DNDEBUG in every usable build). */
OrcModel* orc_model_open(const char* path, int seq_len_override) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) {
        return NULL;
    }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 256) {
        close(fd);
        return NULL;
    }
    void* map = mmap(NULL, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) {
        return NULL;
    }
    int32_t hdr[12];
    memcpy(hdr, map, sizeof hdr);
    if (hdr[0] != 0x7177656E || hdr[1] != 1) {
        munmap(map, (size_t) st.st_size);
        return NULL;
    }
    OrcModel* m = (OrcModel*) calloc(1, sizeof(OrcModel));
    m->map = map;
    m->map_len = (size_t) st.st_size;
    m->dim = hdr[2];
    m->hidden_dim = hdr[3];
    m->n_layers = hdr[4];
    m->n_heads = hdr[5];
    m->n_kv_heads = hdr[6];
    m->vocab_size = hdr[7];
    m->seq_len = hdr[8];
    m->head_dim = hdr[9];
    m->shared_classifier = hdr[10];
    m->group_size = hdr[11];
    if (seq_len_override > 0 && seq_len_override <= m->seq_len) {
        m->seq_len = seq_len_override;
    }
    const int L = m->n_layers, D = m->dim, Hd = m->hidden_dim, gs = m->group_size;
    const int P = m->n_heads * m->head_dim, K = m->n_kv_heads * m->head_dim;

    const uint8_t* cur = (const uint8_t*) map + 256;
    const float* f = (const float*) cur;
    m->att_norm = f;  f += (size_t) L * D;
    m->ffn_norm = f;  f += (size_t) L * D;
    m->out_norm = f;  f += D;
    m->q_norm = f;    f += (size_t) L * m->head_dim;
    m->k_norm = f;    f += (size_t) L * m->head_dim;
    cur = (const uint8_t*) f;

    OrcQ8* e = take_q8(&cur, 1, (size_t) m->vocab_size * D, gs);
    m->emb = *e;
    free(e);
    m->wq = take_q8(&cur, L, (size_t) D * P, gs);
    m->wk = take_q8(&cur, L, (size_t) D * K, gs);
    m->wv = take_q8(&cur, L, (size_t) D * K, gs);
    m->wo = take_q8(&cur, L, (size_t) P * D, gs);
    m->w1 = take_q8(&cur, L, (size_t) D * Hd, gs);
    m->w2 = take_q8(&cur, L, (size_t) Hd * D, gs);
    m->w3 = take_q8(&cur, L, (size_t) D * Hd, gs);
    if (m->shared_classifier) {
        m->cls = m->emb;
    } else {
        OrcQ8* c = take_q8(&cur, 1, (size_t) m->vocab_size * D, gs);
        m->cls = *c;
        free(c);
    }
    if ((size_t) (cur - (const uint8_t*) map) > m->map_len) {
        fprintf(stderr, "[Oracle] checkpoint shorter than its header implies\n");
        orc_model_close(m);
        return NULL;
    }

    const size_t amax = (size_t) (P > Hd ? P : Hd) > (size_t) D ? (size_t) (P > Hd ? P : Hd) : (size_t) D;
    const size_t cache = (size_t) L * m->seq_len * K;
    m->x = (float*) calloc((size_t) D, sizeof(float));
    m->xb = (float*) calloc(amax, sizeof(float));
    m->q = (float*) calloc((size_t) P, sizeof(float));
    m->att = (float*) calloc((size_t) P, sizeof(float));
    m->h1 = (float*) calloc((size_t) Hd, sizeof(float));
    m->h3 = (float*) calloc((size_t) Hd, sizeof(float));
    m->scores = (float*) calloc((size_t) m->n_heads * m->seq_len, sizeof(float));
    m->logits = (float*) calloc((size_t) m->vocab_size, sizeof(float));
    m->k_cache = (float*) calloc(cache, sizeof(float));
    m->v_cache = (float*) calloc(cache, sizeof(float));
    m->aq = (int8_t*) calloc(amax, 1);
    m->as = (float*) calloc(amax / (size_t) gs + 1, sizeof(float));
    if (!m->x || !m->xb || !m->q || !m->att || !m->h1 || !m->h3 || !m->scores || !m->logits
        || !m->k_cache || !m->v_cache || !m->aq || !m->as) {
        orc_model_close(m);
        return NULL;
    }
    return m;
}

void orc_model_close(OrcModel* m) {
    if (!m) {
        return;
    }
    free(m->wq); free(m->wk); free(m->wv); free(m->wo);
    free(m->w1); free(m->w2); free(m->w3);
    free(m->x); free(m->xb); free(m->q); free(m->att); free(m->h1); free(m->h3);
    free(m->scores); free(m->logits); free(m->k_cache); free(m->v_cache);
    free(m->aq); free(m->as);
    free(m->tr_qkv_in_q); free(m->tr_qkv_in_s); free(m->tr_q_rot); free(m->tr_att_out);
    free(m->tr_x_after_att); free(m->tr_x_after_ffn); free(m->tr_h);
    free(m->tr_wo_in_q); free(m->tr_wo_in_s); free(m->tr_ffn_in_q); free(m->tr_ffn_in_s); free(m->tr_w2_in_q);
    free(m->tr_w2_in_s); free(m->tr_cls_in_q); free(m->tr_cls_in_s); free(m->tr_x_final); free(m->tr_x_normed);
    if (m->map) {
        munmap(m->map, m->map_len);
    }
    free(m);
}

int orc_trace_enable(OrcModel* m) {
    const size_t L = (size_t) m->n_layers, D = (size_t) m->dim, Hd = (size_t) m->hidden_dim;
    const size_t P = (size_t) m->n_heads * m->head_dim;
    m->tr_qkv_in_q = (int8_t*) calloc(L * D, 1);
    m->tr_qkv_in_s = (float*) calloc(L * (D / m->group_size), sizeof(float));
    m->tr_q_rot = (float*) calloc(L * P, sizeof(float));
    m->tr_att_out = (float*) calloc(L * P, sizeof(float));
    m->tr_x_after_att = (float*) calloc(L * D, sizeof(float));
    m->tr_x_after_ffn = (float*) calloc(L * D, sizeof(float));
    m->tr_h = (float*) calloc(L * Hd, sizeof(float));
    const size_t gs = (size_t) m->group_size;
    m->tr_wo_in_q = (int8_t*) calloc(L * P, 1);
    m->tr_wo_in_s = (float*) calloc(L * (P / gs), sizeof(float));
    m->tr_ffn_in_q = (int8_t*) calloc(L * D, 1);
    m->tr_ffn_in_s = (float*) calloc(L * (D / gs), sizeof(float));
    m->tr_w2_in_q = (int8_t*) calloc(L * Hd, 1);
    m->tr_w2_in_s = (float*) calloc(L * (Hd / gs), sizeof(float));
    m->tr_cls_in_q = (int8_t*) calloc(D, 1);
    m->tr_cls_in_s = (float*) calloc(D / gs, sizeof(float));
    m->tr_x_final = (float*) calloc(D, sizeof(float));
    m->tr_x_normed = (float*) calloc(D, sizeof(float));
    m->trace_on = m->tr_qkv_in_q && m->tr_qkv_in_s && m->tr_q_rot && m->tr_att_out
                  && m->tr_x_after_att && m->tr_x_after_ffn && m->tr_h && m->tr_wo_in_q && m->tr_wo_in_s
                  && m->tr_ffn_in_q && m->tr_ffn_in_s && m->tr_w2_in_q && m->tr_w2_in_s && m->tr_cls_in_q
                  && m->tr_cls_in_s && m->tr_x_final && m->tr_x_normed;
    return m->trace_on ? 0 : -1;
}

/* reference: src/forward.c:225-350, step for step. */
float* orc_forward(OrcModel* m, int token, int pos) {
    return orc_forward_ex(m, token, pos, 1);
}

/* want_logits == 0: everything but the classifier matmul (:348) -- for prompt positions, whose logits the reference's
 * loops compute and discard (src/completion.c:59-63); the KV cache and the residual stream are the same either way. */
float* orc_forward_ex(OrcModel* m, int token, int pos, int want_logits) {
    const int D = m->dim, Hd = m->hidden_dim, hd = m->head_dim, gs = m->group_size;
    const int P = m->n_heads * hd, K = m->n_kv_heads * hd;

    /* :237 -- embedding row, dequantised (the reference reads it from the fp32
     * table it built with q8_dequantize at load, model.c:206; same values) */
    orc_q8_dequantize(m->x, m->emb.q + (size_t) token * D, m->emb.s + ((size_t) token * D) / gs,
                      D, gs);
    /* note: i/gs inside dequantize restarts at 0 for the row; D % gs == 0 so the
     * row starts on a group boundary and the scales line up */

    for (int l = 0; l < m->n_layers; ++l) {
        float* k_layer = m->k_cache + (size_t) l * m->seq_len * K;
        float* v_layer = m->v_cache + (size_t) l * m->seq_len * K;
        float* k = k_layer + (size_t) pos * K; /* :244-248, written in place */
        float* v = v_layer + (size_t) pos * K;

        orc_rmsnorm(m->xb, m->x, m->att_norm + (size_t) l * D, D);       /* :254 */
        orc_q8_quantize(m->aq, m->as, m->xb, D, gs);                      /* :259 */
        if (m->trace_on) {
            memcpy(m->tr_qkv_in_q + (size_t) l * D, m->aq, (size_t) D);
            memcpy(m->tr_qkv_in_s + (size_t) l * (D / gs), m->as, sizeof(float) * (size_t) (D / gs));
        }
        orc_matmul(m->q, m->aq, m->as, m->wq[l].q, m->wq[l].s, D, P, gs); /* :260 */
        orc_matmul(k, m->aq, m->as, m->wk[l].q, m->wk[l].s, D, K, gs);    /* :261 */
        orc_matmul(v, m->aq, m->as, m->wv[l].q, m->wv[l].s, D, K, gs);    /* :262 */

        const float* gq = m->q_norm + (size_t) l * hd;                    /* :267-268 */
        const float* gk = m->k_norm + (size_t) l * hd;
        for (int h = 0; h < m->n_heads; ++h) {                            /* :270-274 */
            float* qh = m->q + (size_t) h * hd;
            orc_rmsnorm(qh, qh, gq, hd);
            orc_rotary(qh, hd, pos);
        }
        for (int h = 0; h < m->n_kv_heads; ++h) {                         /* :276-280 */
            float* kh = k + (size_t) h * hd;
            orc_rmsnorm(kh, kh, gk, hd);
            orc_rotary(kh, hd, pos);
        }
        if (m->trace_on) {
            memcpy(m->tr_q_rot + (size_t) l * P, m->q, sizeof(float) * (size_t) P);
        }

        orc_attention(m->xb, m->q, k_layer, v_layer, m->scores, m->n_heads, m->n_kv_heads, hd,
                      m->seq_len, pos);                                   /* :286 */
        if (m->trace_on) {
            memcpy(m->tr_att_out + (size_t) l * P, m->xb, sizeof(float) * (size_t) P);
        }

        orc_q8_quantize(m->aq, m->as, m->xb, P, gs);                      /* :291 */
        if (m->trace_on) {
            memcpy(m->tr_wo_in_q + (size_t) l * P, m->aq, (size_t) P);
            memcpy(m->tr_wo_in_s + (size_t) l * (P / gs), m->as, sizeof(float) * (size_t) (P / gs));
        }
        orc_matmul(m->xb, m->aq, m->as, m->wo[l].q, m->wo[l].s, P, D, gs);/* :292-294 */
        for (int i = 0; i < D; ++i) {                                     /* :295-298 */
            m->x[i] += m->xb[i];
        }
        if (m->trace_on) {
            memcpy(m->tr_x_after_att + (size_t) l * D, m->x, sizeof(float) * (size_t) D);
        }

        orc_rmsnorm(m->xb, m->x, m->ffn_norm + (size_t) l * D, D);        /* :303 */
        orc_q8_quantize(m->aq, m->as, m->xb, D, gs);                      /* :308 */
        if (m->trace_on) {
            memcpy(m->tr_ffn_in_q + (size_t) l * D, m->aq, (size_t) D);
            memcpy(m->tr_ffn_in_s + (size_t) l * (D / gs), m->as, sizeof(float) * (size_t) (D / gs));
        }
        orc_matmul(m->h1, m->aq, m->as, m->w1[l].q, m->w1[l].s, D, Hd, gs); /* :309-311 */
        orc_matmul(m->h3, m->aq, m->as, m->w3[l].q, m->w3[l].s, D, Hd, gs); /* :312-314 */
        orc_swiglu(m->h1, m->h3, Hd);                                     /* :319-321 */
        if (m->trace_on) {
            memcpy(m->tr_h + (size_t) l * Hd, m->h1, sizeof(float) * (size_t) Hd);
        }
        orc_q8_quantize(m->aq, m->as, m->h1, Hd, gs);                     /* :326 */
        if (m->trace_on) {
            memcpy(m->tr_w2_in_q + (size_t) l * Hd, m->aq, (size_t) Hd);
            memcpy(m->tr_w2_in_s + (size_t) l * (Hd / gs), m->as, sizeof(float) * (size_t) (Hd / gs));
        }
        orc_matmul(m->xb, m->aq, m->as, m->w2[l].q, m->w2[l].s, Hd, D, gs); /* :327-334 */
        for (int i = 0; i < D; ++i) {                                     /* :335-338 */
            m->x[i] += m->xb[i];
        }
        if (m->trace_on) {
            memcpy(m->tr_x_after_ffn + (size_t) l * D, m->x, sizeof(float) * (size_t) D);
        }
    }

    if (m->trace_on) {
        memcpy(m->tr_x_final, m->x, sizeof(float) * (size_t) D);
    }
    orc_rmsnorm(m->x, m->x, m->out_norm, D);                              /* :344 */
    orc_q8_quantize(m->aq, m->as, m->x, D, gs);                           /* :347 */
    if (m->trace_on) {
        memcpy(m->tr_x_normed, m->x, sizeof(float) * (size_t) D);
        memcpy(m->tr_cls_in_q, m->aq, (size_t) D);
        memcpy(m->tr_cls_in_s, m->as, sizeof(float) * (size_t) (D / gs));
    }
    if (want_logits) {
        orc_matmul(m->logits, m->aq, m->as, m->cls.q, m->cls.s, D, m->vocab_size, gs); /* :348 */
    }
    return m->logits;
}

int orc_argmax(const float* v, int n, float* margin) {
    int best = 0;
    float top = v[0], second = -INFINITY;
    for (int i = 1; i < n; ++i) {
        if (v[i] > top) {
            second = top;
            top = v[i];
            best = i;
        } else if (v[i] > second) {
            second = v[i];
        }
    }
    if (margin) {
        *margin = top - second;
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * Sampler: restatement of reference src/sampler.c and src/xorshift.c. Test
 * infrastructure for the device sampler (qwen_cuda_sample); never part of the product path.
 * ---------------------------------------------------------------------------------------------- */

/* reference src/xorshift.c:7-12 */
uint32_t orc_xorshift_int32(uint64_t* state) {
    *state ^= *state >> 12;
    *state ^= *state << 25;
    *state ^= *state >> 27;
    return (uint32_t) ((*state * 0x2545F4914F6CDD1Dull) >> 32);
}
/* reference src/xorshift.c:14-16: 24 random bits -> [0, 1) */
float orc_xorshift_float(uint64_t* state) { return (float) (orc_xorshift_int32(state) >> 8) / 16777216.0f; }

/* reference src/sampler.c:34-52: the clamps sampler_create applies to top_p and temperature */
void orc_sampler_clamp(float* temperature, float* top_p) {
    const float epsilon = 1e-6f;
    float p = *top_p, t = *temperature;
    if (p > 1.0f || isnan(p) || 1 == isinf(p)) p = 1.0f;
    else if (p < epsilon || -1 == isinf(p)) p = epsilon;
    if (isnan(t) || 1 == isinf(t)) t = 1.0f;
    else if (t < epsilon || -1 == isinf(t)) t = epsilon;
    *top_p = p;
    *temperature = t;
}

typedef struct { float sample; int index; } OrcProb; /* reference include/sampler.h:15-18 */

/* reference src/sampler.c:139-149 (descending by probability; equal elements compare equal) */
static int orc_cmp_dist(const void* a, const void* b) {
    const OrcProb* n = (const OrcProb*) a;
    const OrcProb* m = (const OrcProb*) b;
    if (n->sample > m->sample) return -1;
    if (n->sample < m->sample) return 1;
    return 0;
}

/* reference src/sampler.c:186-201 (sample) with :164-178 (sampler_top_p), :88-113 (sampler_mass_index) and
 * :126-136 (sampler_cdf_index) inlined. logits[] is overwritten with the probabilities like the reference does.
 * temperature / top_p are the already clamped values. `gap` (optional) receives how far r = coin * mass is from the
 * nearest cdf boundary it was compared with, and how far the cut-off mass is from top_p -- the smaller of the two,
 * relative: a device sampler whose probabilities differ in the last bits may legitimately pick a neighbour when
 * the gap is ~1e-6. Returns the token, or -1 if the scratch allocation fails. */
int orc_sample(float* logits, int vocab_size, float temperature, float top_p, float coin, float* gap) {
    for (int q = 0; q < vocab_size; q++) logits[q] /= temperature;
    orc_softmax(logits, vocab_size);
    OrcProb* dist = (OrcProb*) malloc((size_t) vocab_size * sizeof(OrcProb));
    if (!dist) return -1;
    for (int i = 0; i < vocab_size; i++) {
        dist[i].index = i;
        dist[i].sample = logits[i];
    }
    qsort(dist, (size_t) vocab_size, sizeof(OrcProb), orc_cmp_dist);
    /* sampler_mass_index */
    float mass = 0.0f, g = INFINITY;
    int id = vocab_size - 1;
    for (int i = 0; i < vocab_size; i++) {
        const float before = mass;
        mass += dist[i].sample;
        if (mass > top_p) {
            id = i;
            g = fminf(fabsf(mass - top_p), fabsf(top_p - before)) / fmaxf(top_p, 1e-30f);
            break;
        }
    }
    const float epsilon = 1e-3f;
    if (mass < epsilon) {
        for (int i = 0; i <= id; i++) mass += dist[i].sample;
    }
    /* sampler_cdf_index(dist, n = id, coin, mass): note the inclusive bound i <= n and the fallback dist[n - 1] */
    float cdf = 0.0f;
    const float r = coin * mass;
    int tok = -2;
    for (int i = 0; i <= id; i++) {
        cdf += dist[i].sample;
        if (r < cdf) {
            tok = dist[i].index;
            const float lo = cdf - dist[i].sample;
            g = fminf(g, fminf(fabsf(cdf - r), fabsf(r - lo)) / fmaxf(mass, 1e-30f));
            break;
        }
    }
    if (tok == -2) tok = dist[id > 0 ? id - 1 : 0].index; /* reference reads dist[n - 1]; n = 0 would read dist[-1] */
    if (gap) *gap = g;
    free(dist);
    return tok;
}
