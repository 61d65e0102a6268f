/* host_tokenizer.c -- the reference's tokenizer API (include/tokenizer.h) with a hash-table vocabulary lookup.
 *
 * SURVEY.md 8f-4. The reference answers every token -> id question by strcmp over the whole vocabulary
 * (src/tokenizer.c:150-168) and asks it once per adjacent pair per merge round (src/tokenizer.c:229-281): O(T^2 * V)
 * string compares per prompt. Here: an open-addressing table built once at load (FNV-1a over the token bytes, lowest id
 * wins for duplicate strings, exactly what the linear scan returns), and an array of merge candidates -- the id of
 * token[i] + token[i+1] if that string is in the vocabulary -- of which only the two entries next to a merge are looked up
 * again. A round is then a scan of floats; the ids produced are the reference's, including its tie rule (first candidate
 * with the strictly greatest score wins). Host code only.
 */
#include "../../include/tokenizer.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct TokPriv {
    Tokenizer pub;   /* must stay first: callers hold a Tokenizer* */
    int* table;      /* open addressing, -1 = empty, else token id */
    uint32_t mask;   /* table size - 1 (power of two) */
} TokPriv;

static uint32_t tok_hash(const char* s, size_t n) {
    uint32_t h = 2166136261u;
    for (size_t i = 0; i < n; ++i) {
        h = (h ^ (uint8_t) s[i]) * 16777619u;
    }
    return h;
}

static int tok_lookup(const TokPriv* p, const char* s, size_t n) {
    for (uint32_t i = tok_hash(s, n) & p->mask;; i = (i + 1) & p->mask) {
        const int id = p->table[i];
        if (id < 0) {
            return -1;
        }
        const char* t = p->pub.entries[id].token;
        if (strlen(t) == n && memcmp(t, s, n) == 0) {
            return id;
        }
    }
}

static int tok_build_table(TokPriv* p) {
    uint32_t size = 1024;
    while (size < 2u * (uint32_t) p->pub.vocab_size) {
        size <<= 1;
    }
    p->table = (int*) malloc(sizeof(int) * size);
    if (!p->table) {
        return -1;
    }
    for (uint32_t i = 0; i < size; ++i) {
        p->table[i] = -1;
    }
    p->mask = size - 1;
    for (int id = 0; id < p->pub.vocab_size; ++id) { /* ascending ids: a duplicate string keeps its LOWEST id, as strcmp scanning does */
        const char* t = p->pub.entries[id].token;
        const size_t n = strlen(t);
        uint32_t i = tok_hash(t, n) & p->mask;
        int dup = 0;
        while (p->table[i] >= 0) {
            const char* o = p->pub.entries[p->table[i]].token;
            if (strlen(o) == n && memcmp(o, t, n) == 0) {
                dup = 1;
                break;
            }
            i = (i + 1) & p->mask;
        }
        if (!dup) {
            p->table[i] = id;
        }
    }
    return 0;
}

static void tok_release(TokPriv* p, int n_loaded) {
    for (int k = 0; k < n_loaded; ++k) {
        free(p->pub.entries[k].token);
    }
    free(p->pub.entries);
    free(p->table);
    free(p);
}

/* reference: src/tokenizer.c:17-120 -- same file format (appendix A of SURVEY.md), same banners, NULL on any failure */
Tokenizer* tokenizer_create(const char* prefix) {
    if (!prefix) {
        return NULL;
    }
    const char* suffix = ".tokenizer";
    char* path = (char*) calloc(strlen(prefix) + strlen(suffix) + 1, 1);
    if (!path) {
        return NULL;
    }
    strcpy(path, prefix);
    strcat(path, suffix);
    FILE* f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "[Tokenizer] Failed to open %s\n", path);
        free(path);
        return NULL;
    }
    free(path);
    TokPriv* p = (TokPriv*) calloc(1, sizeof(TokPriv));
    if (!p) {
        fclose(f);
        return NULL;
    }
    Tokenizer* t = &p->pub;
    if (fread(&t->magic, sizeof(uint32_t), 1, f) != 1 || fread(&t->version, sizeof(int32_t), 1, f) != 1
        || (uint32_t) t->magic != QTKN_MAGIC || t->version != QTKN_VERSION) {
        fprintf(stderr, "[Tokenizer] Invalid tokenizer format.\n");
        fclose(f);
        free(p);
        return NULL;
    }
    if (fread(&t->vocab_size, sizeof(int32_t), 1, f) != 1 || fread(&t->max_len, sizeof(int32_t), 1, f) != 1
        || fread(&t->special, sizeof(TokenSpecial), 1, f) != 1 || t->vocab_size <= 0 || t->max_len <= 0) {
        fprintf(stderr, "[Tokenizer] Invalid tokenizer header.\n");
        fclose(f);
        free(p);
        return NULL;
    }
    t->entries = (TokenEntry*) calloc((size_t) t->vocab_size, sizeof(TokenEntry));
    if (!t->entries) {
        fclose(f);
        free(p);
        return NULL;
    }
    for (int i = 0; i < t->vocab_size; ++i) {
        float score;
        int length;
        char* buf = NULL;
        if (fread(&score, sizeof(float), 1, f) != 1 || fread(&length, sizeof(int), 1, f) != 1 || length < 0
            || !(buf = (char*) calloc((size_t) length + 1, 1)) || fread(buf, 1, (size_t) length, f) != (size_t) length) {
            fprintf(stderr, "[Tokenizer] Token read error at index %d\n", i);
            free(buf);
            fclose(f);
            tok_release(p, i);
            return NULL;
        }
        t->entries[i].score = score;
        t->entries[i].token = buf;
    }
    fclose(f);
    if (tok_build_table(p)) {
        tok_release(p, t->vocab_size);
        return NULL;
    }
    fprintf(stderr, "[Tokenizer] magic=%x\n", t->magic);
    fprintf(stderr, "[Tokenizer] version=%d\n", t->version);
    fprintf(stderr, "[Tokenizer] vocab_size=%d\n", t->vocab_size);
    fprintf(stderr, "[Tokenizer] max_len=%d\n", t->max_len);
    fprintf(stderr, "[Tokenizer] bos=%d\n", t->special.bos);
    fprintf(stderr, "[Tokenizer] eos=%d\n", t->special.eos);
    fprintf(stderr, "[Tokenizer] eot=%d\n", t->special.eot);
    return t;
}

void tokenizer_free(Tokenizer* t) {
    if (t) {
        tok_release((TokPriv*) t, t->vocab_size);
    }
}

/* reference: src/tokenizer.c:141-147 */
char* tokenizer_id_to_token(Tokenizer* t, int id) {
    if (!t || !t->entries || id < 0 || id >= t->vocab_size) {
        fprintf(stderr, "[Tokenizer] ERROR: Invalid id! %d\n", id);
        return NULL;
    }
    return t->entries[id].token;
}

/* reference: src/tokenizer.c:150-168 */
int tokenizer_token_to_id(Tokenizer* t, const char* token) {
    if (!t || !t->entries || !token) {
        fprintf(stderr, "[Tokenizer] ERROR: Invalid token! %s\n", token ? token : "(null)");
        return -1;
    }
    return tok_lookup((const TokPriv*) t, token, strlen(token));
}

/* id of the string entries[a].token + entries[b].token, or -1 (reference: src/tokenizer.c:238-247; the reference's buffer
 * holds 2 * max_len bytes, which two tokens always fit) */
static int tok_pair(const TokPriv* p, int a, int b, char* buf) {
    const char* sa = p->pub.entries[a].token;
    const char* sb = p->pub.entries[b].token;
    const size_t na = strlen(sa), nb = strlen(sb);
    if (na + nb > (size_t) 2 * p->pub.max_len) { /* cannot happen for a well-formed file; the reference would truncate */
        return -1;
    }
    memcpy(buf, sa, na);
    memcpy(buf + na, sb, nb);
    return tok_lookup(p, buf, na + nb);
}

/* reference: src/tokenizer.c:176-287 */
void tokenizer_encode(Tokenizer* t, char* text, int* ids, int* n_ids) {
    *n_ids = 0;
    if (!t || !t->entries || !text || !ids) {
        return;
    }
    const TokPriv* p = (const TokPriv*) t;
    int n = 0;
    /* bytes and "<...>" specials -> ids (src/tokenizer.c:188-227) */
    for (char* bytes = text; *bytes;) {
        int id = -1;
        if (*bytes == '<') { /* up to max_len bytes ending in '>' may be one special token */
            for (int k = 0; bytes[k] && k < t->max_len; ++k) {
                if (bytes[k] == '>') {
                    id = tok_lookup(p, bytes, (size_t) k + 1);
                    if (id != -1) {
                        bytes += k + 1;
                    }
                    break;
                }
            }
        }
        if (id == -1) {
            const char c = *bytes++;
            id = tok_lookup(p, &c, 1);
            if (id == -1) {
                fprintf(stderr, "[Tokenizer] Warning: Unknown character `%c` (codepoint %d)\n", c, c);
            }
        }
        if (id != -1) {
            ids[n++] = id;
        }
    }
    /* greedy merges: every round takes the adjacent pair whose concatenation has the strictly greatest score, leftmost
     * first (src/tokenizer.c:229-281). pair[i] = id of ids[i] + ids[i + 1] or -1; a merge at i changes pair[i - 1] and
     * pair[i] only. */
    if (n >= 2) {
        int* pair = (int*) malloc(sizeof(int) * (size_t) n);
        char* buf = (char*) malloc((size_t) 2 * t->max_len + 1);
        if (pair && buf) {
            for (int i = 0; i + 1 < n; ++i) {
                pair[i] = tok_pair(p, ids[i], ids[i + 1], buf);
            }
            for (;;) {
                float best_score = -1e10f;
                int best = -1;
                for (int i = 0; i + 1 < n; ++i) {
                    if (pair[i] != -1 && t->entries[pair[i]].score > best_score) {
                        best_score = t->entries[pair[i]].score;
                        best = i;
                    }
                }
                if (best == -1) {
                    break;
                }
                ids[best] = pair[best];
                memmove(&ids[best + 1], &ids[best + 2], sizeof(int) * (size_t) (n - best - 2));
                memmove(&pair[best + 1], &pair[best + 2], sizeof(int) * (size_t) (n - best - 2));
                --n;
                if (best > 0) {
                    pair[best - 1] = tok_pair(p, ids[best - 1], ids[best], buf);
                }
                if (best + 1 < n) {
                    pair[best] = tok_pair(p, ids[best], ids[best + 1], buf);
                }
            }
        }
        free(pair);
        free(buf);
    }
    *n_ids = n;
}
