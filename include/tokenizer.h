/*
 * tokenizer.h -- text <-> token ids, B200 build.
 *
 * ABI mirror of the reference header (reference: include/tokenizer.h:24-148): same structs, field order and five
 * prototypes, so src/qwen.c:21, src/completion.c:47,69 and examples/tokenizer.c compile against it unchanged.
 * The implementation (qwen3.c_b200/csrc/host_tokenizer.c) answers tokenizer_token_to_id from a hash table and keeps the
 * merge candidates of tokenizer_encode incrementally, where the reference scans the whole vocabulary with strcmp for
 * every lookup (src/tokenizer.c:150-168: O(T^2 * V) per prompt -- minutes for a 2 k-character prompt before the GPU
 * starts). Results are identical id for id (tests/test_tokenizer_cpu.py against the reference's compiled tokenizer.c).
 * Host code only: nothing here touches the device.
 */
#ifndef QWEN_TOKENIZER_H
#define QWEN_TOKENIZER_H

#ifdef __cplusplus
extern "C" {
#endif

#define QTKN_MAGIC 0x71746B6E   /* "qtkn" */
#define QTKN_VERSION 2
#define QTKN_VOCAB_SIZE 151936
#define QTKN_MAX_SEQ_LEN 32768

typedef struct TokenEntry {
    char* token; /* null-terminated UTF-8 */
    float score; /* merge rank score, higher is better */
} TokenEntry;

typedef struct TokenSpecial {
    int bos, eos, eot, pad; /* core ids */
    int bor, eor;           /* think */
    int btc, etc;           /* tool call */
    int btr, etr;           /* tool response */
} TokenSpecial;

typedef struct Tokenizer {
    TokenEntry* entries; /* id -> token */
    TokenSpecial special;
    int magic;
    int version;
    int vocab_size;
    int max_len; /* longest token in bytes */
} Tokenizer;

/* Loads "<prefix>.tokenizer" (reference: src/tokenizer.c:17-120). NULL on failure. The returned object is larger than
 * Tokenizer (the lookup table sits behind it): release it with tokenizer_free only. */
Tokenizer* tokenizer_create(const char* prefix);
void tokenizer_free(Tokenizer* t);
char* tokenizer_id_to_token(Tokenizer* t, int id);         /* src/tokenizer.c:141-147 */
int tokenizer_token_to_id(Tokenizer* t, const char* token); /* src/tokenizer.c:150-168: lowest id with that string, or -1 */
/* Bytes (and "<...>" specials) -> ids, then greedy merges by best score, leftmost on ties (src/tokenizer.c:176-287). */
void tokenizer_encode(Tokenizer* t, char* text, int* ids, int* n_ids);

#ifdef __cplusplus
}
#endif
#endif /* QWEN_TOKENIZER_H */
