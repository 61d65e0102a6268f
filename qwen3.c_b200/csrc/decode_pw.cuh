// decode_pw.cuh -- the decode step with PER-WARP weight / KV pipelines (k_decode_pw). Included by decode_mega.cu inside
// its anonymous namespace: the flow arena, the hand-off polling, the prologues (RMSNorm + Q8_0 quantise), the combine and
// the host side are shared with the ring variant (k_decode), which stays the kernel of tensor-parallel contexts.
//
// Why: in the ring variant one producer thread fills 28 KB slots that all 15 consumer warps share. Measured on B200
// (profiles/r2_k_decode_summary.md): a warp owns a unit in one tile of three but has to VISIT every tile -- wait, __syncwarp,
// arrive, ~400 cycles each, 28 % of the w1/w3 phase -- because an mbarrier phase can only be told from its neighbour by
// parity, so a warp that skipped tiles would alias phases (tried: it reads tiles that have not landed). The attention tiles
// of a group of four fill the whole ring, so the next group's KV rows are requested only when the group is done, and a w2
// row pair fills a slot, so only `ring depth` warps have work at a time.
//
// Here every warp is its own pipeline: it owns 12 KB of shared memory and three mbarriers that nobody else waits on
// and walks only its own work. Weight copies are issued FOR it by the issue server (warp 15, see PwCtl below); KV copies
// by one elected lane of the warp itself:
//   * GEMV phases: work unit = 2 rows; a row is cut into chunks of <= 11 records (2992 B) so that two stages of
//     [row a chunk | row b chunk] fit the region; the lane partial sums ride across a unit's chunks. After consuming a
//     stage the warp publishes a counter and the server refills it with the warp's NEXT item -- across phase boundaries: weights do not depend on activations, so a
//     warp waiting at a hand-off already has its first two items of the next matrix in flight (15 x 12 KB per SM).
//   * attention: the warps of a head group split the CTA's cached positions to the position; a warp streams blocks of 8
//     positions as [8 K rows] and [8 V rows] items through three 4 KB sub-stages; the K rows' sub-stage is handed back
//     right after the scores, before the V rows are needed. Its first items are issued when its last QKV unit is done,
//     i.e. they travel during the q hand-off.
//   * the CTA merges the 15 warp states with three barriers (dump, weights, weighted sum) instead of a four-level tree.
// The issue cursor runs at most two GEMV items ahead of consumption and is held at the attention block (its shared memory
// is needed for KV rows) until the merge is done.

constexpr int kPwRegion = 12288;  // bytes of shared memory per consumer warp
constexpr int kPwStage = 6144;    // GEMV stage: [row a chunk <= 3072 B][row b chunk <= 3072 B]; two stages per region
constexpr int kPwSub = 4096;      // attention sub-stage: 8 K rows or 8 V rows of 512 B; three per region
constexpr int kPwChunkRecs = 11;  // records per row chunk: 11 * 272 = 2992 B <= kPwStage / 2
constexpr int kPwBlock = 8;       // positions per attention block (32 scores = 8 positions x 4 heads per butterfly)

// per CTA, filled once per launch in shared memory: the divisions of the work split happen once, not per phase
struct PwTables {
    int r0[5], r1[5];                               // this CTA's row range of every matrix
    int a_kvh, a_j, a_n, a_lo, a_hi, a_active;      // attention split (attn_split)
};

// Consumer warps do not issue their own weight copies: measured (scripts/ubench/tmaissue.cu, unit profile) an item costs the
// issuing warp ~50 cycles for mbarrier.expect_tx plus ~135-150 per cp.async.bulk -- with the address arithmetic ~680 cycles
// per item on a warp's critical path, half of what the dot products of the item take. Warp 15 is the ISSUE SERVER instead:
// its lane w serves consumer warp w. The consumer only publishes how many stages it has consumed (one shared-memory store),
// the server lane keeps the warp's cursor in its own registers and refills a stage as soon as it is free -- at most two
// items ahead, across phase boundaries, held at the attention block (the region then carries KV rows) until the consumer
// reports its merge done. Lane 15 of the server is the HBM -> L2 prefetcher (see pw_server).
struct PwCtl {
    volatile unsigned freed[16];  // per consumer warp: GEMV stages consumed so far (monotonic)
    volatile int unblocked[16];   // per consumer warp: last layer whose attention block it has finished (-1: none)
};

// warp 15. Lane w < 15: cursor of consumer warp w over its GEMV items -- phase ph (4 * layer + {0 QKV, 1 WO, 2 W13, 3 W2};
// 4 * layers = classifier), units u = w, w + 15, ... of the CTA's row range, chunks of a unit -- item n goes to stage n & 1.
// Lane 15: asks for everything the CTA reads in sub-phase sp (5 * layer + {0 QKV, 1 KV rows, 2 WO, 3 W1/W3, 4 W2};
// 5 * layers = the first part of the classifier) to be brought from HBM into L2 `l2_ahead` sub-phases before the consumers
// get there (cp.async.bulk.prefetch.L2). Measured on B200 (scripts/ubench/l2rate.cu): bulk copies into shared memory issued
// by 15 warps run at 16-20 TB/s when their source is in L2 against ~7 TB/s from HBM, and an HBM -> L2 prefetch stream barely
// slows the hand-off polls (scripts/ubench/handoff.cu): HBM keeps streaming through the hand-offs.
__device__ void pw_server(const Shared& sh, const MegaParams& p, const PwTables& tb, PwCtl* ctl, volatile int* progress, volatile int* abort_flag) {
    const int lane = threadIdx.x & 31;
    const int nph = 4 * p.layers_run, nsp = 5 * p.layers_run;
    int ph = 4 * p.l_begin, c = 0, C = 1, units_left = 0, rows_left = 0, rb = 0, cb = 0, cb_last = 0;
    unsigned n = 0;
    bool live = false, done = lane >= kConsumerWarps;
    const uint8_t* src = nullptr;
    int sp = 5 * p.l_begin;
    bool pf_done = !(lane == kConsumerWarps && p.l2_ahead > 0);
#pragma unroll 1
    for (;;) {
        if (*abort_flag) return;
        bool did = false;
        if (!done) {
#pragma unroll 1
            while (!live) { // enter the next phase in which the warp has units
                if (ph > nph) {
                    done = true;
                    break;
                }
                const int k = ph == nph ? 4 : (ph & 3);
                if (k == 1 && ctl->unblocked[lane] < (ph >> 2)) break; // its region still carries KV rows
                const MatDesc& m = p.mat[k];
                const int r1 = tb.r1[k];
                const int row = tb.r0[k] + 2 * lane;
                if (row >= r1) { // no units here
                    ++ph;
                    continue;
                }
                const int recs = qw_sg_per_row(m.n);
                C = m.C;
                c = 0;
                units_left = (r1 - row + 2 * kConsumerWarps - 1) / (2 * kConsumerWarps);
                rb = recs * QW_SG_BYTES;
                src = m.base + (k == 4 ? 0 : (size_t) (ph >> 2) * m.stride) + (size_t) row * rb;
                rows_left = r1 - row;
                cb = m.cr * QW_SG_BYTES;
                cb_last = (recs - (m.C - 1) * m.cr) * QW_SG_BYTES;
                live = true;
            }
            if (live && n - ctl->freed[lane] < 2u) {
                __threadfence_block(); // the consumer's reads of the stage happened before it published `freed`
                const bool last_chunk = c == C - 1;
                const uint32_t bytes = (uint32_t) (last_chunk ? cb_last : cb);
                const bool two = rows_left >= 2;
                const unsigned stage = n & 1u;
                const uint32_t bar = sh.full + (lane * 3 + stage) * 8;
                const uint32_t dst = smem_u32(sh.ring + (size_t) lane * kPwRegion + stage * kPwStage);
                mbar_expect_tx(bar, two ? 2 * bytes : bytes);
                if (C == 1 && two) { // whole rows: the pair is one contiguous range (one 5 KB copy runs at twice the rate of two 2.7 KB ones)
                    bulk_g2s(dst, src, 2 * bytes, bar);
                } else {
                    bulk_g2s(dst, src, bytes, bar);
                    if (two) bulk_g2s(dst + kPwStage / 2, src + rb, bytes, bar);
                }
                ++n;
                if (last_chunk) { // on to the warp's next unit, 15 units = 30 rows further
                    src += (size_t) 2 * kConsumerWarps * rb - (size_t) (C - 1) * cb;
                    rows_left -= 2 * kConsumerWarps;
                    c = 0;
                    if (--units_left == 0) {
                        live = false;
                        ++ph;
                    }
                } else {
                    src += cb;
                    ++c;
                }
                did = true;
            }
        }
        if (!pf_done) {
            if (sp > nsp) {
                pf_done = true;
            } else if (*progress + p.l2_ahead >= sp) {
                prefetch_subphase(p, sp);
                ++sp;
                did = true;
            }
        }
        if (__all_sync(0xffffffffu, done && pf_done)) return;
        if (!__any_sync(0xffffffffu, did)) __nanosleep(40);
    }
}

// wait for the warp's own barrier b (0..2); bpar holds the phase parity of each
__device__ __forceinline__ void pw_wait(const Shared& sh, const MegaParams& p, int warp, unsigned b, unsigned& bpar, int code) {
    mbar_wait(sh, p, sh.full + (warp * 3 + b) * 8, (bpar >> b) & 1u, code);
    bpar ^= 1u << b;
}

// one chunk (groups = 4 * records of the chunk) of two weight rows against the matching part of the activation vector;
// the same lane roles as the ring variant's consume_mat: full steps of 32 groups (lane = group), then the tail
__device__ __forceinline__ void pw_gemv_chunk(const uint8_t* xq, const uint8_t* rowa, const uint8_t* rowb, int groups, int lane, float& acca, float& accb) {
    const int rot = lane & 2;
    const int nfull = groups >> 5, rem = groups & 31;
    const bool tail_split = rem > 0 && rem <= 16;
    const bool t_rowb = tail_split && lane >= rem;
    const int tG = rem == 0 ? -1 : tail_split ? (lane < 2 * rem ? nfull * 32 + (t_rowb ? lane - rem : lane) : -1) : (lane < rem ? nfull * 32 + lane : -1);
#pragma unroll 1
    for (int k = 0; k <= nfull; ++k) { // full steps, then the tail when more than 16 lanes have a group in it
        const int G = k < nfull ? k * 32 + lane : (tail_split ? -1 : tG);
        if (G >= 0) gemv_step2(xq, rowa, rowb, G, rot, acca, accb);
    }
    if (tail_split && tG >= 0) { // a tail of <= 16 groups: lanes [0, rem) take row a, lanes [rem, 2 rem) row b
        const float t = gemv_step1(xq, t_rowb ? rowb : rowa, tG, rot);
        if (t_rowb) accb = __fadd_rn(accb, t); else acca = __fadd_rn(acca, t);
    }
}

// GEMV phase k of this CTA: units u = warp, warp + 15, ... of its row range, each streamed chunk by chunk through the warp's
// two stages. kind / staging / epilogue as in consume_mat (kinds 0, 1, 2; no tensor-parallel kind here).
__device__ __forceinline__ void pw_consume_mat(const Shared& sh, const MegaParams& p, const PwTables& tb, int k, PwCtl* ctl, unsigned& cons_n,
                                               unsigned& bpar, float* out, const float* resid) {
    const MatDesc& m = p.mat[k];
    const int r0 = tb.r0[k], r1 = tb.r1[k];
    const int KIND = m.kind;
    const int recs = qw_sg_per_row(m.n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nrows = r1 - r0;
    const int total = (nrows + 1) >> 1;
    const bool staged = m.stage != 0;
    float* stage = sh.scr;
    const uint8_t* region = sh.ring + (size_t) warp * kPwRegion;
    float pa0 = 0.0f, pb0 = 0.0f, xres = 0.0f; // lane partials of a unit whose reduction waits for the next unit
    int grow0 = 0;
    bool two0 = false, pend = false;
#pragma unroll 1
    for (int u = warp; u < total; u += kConsumerWarps) {
        const int grow = r0 + 2 * u;
        const bool two = grow + 1 < r1;
        if (KIND == 1 && (lane & 7) == 0 && ((lane & 16) != 0) == pend && ((lane & 8) == 0 || two))
            xres = __uint_as_float(ldf_u32(resid + grow + ((lane >> 3) & 1))); // issued early: hides the L2 round trip
        float acca = 0.0f, accb = 0.0f;
#pragma unroll 1
        for (int c = 0; c < m.C; ++c) {
            const unsigned s = cons_n & 1u;
            pw_wait(sh, p, warp, s, bpar, 3);
            const int rec0 = c * m.cr;
            const uint8_t* rowa = region + s * kPwStage;
            if (!(p.dbg_mode & 1)) // debug mode bit 0: skip the dot products (timing only)
                pw_gemv_chunk(sh.xq + (size_t) rec0 * QW_SG_BYTES, rowa, two ? rowa + (m.C == 1 ? recs * QW_SG_BYTES : kPwStage / 2) : rowa,
                              4 * min(m.cr, recs - rec0), lane, acca, accb);
            // every byte of the stage has been consumed by a dp4a: tell the issue server, it refills the stage with the warp's
            // next item (possibly of the next matrix)
            __syncwarp();
            ++cons_n;
            if (lane == 0) {
                __threadfence_block();
                ctl->freed[warp] = cons_n;
            }
        }
        // cross-lane reduction + epilogue for two units (4 rows) at a time: see consume_mat
        const bool last_unit = u + kConsumerWarps >= total;
        if (!pend && !last_unit) {
            pa0 = acca; pb0 = accb; grow0 = grow; two0 = two;
            pend = true;
        } else {
            const float v0 = pend ? pa0 : acca, v1 = pend ? pb0 : accb, v2 = pend ? acca : 0.0f, v3 = pend ? accb : 0.0f;
            const bool hi = (lane & 16) != 0, mid = (lane & 8) != 0;
            float t0 = __fadd_rn(hi ? v2 : v0, __shfl_xor_sync(0xffffffffu, hi ? v0 : v2, 16));
            float t1 = __fadd_rn(hi ? v3 : v1, __shfl_xor_sync(0xffffffffu, hi ? v1 : v3, 16));
            float r = __fadd_rn(mid ? t1 : t0, __shfl_xor_sync(0xffffffffu, mid ? t0 : t1, 8));
            r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
            r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
            r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
            const int ug = (pend && !hi) ? grow0 : grow;
            const bool utwo = (pend && !hi) ? two0 : two;
            const bool live = (!hi || pend) && (!mid || utwo);
            if (KIND == 2 && !staged) {
                const float gate = __shfl_xor_sync(0xffffffffu, r, 8);
                if ((lane & 15) == 0 && live) stf_f32(out + (ug >> 1), __fmul_rn(silu_ref(r), gate));
            } else if (KIND == 2) {
                // staged: the raw w1 / w3 sums wait in shared memory; SiLU * gate (expf, an IEEE division: ~400 cycles of latency
                // on two lanes) is done for all pairs at once by the whole CTA before the bulk store
                if ((lane & 7) == 0 && live) stage[2048 + ug + (mid ? 1 : 0) - r0] = r;
            } else if ((lane & 7) == 0 && live) {
                const float o = KIND == 1 ? __fadd_rn(xres, r) : r;
                if (staged) stage[ug + (mid ? 1 : 0) - r0] = o; else stf_f32(out + ug + (mid ? 1 : 0), o);
            }
            pend = false;
        }
    }
    if (staged) { // one TMA bulk store per CTA (see consume_mat)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bar_consumers();
        if (KIND == 2) { // swiglu (forward.c:134-139) of the CTA's row pairs, one pair per thread
            for (int i = threadIdx.x; i < (nrows >> 1); i += kConsumerThreads) stage[i] = __fmul_rn(silu_ref(stage[2048 + 2 * i]), stage[2048 + 2 * i + 1]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_consumers();
        }
        const int nout = KIND == 2 ? nrows >> 1 : nrows;
        float* dst = out + (KIND == 2 ? r0 >> 1 : r0);
        if (threadIdx.x == 0 && nrows > 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(stage)), "r"(nout * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        flush_stores();
        bar_consumers(); // the next prologue rewrites sh.xq: no warp may still be reading it (the staged branch has its barrier)
    }
}

// ---- attention items of one warp: item a = block a / 2 of its positions, K rows (a even) or V rows (a odd), sub-stage a % 3
struct PwAttn {
    int pos0, npos;   // the warp's cached positions [pos0, pos0 + npos) of the CTA's kv head
    int nitems;       // 2 * blocks
    int issued;       // items issued
    size_t base;      // float offset of (layer, kv head, position 0) in the caches
};
__device__ __forceinline__ void pw_attn_issue(const Shared& sh, const MegaParams& p, PwAttn& a, int warp, int lane) {
    if (a.issued >= a.nitems) return;
    const int blk = a.issued >> 1;
    const int cnt = min(kPwBlock, a.npos - blk * kPwBlock);
    if (lane == 0) {
        const unsigned sub = (unsigned) a.issued % 3u;
        const uint32_t bar = sh.full + (warp * 3 + sub) * 8;
        const float* src = ((a.issued & 1) ? p.v_cache : p.k_cache) + a.base + (size_t) (a.pos0 + blk * kPwBlock) * 128;
        mbar_expect_tx(bar, (uint32_t) cnt * 512u);
        bulk_g2s(smem_u32(sh.ring + (size_t) warp * kPwRegion + sub * kPwSub), src, (uint32_t) cnt * 512u, bar);
    }
    ++a.issued;
}
// the warp's share of the attention block of layer l: set up after its last QKV unit so that the first three items (K rows
// and V rows of block 0, K rows of block 1) travel while the CTA waits for q
template <int KV_MUL>
__device__ __forceinline__ PwAttn pw_attn_setup(const MegaParams& p, const PwTables& tb, int l, int warp) {
    constexpr int HW = KV_MUL < 4 ? KV_MUL : 4, NHG = KV_MUL / HW;
    PwAttn a{0, 0, 0, 0, 0};
    if (!tb.a_active) return a;
    const int hg = warp % NHG, wi = warp / NHG, wn = (kConsumerWarps - hg + NHG - 1) / NHG;
    const int np = tb.a_hi - tb.a_lo;
    const int my0 = np * wi / wn, my1 = np * (wi + 1) / wn;
    a.pos0 = tb.a_lo + my0;
    a.npos = my1 - my0;
    a.nitems = 2 * ((a.npos + kPwBlock - 1) / kPwBlock);
    a.base = ((size_t) l * p.KVHl + tb.a_kvh) * p.S * 128;
    return a;
}

template <int KV_MUL>
__device__ void pw_consume_attn(const Shared& sh, const MegaParams& p, const PwTables& tb, int l, PwAttn& at, PwCtl* ctl, unsigned& bpar) {
    constexpr int HW = KV_MUL < 4 ? KV_MUL : 4; // heads per warp
    constexpr int NHG = KV_MUL / HW;            // head groups per KV head
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* sq = sh.scr;
    float* fl = flow_layer(p, l);
    const float* qkv = fl + p.o_qkv;
    const float* gq = p.q_norm + (size_t) l * 128;
    const float* gk = p.k_norm + (size_t) l * 128;
    if (!tb.a_active) { // this block takes no part in attention
        if (lane == 0) ctl->unblocked[warp] = l;
        return;
    }
    const int kvh = tb.a_kvh, my_slot = tb.a_j;
    const int npos = tb.a_hi - tb.a_lo;
    const bool own_last = tb.a_j == tb.a_n - 1; // this block also takes the step's own position
    const int hg = warp % NHG, wi = warp / NHG, wn = (kConsumerWarps - hg + NHG - 1) / NHG;
    if (npos == 0 && !own_last) {
        // nothing to attend over here (short context): the kv head's combine tasks still expect this block's (m, l)
        if (warp == 0 && lane < KV_MUL) {
            float* dst = fl + p.o_part + ((size_t) (kvh * p.part_slots + my_slot) * KV_MUL + lane) * kPartStride;
            stf_f32(dst + 128, -INFINITY);
            stf_f32(dst + 129, 0.0f);
            flush_stores();
        }
        if (lane == 0) ctl->unblocked[warp] = l;
        return;
    }
    // ---- prologue: warp j < KV_MUL prepares query head j; warps KV_MUL, KV_MUL + 1 this step's K, V row (as in consume_attn)
    bar_consumers(); // previous users of the scratch are done
    if (warp < KV_MUL + (own_last ? 2 : 0)) {
        const float* src = warp < KV_MUL ? qkv + (size_t) (kvh * KV_MUL + warp) * 128
                           : warp == KV_MUL ? qkv + p.Pl + (size_t) kvh * 128
                                            : qkv + p.Pl + p.Kl + (size_t) kvh * 128;
        float x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = __uint_as_float(ldf_u32(src + lane + 32 * k));
        {
            unsigned long long t_start = 0;
#pragma unroll 1
            for (unsigned n = 1;; ++n) {
                bool busy = false;
#pragma unroll
                for (int k = 0; k < 4; ++k) busy |= __float_as_uint(x[k]) == kSent;
                if (!busy) break;
                if ((n & 255u) == 0 && poll_timed_out(sh.abort_flag, p.err, t_start, 13)) break;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (__float_as_uint(x[k]) == kSent) x[k] = __uint_as_float(ldf_u32(src + lane + 32 * k));
            }
        }
        float* dst = sq + warp * 128;
        if (warp <= KV_MUL) {
            head_norm_rope_warp(dst, x, warp < KV_MUL ? gq : gk, p, lane);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) dst[lane + 32 * k] = x[k];
        }
        if (warp >= KV_MUL) { // this step's K / V row goes into the cache for later steps
            __syncwarp();
            float* cache = (warp == KV_MUL ? p.k_cache : p.v_cache) + (((size_t) l * p.KVHl + kvh) * p.S + p.pos) * 128;
            *reinterpret_cast<float4*>(cache + lane * 4) = *reinterpret_cast<const float4*>(dst + lane * 4);
        }
    }
    bar_consumers();
    stamp(p, l, 1);
    // ---- this warp's blocks: K rows -> scores and softmax statistics -> hand the K sub-stage back -> V rows -> accumulate
    float4 q[HW];
#pragma unroll
    for (int j = 0; j < HW; ++j) q[j] = *reinterpret_cast<const float4*>(sq + (hg * HW + j) * 128 + lane * 4);
    AttnState<HW> st;
    attn_state_reset(st);
    const uint8_t* region = sh.ring + (size_t) warp * kPwRegion;
    const int nblk = at.nitems >> 1;
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
        const int cnt = min(kPwBlock, at.npos - b * kPwBlock);
        const unsigned sk = (unsigned) (2 * b) % 3u, sv = (unsigned) (2 * b + 1) % 3u;
        pw_wait(sh, p, warp, sk, bpar, 4);
        const float e = attn_block_scores<HW>(reinterpret_cast<const float*>(region + sk * kPwSub), cnt, q, st, lane);
        __syncwarp();
        pw_attn_issue(sh, p, at, warp, lane); // into the K rows' sub-stage
        pw_wait(sh, p, warp, sv, bpar, 4);
        attn_block_pv<HW>(reinterpret_cast<const float*>(region + sv * kPwSub), cnt, e, st, lane);
        __syncwarp();
        pw_attn_issue(sh, p, at, warp, lane);
    }
    // this step's own position, from shared memory: the last warp of each head group takes it
    if (own_last && wi == wn - 1)
        attn_one_row<HW>(*reinterpret_cast<const float4*>(sq + KV_MUL * 128 + lane * 4),
                         *reinterpret_cast<const float4*>(sq + (KV_MUL + 1) * 128 + lane * 4), q, st);
    // ---- merge the 15 warp states: every warp dumps its state into its own (now idle) region, 16 threads per head
    // compute the common maximum, the weights exp(m_w - M) and the sum, then all threads form the weighted sums
    {
        float* dump = reinterpret_cast<float*>(sh.ring + (size_t) warp * kPwRegion);
#pragma unroll
        for (int j = 0; j < HW; ++j) {
            *reinterpret_cast<float4*>(dump + j * kPartStride + lane * 4) = st.acc[j];
            if (lane == 0) {
                dump[j * kPartStride + 128] = st.m[j];
                dump[j * kPartStride + 129] = st.l[j];
            }
        }
    }
    bar_consumers(); // dumps visible; sq is dead from here
    stamp(p, l, 9);
    float* wts = sh.scr;                        // [KV_MUL][16] weights
    float* pub = sh.scr + KV_MUL * 16;          // [KV_MUL][kPartStride] the CTA's partial, as published
    if (tid < (KV_MUL * 16 < 32 ? 32 : KV_MUL * 16)) { // whole warps: the 16-lane shuffles below use the full mask
        const int jj = tid >> 4, w = tid & 15;  // head of the kv head, contributing warp
        const bool has = jj < KV_MUL && w < kConsumerWarps && (w % NHG) == jj / HW;
        const float* d = reinterpret_cast<const float*>(sh.ring + (size_t) w * kPwRegion) + (jj % HW) * kPartStride;
        const float mw = has ? d[128] : -INFINITY, lw = has ? d[129] : 0.0f;
        float M = mw;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
        const float wt = (mw == -INFINITY) ? 0.0f : expf(__fsub_rn(mw, M));
        float Ls = __fmul_rn(lw, wt);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) Ls = __fadd_rn(Ls, __shfl_xor_sync(0xffffffffu, Ls, o));
        if (jj < KV_MUL) {
            wts[tid] = wt;
            if (w == 0) {
                pub[jj * kPartStride + 128] = M;
                pub[jj * kPartStride + 129] = Ls;
                pub[jj * kPartStride + 130] = 0.0f; // pad words: never read
                pub[jj * kPartStride + 131] = 0.0f;
            }
        }
    }
    bar_consumers();
#pragma unroll 1
    for (int o = tid; o < KV_MUL * 128; o += kConsumerThreads) {
        const int jj = o >> 7, d = o & 127;
        float A = 0.0f;
#pragma unroll 1
        for (int w = (jj / HW); w < kConsumerWarps; w += NHG) {
            const float wt = wts[jj * 16 + w];
            if (wt != 0.0f) A = __fmaf_rn(reinterpret_cast<const float*>(sh.ring + (size_t) w * kPwRegion)[(jj % HW) * kPartStride + d], wt, A);
        }
        pub[jj * kPartStride + d] = A;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    bar_consumers(); // every dump has been read: the regions may take weights again
    if (threadIdx.x == 0) {
        float* gdst = fl + p.o_part + ((size_t) (kvh * p.part_slots + my_slot) * KV_MUL) * kPartStride;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(pub)), "r"(KV_MUL * kPartStride * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // combine_attn reuses the scratch after its barrier
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the dumps were generic-proxy writes; the next bulk copies land on them
    if (lane == 0) ctl->unblocked[warp] = l; // the region takes weights again
}

template <int KV_MUL>
__device__ void consumer_pw(const Shared& sh, const MegaParams& p, PwTables& tb, PwCtl* ctl, volatile int* progress) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the work split of this CTA, once
    if (threadIdx.x < 5) {
        int r0, r1;
        cta_rows(p.mat[threadIdx.x], p.perm, r0, r1);
        tb.r0[threadIdx.x] = r0;
        tb.r1[threadIdx.x] = r1;
    } else if (threadIdx.x == 32) {
        const AttnSplit a = attn_split(p);
        tb.a_kvh = a.kvh; tb.a_j = a.j; tb.a_n = a.n; tb.a_lo = a.p_lo; tb.a_hi = a.p_hi; tb.a_active = a.active ? 1 : 0;
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory"); // tables ready: consumers and the issue server (warp 15)
    unsigned cons_n = 0, bpar = 0;
    // refill the OTHER arena with the sentinel for the next launch (nobody reads it during this one)
    {
        const unsigned per = (unsigned) ((p.flow_words / 4 + gridDim.x - 1) / gridDim.x); // 16-byte pieces per CTA
        const unsigned a = per * blockIdx.x, b = min((unsigned) (p.flow_words / 4), a + per);
        uint4* dst = reinterpret_cast<uint4*>(p.flow_other);
        const uint4 s4 = make_uint4(kSent, kSent, kSent, kSent);
#pragma unroll 1
        for (unsigned i = a + threadIdx.x; i < b; i += kConsumerThreads) dst[i] = s4;
    }
    // the residual stream starts as the dequantised embedding row (forward.c:237): every CTA contributes its slice
    {
        const int tok = p.token_dev ? *p.token_dev : p.token;
        const uint8_t* row = p.w_emb + (size_t) tok * qw_row_bytes(p.D);
        const int c0 = (int) ((unsigned) p.D * blockIdx.x / gridDim.x), c1 = (int) ((unsigned) p.D * (blockIdx.x + 1) / gridDim.x);
#pragma unroll 1
        for (int c = c0 + threadIdx.x; c < c1; c += kConsumerThreads) {
            const uint8_t* rec = row + (size_t) (c >> 8) * QW_SG_BYTES;
            const float sc = *reinterpret_cast<const float*>(rec + 256 + ((c >> 6) & 3) * 4);
            stf_f32(p.flow_x0 + c, p.x_inject ? p.x_inject[c] : __fmul_rn((float) reinterpret_cast<const int8_t*>(rec)[c & 255], sc));
        }
        flush_stores();
    }
    const float* xprev = p.flow_x0;
    const int nph = 4 * p.layers_run;
    PwAttn at{0, 0, 0, 0, 0};
#pragma unroll 1
    for (int ph = 4 * p.l_begin; ph <= nph; ++ph) {
        const int l = ph >> 2, k = ph == nph ? 4 : (ph & 3);
        float* fl = flow_layer(p, k == 4 ? 0 : l);
        const int lp = k == 4 ? p.L : l; // profile row
        float* out;
        const float* resid = nullptr;
        stamp(p, lp, 4 * (k & 3));
        if (threadIdx.x == 0) *progress = k == 4 ? 5 * p.layers_run : 5 * l + (k == 0 ? 0 : k == 1 ? 1 : k + 1);
        if (k == 1) { // attention block, second half (forward.c:261-298)
            pw_consume_attn<KV_MUL>(sh, p, tb, l, at, ctl, bpar);
            stamp(p, lp, 13);
            if (threadIdx.x == 0) *progress = 5 * l + 2;
            combine_attn(sh, p, l);
            stamp(p, lp, 5);
            prologue_load_codes(sh, p, reinterpret_cast<const uint8_t*>(fl + p.o_attq), p.Pl);
            out = fl + p.o_xa;
            resid = xprev;
        } else {
            const float *src, *nw;
            int n = p.D;
            if (k == 0) { // forward.c:254-259
                src = xprev; nw = p.att_norm + (size_t) l * p.D; out = fl + p.o_qkv;
            } else if (k == 2) { // forward.c:303-318
                src = fl + p.o_xa; nw = p.ffn_norm + (size_t) l * p.D; out = fl + p.o_h;
            } else if (k == 3) { // forward.c:319-338
                src = fl + p.o_h; nw = nullptr; n = p.Hdl; out = fl + p.o_xb; resid = fl + p.o_xa;
            } else { // final norm + classifier (forward.c:344-348)
                src = xprev; nw = p.out_norm; out = p.logits;
            }
            prologue_quant<false>(sh, p, src, n, nw, -1);
        }
        if (p.dbg_codes && blockIdx.x == 0) { // debug: what this GEMV is fed (every CTA holds the same vector)
            const int pieces = (int) (qw_row_bytes(p.mat[k].n) / 16);
            uint4* dst = reinterpret_cast<uint4*>(p.dbg_codes + (size_t) (k == 4 ? 4 * p.L : 4 * l + k) * p.dbg_stride);
            for (int i = threadIdx.x; i < pieces; i += kConsumerThreads) dst[i] = reinterpret_cast<const uint4*>(sh.xq)[i];
        }
        stamp(p, lp, 4 * (k & 3) + 2);
        pw_consume_mat(sh, p, tb, k, ctl, cons_n, bpar, out, resid);
        if (k == 0) { // the warp's QKV units are done: its region is free, the first KV rows of the layer can travel during the q hand-off
            at = pw_attn_setup<KV_MUL>(p, tb, l, warp);
            pw_attn_issue(sh, p, at, warp, lane);
            pw_attn_issue(sh, p, at, warp, lane);
            pw_attn_issue(sh, p, at, warp, lane);
        }
        stamp(p, lp, 4 * (k & 3) + 3);
        if (k == 3) xprev = fl + p.o_xb;
    }
}

template <int KV_MUL>
__global__ void __launch_bounds__(kThreads, 1) k_decode_pw(const __grid_constant__ MegaParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ int abort_flag;
    __shared__ int progress; // sub-phase the consumers have reached (read by the L2 prefetcher)
    __shared__ PwTables tables;
    __shared__ PwCtl ctl;
    Shared sh;
    sh.ring = smem; // the 15 warp regions
    sh.xq = smem + p.off_xq;
    sh.scr = reinterpret_cast<float*>(smem + p.off_scr);
    sh.xres = nullptr;
    sh.misc = reinterpret_cast<float*>(smem + p.off_misc);
    sh.full = smem_u32(smem + p.off_bar); // [15 warps][3] barriers, each waited on by its warp only
    sh.empty = 0;
    sh.abort_flag = &abort_flag;
    if (threadIdx.x == 0) {
        abort_flag = 0;
        progress = 5 * p.l_begin;
        for (int w = 0; w < 16; ++w) {
            ctl.freed[w] = 0u;
            ctl.unblocked[w] = -1;
        }
        for (int s = 0; s < 3 * kConsumerWarps; ++s) mbar_init(sh.full + s * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= kConsumerThreads) {
        // the server reads the row tables the consumers fill first: wait for them
        asm volatile("bar.sync 2, %0;" ::"n"(kThreads) : "memory");
        pw_server(sh, p, tables, &ctl, &progress, &abort_flag);
        return;
    }
    consumer_pw<KV_MUL>(sh, p, tables, &ctl, &progress);
}
