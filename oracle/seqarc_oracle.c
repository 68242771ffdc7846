/*
 * TEST INFRASTRUCTURE -- CPU oracle for the SeqArc v1.6 entropy-coding hot path.
 * See seqarc_oracle.h for the rules about who may use this file.
 *
 * Every function restates, in plain C, what the reference binary /root/reference/SeqArc-1.6
 * does at the cited virtual address (the reference ships no source).
 * Reproduce an address with:  objdump -d -C --no-show-raw-insn SeqArc-1.6 | less
 */
#include "seqarc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ Encap vints
 * Encap::setID @0x420720, Encap::setSize @0x420780. */
static int put_id(uint8_t *p, uint64_t id) {
    int n = 1;
    while (n < 8 && id > (((uint64_t)1 << (7 * n)) - 2)) n++;
    uint64_t v = id | ((uint64_t)1 << (7 * n));
    for (int i = 0; i < n; i++) p[i] = (uint8_t)(v >> (8 * (n - 1 - i)));
    return n;
}
static void put_size(uint8_t *p, uint64_t size, int n) {
    uint64_t v = size | ((uint64_t)1 << (7 * n));
    for (int i = 0; i < n; i++) p[i] = (uint8_t)(v >> (8 * (n - 1 - i)));
}
static int get_vint(const uint8_t *p, uint64_t *v) {
    int n = 1;
    uint8_t m = 0x80;
    while (n <= 8 && !(p[0] & m)) { n++; m >>= 1; }
    uint64_t x = p[0] & (m - 1);
    for (int i = 1; i < n; i++) x = (x << 8) | p[i];
    *v = x;
    return n;
}
static void put_u32le(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static uint32_t get_u32le(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

/* ------------------------------------------------------------------ RangeCoder
 * Inlined everywhere in the reference; encode e.g. @0x422010-0x422085, flush @0x424a30,
 * decode @0x429760-0x429880 (Shelwien-style carry-less coder, 64-bit low / 32-bit range). */
typedef struct {
    uint64_t low;
    uint32_t range, code;
    const uint8_t *in;
    uint8_t *out;
} RC;

static void rc_start_enc(RC *rc, uint8_t *out) { rc->low = 0; rc->range = 0xffffffffu; rc->out = out; }
static void rc_finish_enc(RC *rc) {
    for (int i = 0; i < 8; i++) { *rc->out++ = (uint8_t)(rc->low >> 56); rc->low <<= 8; }
}
/* test hook: when set, every coded (cum, freq, tot) is appended as cum | freq<<16 | tot<<32 */
static uint64_t *g_trace;
static uint64_t g_ntrace;
static void rc_encode(RC *rc, uint32_t cum, uint32_t freq, uint32_t tot) {
    if (g_trace) g_trace[g_ntrace++] = (uint64_t)cum | ((uint64_t)freq << 16) | ((uint64_t)tot << 32);
    rc->range /= tot;
    rc->low += (uint32_t)(cum * rc->range);
    rc->range *= freq;
    if (cum + freq > tot) abort();
    while (rc->range < (1u << 24)) {
        if ((rc->low ^ (rc->low + rc->range)) >> 56)
            rc->range = (((uint32_t)rc->low) | 0xffffffu) - (uint32_t)rc->low;
        *rc->out++ = (uint8_t)(rc->low >> 56);
        rc->range <<= 8;
        rc->low <<= 8;
    }
}
static void rc_start_dec(RC *rc, const uint8_t *in) {
    rc->low = 0; rc->range = 0xffffffffu; rc->code = 0; rc->in = in;
    for (int i = 0; i < 8; i++) rc->code = (rc->code << 8) | *rc->in++;
}
static uint32_t rc_getfreq(RC *rc, uint32_t tot) { rc->range /= tot; return rc->code / rc->range; }
static void rc_decode(RC *rc, uint32_t cum, uint32_t freq) {
    uint32_t t = cum * rc->range;
    rc->low += t;
    rc->code -= t;
    rc->range *= freq;
    while (rc->range < (1u << 24)) {
        if ((rc->low ^ (rc->low + rc->range)) >> 56)
            rc->range = (((uint32_t)rc->low) | 0xffffffu) - (uint32_t)rc->low;
        rc->code = (rc->code << 8) | *rc->in++;
        rc->range <<= 8;
        rc->low <<= 8;
    }
}

/* ------------------------------------------------------------------ SIMPLE_MODEL<N>
 * ctor pattern @0x42f2b0-0x42f2ea; encodeSymbol inlined e.g. @0x4221fb-0x4222f1 and
 * @0x424ec0-0x424fa1.  Flat layout: w[0]=TotFreq w[1]=BubCnt w[2]=sentinel{sym 0,freq 0xffe0}
 * w[3..3+N-1]=F[i]{u16 sym,u16 freq}  w[3+N]=terminator (freq 0).  Stride N+4 words. */
#define SM_STRIDE(N) ((N) + 4)
#define SM_SYM(w) ((w) & 0xffffu)
#define SM_FREQ(w) ((w) >> 16)
#define SM_PACK(s, f) (((uint32_t)(f) << 16) | (uint32_t)(s))

static void sm_init(uint32_t *m, int N) {
    m[0] = (uint32_t)N; m[1] = 0; m[2] = SM_PACK(0, 0xffe0);
    for (int i = 0; i < N; i++) m[3 + i] = SM_PACK(i, 1);
    m[3 + N] = 0;
}
/* model update shared by encode/decode/training: k = index of coded symbol in F[] */
static void sm_update(uint32_t *m, int k) {
    uint32_t *F = m + 3;
    F[k] = SM_PACK(SM_SYM(F[k]), (SM_FREQ(F[k]) + 8) & 0xffff);
    uint32_t tot = m[0] + 8;
    if (tot > 0xffe0) {                 /* halve: @0x424fc0-0x425001 */
        tot = 0;
        for (uint32_t *p = F; SM_FREQ(*p) != 0; p++) {
            uint32_t f = SM_FREQ(*p);
            f -= f >> 1;
            *p = SM_PACK(SM_SYM(*p), f);
            tot += f;
        }
    }
    m[0] = tot;
    if (((++m[1]) & 15) == 0 && SM_FREQ(F[k]) > SM_FREQ(F[k - 1])) {   /* F[-1] = sentinel */
        uint32_t t = F[k]; F[k] = F[k - 1]; F[k - 1] = t;
    }
}
static void sm_encode(uint32_t *m, RC *rc, uint32_t sym) {
    uint32_t *F = m + 3, acc = 0;
    int k = 0;
    while (SM_SYM(F[k]) != sym) { acc += SM_FREQ(F[k]); k++; }
    rc_encode(rc, acc, SM_FREQ(F[k]), m[0]);
    sm_update(m, k);
}
static void sm_train(uint32_t *m, uint32_t sym) {   /* *_formodel twins: update without coding */
    uint32_t *F = m + 3;
    int k = 0;
    while (SM_SYM(F[k]) != sym) k++;
    sm_update(m, k);
}
/* decodeSymbol: inlined e.g. @0x42a7d0-0x42a960 */
static uint32_t sm_decode(uint32_t *m, RC *rc) {
    uint32_t *F = m + 3;
    uint32_t target = rc_getfreq(rc, m[0]);
    uint32_t acc = 0;
    int k = 0;
    while (acc + SM_FREQ(F[k]) <= target) { acc += SM_FREQ(F[k]); k++; }
    rc_decode(rc, acc, SM_FREQ(F[k]));
    uint32_t sym = SM_SYM(F[k]);
    sm_update(m, k);
    return sym;
}

/* ------------------------------------------------------------------ context object */
static const uint8_t seq_val_table_init[] = "ACGTNMRYKSWHBVD";   /* this+0x15a0 */
static uint8_t seq_val[256];                                      /* tables @0x44b800 / @0x44bd00 */

struct sqo_ctx {
    sqo_params p;
    int NS;                 /* this+0x1494 = 7 + Slevel */
    uint32_t nseq;          /* this+0x0    = 4^NS */
    uint32_t seqmask;       /* this+0x1578 */
    uint32_t nqual;         /* this+0x4    = 0x10000 (Qlevel<=2) | 0x100000 */
    uint8_t *seqtab;        /* this+0x1498 */
    uint32_t *qual;         /* this+0x1500: SIMPLE_MODEL<95>[nqual] */
    uint32_t *name_prefix, *name_suffix, *name_len;   /* SIMPLE_MODEL<256>[256] x3 */
    uint32_t *name_middle;                            /* SIMPLE_MODEL<128>[8192] */
    uint32_t len1[SM_STRIDE(256)], len2[SM_STRIDE(256)], len3[SM_STRIDE(256)], len4[SM_STRIDE(256)], same_len[SM_STRIDE(2)];
    int32_t last_len;       /* this+0x1060: only the decoders write it */
    uint8_t last_name[1024];
    int32_t lnl, lpl, lsl;  /* this+0x1488/0x148c/0x1490 */
    uint32_t knb[SM_STRIDE(64)];
    uint32_t kbits[64][SM_STRIDE(2)];
    /* decode-side scratch (membuf+0x150.. arrays) */
    uint8_t *d_tip, *d_ch, *d_maxq;
    uint32_t *d_cnt, *d_pos;
    uint32_t n_tip, n_ch, n_maxq, n_cnt, n_pos;
    /* aligned mode (decode only, single-end): the reference sequence and the side arrays of the block (membuf+0xe0..0x138) */
    const uint8_t *ref; uint64_t reflen;
    uint8_t *a_order, *a_pos, *a_mis, *a_rev, *a_cigal, *a_cigav;
    uint32_t n_order, n_align, n_posbits, n_cigal, n_cigav;
    uint32_t al_i, al_pos, al_cigal, al_cigav;      /* cursors: this+0x1758, +0x1718, +0x1720, +0x175c */
    uint8_t *a_perel; uint32_t n_perel, n_insbits, al_perel;   /* paired-end: PERelation array (membuf+0x140), PEInsrNum (+0x28), cursor this+0x1764 */
    uint64_t al_cur;                                 /* this+0x1710: the position computed last */
    /* encode side: the aligner's arrays for the NEXT sqo_encode_block call (sqo_set_align; the caller keeps them alive) */
    const uint8_t *e_order, *e_pos, *e_mis, *e_rev, *e_cigal, *e_cigav;
    uint32_t e_align, e_posbits, e_ncigal, e_ncigav;
    const uint8_t *e_perel; uint32_t e_nperel, e_insbits;     /* paired-end (sqo_set_align_pe) */
};

static void init_tables(void) {
    if (seq_val['C'] == 1) return;
    memset(seq_val, 4, 256);
    for (int i = 0; i < 15; i++) {
        seq_val[seq_val_table_init[i]] = i;
        seq_val[seq_val_table_init[i] | 0x20] = i;
    }
}

sqo_ctx *sqo_create(const sqo_params *p) {   /* ctor @0x42f2a0 + initModel @0x42ed00 */
    init_tables();
    sqo_ctx *c = (sqo_ctx *)calloc(1, sizeof(*c));
    c->p = *p;
    c->NS = 7 + p->slevel;
    c->nseq = 1u << (2 * c->NS);
    c->seqmask = c->nseq - 1;
    c->nqual = p->qlevel > 2 ? 0x100000u : 0x10000u;
    c->seqtab = (uint8_t *)malloc((size_t)c->nseq * 4);
    c->qual = (uint32_t *)malloc((size_t)c->nqual * SM_STRIDE(95) * 4);
    c->name_prefix = (uint32_t *)malloc(256 * SM_STRIDE(256) * 4);
    c->name_suffix = (uint32_t *)malloc(256 * SM_STRIDE(256) * 4);
    c->name_len = (uint32_t *)malloc(256 * SM_STRIDE(256) * 4);
    c->name_middle = (uint32_t *)malloc(8192 * SM_STRIDE(128) * 4);
    c->last_len = 0;
    return c;
}
void sqo_destroy(sqo_ctx *c) {
    if (!c) return;
    free(c->seqtab); free(c->qual); free(c->name_prefix); free(c->name_suffix); free(c->name_len);
    free(c->name_middle); free(c->d_tip); free(c->d_ch); free(c->d_maxq); free(c->d_cnt); free(c->d_pos);
    free(c->a_perel); free(c->a_order); free(c->a_pos); free(c->a_mis); free(c->a_rev); free(c->a_cigal); free(c->a_cigav);
    free(c);
}

/* ------------------------------------------------------------------ model snapshot
 * SaveModelToMem @0x420c00 packs SIMPLE_MODEL<N> as u32le Tot, u32le Bub, N x {u8 sym, u16le freq};
 * unpack loops e.g. @0x424210-0x4242b6 (name), @0x42705c-0x42712e (qual).  Sentinel and
 * terminator are not stored. Layout: prefix[256] suffix[256] len[256] (776 B each),
 * middle[8192] (392 B each), seq table (4^NS*4 B), qual[nqual] (293 B each). */
static const uint8_t *sm_unpack(uint32_t *m, int N, const uint8_t *s) {
    m[0] = get_u32le(s); m[1] = get_u32le(s + 4); m[2] = SM_PACK(0, 0xffe0);
    s += 8;
    for (int i = 0; i < N; i++, s += 3) m[3 + i] = SM_PACK(s[0], s[1] | (s[2] << 8));
    m[3 + N] = 0;
    return s;
}
static uint8_t *sm_pack(const uint32_t *m, int N, uint8_t *s) {
    put_u32le(s, m[0]); put_u32le(s + 4, m[1]);
    s += 8;
    for (int i = 0; i < N; i++, s += 3) { s[0] = (uint8_t)SM_SYM(m[3 + i]); s[1] = SM_FREQ(m[3 + i]) & 0xff; s[2] = SM_FREQ(m[3 + i]) >> 8; }
    return s;
}
static uint64_t snap_name_bytes(void) { return 3ull * 256 * 776 + 8192ull * 392; }
uint64_t sqo_model_size(const sqo_params *p) {
    uint64_t nseq = 1ull << (2 * (7 + p->slevel));
    uint64_t nqual = p->qlevel > 2 ? 0x100000u : 0x10000u;
    return (p->id_rule ? 0 : snap_name_bytes()) + nseq * 4 + nqual * 293;
}
static const uint8_t *snap_seq_ptr(const sqo_ctx *c) { return c->p.model + (c->p.id_rule ? 0 : snap_name_bytes()); }
static const uint8_t *snap_qual_ptr(const sqo_ctx *c) { return snap_seq_ptr(c) + (uint64_t)c->nseq * 4; }

static void reset_name_models(sqo_ctx *c) {     /* compressName @0x4241e8-0x42452c / @0x424604-0x4247b9 */
    if (c->p.usemodel) {
        const uint8_t *s = c->p.model;
        for (int i = 0; i < 256; i++) s = sm_unpack(c->name_prefix + i * SM_STRIDE(256), 256, s);
        for (int i = 0; i < 256; i++) s = sm_unpack(c->name_suffix + i * SM_STRIDE(256), 256, s);
        for (int i = 0; i < 256; i++) s = sm_unpack(c->name_len + i * SM_STRIDE(256), 256, s);
        for (int i = 0; i < 8192; i++) s = sm_unpack(c->name_middle + i * SM_STRIDE(128), 128, s);
    } else {
        for (int i = 0; i < 256; i++) {
            sm_init(c->name_prefix + i * SM_STRIDE(256), 256);
            sm_init(c->name_suffix + i * SM_STRIDE(256), 256);
            sm_init(c->name_len + i * SM_STRIDE(256), 256);
        }
        for (int i = 0; i < 8192; i++) sm_init(c->name_middle + i * SM_STRIDE(128), 128);
    }
    memset(c->last_name, ' ', 1024);            /* @0x42452c-0x42457a */
    c->lnl = c->lpl = c->lsl = 0;
}
static void reset_seq_model(sqo_ctx *c) {       /* compressSeq @0x424931-0x424963 / @0x424af8 */
    if (c->p.usemodel) memcpy(c->seqtab, snap_seq_ptr(c), (size_t)c->nseq * 4);
    else memset(c->seqtab, 3, (size_t)c->nseq * 4);
}
static void reset_qual_models(sqo_ctx *c) {     /* compressQual @0x426f10-0x426f84 / @0x42705c-0x42712e */
    if (c->p.usemodel) {
        const uint8_t *s = snap_qual_ptr(c);
        for (uint32_t i = 0; i < c->nqual; i++) s = sm_unpack(c->qual + (size_t)i * SM_STRIDE(95), 95, s);
    } else {
        for (uint32_t i = 0; i < c->nqual; i++) sm_init(c->qual + (size_t)i * SM_STRIDE(95), 95);
    }
}

/* ------------------------------------------------------------------ Len stream
 * compressLen @0x424120, compressLen_short @0x423f50, encode_len_short @0x4239a0.  A block holding a read
 * longer than 65,535 bases (membuf+0x2, low bit of the block-index length field) goes through
 * compressLen_long @0x423710 / encode_len_long @0x422e70: four length bytes with four SIMPLE_MODEL<256>
 * (this+0x8, +0x418, +0x828, +0xc38), the same "last_len is never written" comparison. */
int64_t sqo_stream_len(sqo_ctx *c, const sqo_block *b, uint8_t *out) {
    RC rc; rc_start_enc(&rc, out);
    sm_init(c->len1, 256); sm_init(c->len2, 256); sm_init(c->same_len, 2);
    if (b->longflag) { sm_init(c->len3, 256); sm_init(c->len4, 256); }
    for (uint32_t i = 0; i < b->nreads; i++) {
        int32_t len = b->seqlen[i];
        if (len != 0) {   /* cmp with this+0x1060, which the encoder never writes (ctor value 0) */
            sm_encode(c->same_len, &rc, 0);
            sm_encode(c->len1, &rc, len & 0xff);
            sm_encode(c->len2, &rc, (len >> 8) & 0xff);
            if (b->longflag) {
                sm_encode(c->len3, &rc, (len >> 16) & 0xff);
                sm_encode(c->len4, &rc, ((uint32_t)len >> 24) & 0xff);
            }
        } else {
            sm_encode(c->same_len, &rc, 1);
        }
    }
    rc_finish_enc(&rc);
    return rc.out - out;
}

/* ------------------------------------------------------------------ Name stream
 * encode_name @0x421070 (fqzcomp nlevel-1 name model; SURVEY App. F pseudo-code). */
static void encode_name(sqo_ctx *c, RC *rc, const uint8_t *name, int n, int train) {
    uint8_t *last = c->last_name;
    int lnl = c->lnl, p = 0, s = 0;
    if (n > 0 && lnl > 0 && name[0] == last[0]) {
        p = 1;
        while (p < n && p < lnl && name[p] == last[p]) p++;
    }
    if (n > 0 && lnl > 0) {
        int i = n - 1, j = lnl - 1;
        while (i >= 0 && j >= 0 && name[i] == last[j]) { i--; j--; }
        s = n - 1 - i;
    }
    if (n - s - p < 0) s = n - p;
#define ENC(m, N, ctx, sym) do { uint32_t *mm = (m) + (size_t)(ctx) * SM_STRIDE(N); if (train) sm_train(mm, (sym)); else sm_encode(mm, rc, (sym)); } while (0)
    ENC(c->name_prefix, 256, c->lpl, (uint32_t)p);
    ENC(c->name_suffix, 256, c->lsl, (uint32_t)s);
    ENC(c->name_len, 256, lnl, (uint32_t)n);
    c->lpl = p; c->lsl = s;
    int len2 = n - s, lc2 = (p != 0), i = p, j = p, k = 0;
    while (i < len2) {
        int v = ((int)last[j] - 32) * 2 + lc2 + k * 64;
        int ctx = v % 8192;
        uint8_t ch = name[i];
        ENC(c->name_middle, 128, ctx, (uint32_t)(ch & 0x7f));
        if (ch == ' ' && last[j] != ' ') j++;
        if (ch != ' ' && last[j] == ' ') j--;
        if (ch == ':' && last[j] != ':') j++;
        if (ch != ':' && last[j] == ':') j--;
        if (ch == ':' || ch == ' ') k = (k + 3) & ~3;
        if (j < 0) { j = 0; lc2 = 0; }
        else { lc2 = (ch == last[j]); j++; }
        k++; i++;
    }
#undef ENC
    memcpy(last, name, (size_t)n);
    c->lnl = n;
}
int64_t sqo_stream_name(sqo_ctx *c, const sqo_block *b, uint8_t *out) {   /* compressName @0x4241a0 */
    RC rc; rc_start_enc(&rc, out);
    reset_name_models(c);
    const uint8_t *p = b->ids;
    for (uint32_t i = 0; i < b->nreads; i++) { encode_name(c, &rc, p, b->idlen[i], 0); p += b->idlen[i]; }
    rc_finish_enc(&rc);
    return rc.out - out;
}

/* ------------------------------------------------------------------ Seq stream
 * encode_seq @0x421f30; base model = 4 byte counters per context (fqzcomp BASE_MODEL<uint8_t>). */
static void encode_seq(sqo_ctx *c, RC *rc, const uint8_t *seq, int len, int train) {
    uint32_t last = 0x7616c7u & c->seqmask;
    for (int i = 0; i < len; i++) {
        uint32_t bb = seq_val[seq[i]];
        if (bb > 3) continue;                               /* @0x421f79: non-ACGT skipped */
        uint8_t *S = c->seqtab + (size_t)last * 4;
        uint32_t tot = S[0] + S[1] + S[2] + S[3];
        if (tot > 253) {                                    /* @0x421fb0 */
            for (int j = 0; j < 4; j++) S[j] -= S[j] >> 1;
            tot = S[0] + S[1] + S[2] + S[3];
        }
        if (!train) {
            uint32_t cum = 0;
            for (uint32_t j = 0; j < bb; j++) cum += S[j];
            rc_encode(rc, cum, S[bb], tot);
        }
        S[bb]++;                                            /* @0x422087 */
        last = (last * 4 + bb) & c->seqmask;                /* @0x422090 */
    }
}
int64_t sqo_stream_seq(sqo_ctx *c, const sqo_block *b, uint8_t *out) {    /* compressSeq @0x4248a0 */
    RC rc; rc_start_enc(&rc, out);
    reset_seq_model(c);
    const uint8_t *p = b->seq;
    for (uint32_t i = 0; i < b->nreads; i++) { encode_seq(c, &rc, p, b->seqlen[i], 0); p += b->seqlen[i]; }
    rc_finish_enc(&rc);
    return rc.out - out;
}

/* ------------------------------------------------------------------ Qual stream
 * encode_qual @0x422180 (context update @0x422474-0x422505). */
static void encode_qual(sqo_ctx *c, RC *rc, const uint8_t *q, int len, int train) {
    int len2 = len;
    while (len2 > 0 && q[len2 - 1] == '#') len2--;
    uint32_t last = 0;
    int delta = 5, q1 = 0, q2 = 0;
    for (int i = 0; i < len2; i++) {
        uint32_t qq = (uint8_t)(q[i] - 33);
        uint32_t *m = c->qual + (size_t)last * SM_STRIDE(95);
        if (train) sm_train(m, qq); else sm_encode(m, rc, qq);
        last = (uint32_t)(((q1 > q2 ? q1 : q2) << 6) + (int)qq) & 0xfff;
        if (c->p.qlevel > 1) {
            last += (uint32_t)(q1 == q2) << 12;
            delta += (q1 > (int)qq) * (q1 - (int)qq);
            last += (uint32_t)((delta < 56 ? delta : 56) & 0xf8) << 10;
        }
        if (c->p.qlevel > 2) last += (i <= 111) ? (uint32_t)((i + 15) & 0x78) << 13 : 0xf0000u;
        q2 = q1; q1 = (int)qq;
    }
    if (len2 != len) {
        uint32_t *m = c->qual + (size_t)last * SM_STRIDE(95);
        if (train) sm_train(m, 94); else sm_encode(m, rc, 94);
    }
}

/* rblock @0x426c10 (R-Block lossy smoothing).  End of data = n (the reference
 * scans for a '\n' that is never written; the newline-guard shim gives these semantics). */
static double rnd(double x) { return floor(x + 0.5); }
void sqo_rblock(uint8_t *q, uint64_t n, double r) {
    if (n == 0) return;
    uint64_t start = 0, i = 1;
    int mn = q[0], mx = q[0];
    for (;;) {
        if (i >= n) { memset(q + start, (int)(uint8_t)rnd(sqrt((double)mn * mx)), i - start); return; }
        int ch = q[i];
        if (ch >= mn && ch <= mx) { i++; continue; }
        if (ch > mx) {
            double g = rnd(sqrt((double)ch * mn));
            if (r > g / mn && r > ch / g) { mx = ch; i++; continue; }
        } else {
            double g = rnd(sqrt((double)ch * mx));
            if (r > g / ch && r > mx / g) { mn = ch; i++; continue; }
        }
        memset(q + start, (int)(uint8_t)rnd(sqrt((double)mn * mx)), i - start);
        start = i; mn = mx = ch; i++;
    }
}

int64_t sqo_stream_qual(sqo_ctx *c, const sqo_block *b, uint8_t *out) {   /* compressQual @0x426e80 */
    RC rc; rc_start_enc(&rc, out);
    reset_qual_models(c);
    if (c->p.lossy) sqo_rblock(b->qual, b->seqbytes, c->p.lossy_r);      /* in place, as the reference */
    const uint8_t *p = b->qual;
    for (uint32_t i = 0; i < b->nreads; i++) { encode_qual(c, &rc, p, b->seqlen[i], 0); p += b->seqlen[i]; }
    rc_finish_enc(&rc);
    return rc.out - out;
}

/* ------------------------------------------------------------------ Dege streams
 * DegeInfoProcess @0x433a10 builds five arrays per block from (seq, qual) BEFORE lossy smoothing:
 *   tip[nreads]   1 if the read has any non-ACGT base                       (membuf+0x150)
 *   ch[]          the raw non-ACGT characters, in order                      (membuf+0x160)
 *   maxq[]        per flagged read: max quality char over its non-ACGT bases (membuf+0x170)
 *   cnt[]         per flagged read: number of ACGT bases with quality <= maxq (membuf+0x180)
 *   pos[]         per such base: # of non-ACGT bases (with q<=maxq) since the previous one (membuf+0x190)
 */
typedef struct {
    uint8_t *tip, *ch, *maxq;
    uint32_t *cnt, *pos;
    uint32_t n_tip, n_ch, n_maxq, n_cnt, n_pos;
} dege_t;

static void dege_build(const sqo_block *b, dege_t *d) {
    uint64_t nb = b->seqbytes ? b->seqbytes : 1;
    d->tip = (uint8_t *)malloc(b->nreads + 1); d->ch = (uint8_t *)malloc(nb);
    d->maxq = (uint8_t *)malloc(b->nreads + 1); d->cnt = (uint32_t *)malloc(4ull * (b->nreads + 1));
    d->pos = (uint32_t *)malloc(4 * nb);
    d->n_tip = d->n_ch = d->n_maxq = d->n_cnt = d->n_pos = 0;
    const uint8_t *s = b->seq, *q = b->qual;
    for (uint32_t r = 0; r < b->nreads; r++) {
        int len = b->seqlen[r], cnt = 0;
        uint8_t maxq = 0;
        for (int i = 0; i < len; i++)
            if (seq_val[s[i]] > 3) {
                d->ch[d->n_ch++] = s[i]; cnt++;
                if ((int)(int8_t)q[i] > (int)maxq) maxq = q[i];
            }
        if (cnt == 0) { d->tip[d->n_tip++] = 0; }
        else {
            d->tip[d->n_tip++] = 1;
            d->maxq[d->n_maxq++] = maxq;
            uint32_t gap = 0, n = 0;
            for (int i = 0; i < len; i++) {
                if ((int)maxq < (int)(int8_t)q[i]) continue;
                if (seq_val[s[i]] > 3) gap++;
                else { d->pos[d->n_pos++] = gap; n++; gap = 0; }
            }
            d->cnt[d->n_cnt++] = n;
        }
        s += len; q += len;
    }
}
static void dege_free(dege_t *d) { free(d->tip); free(d->ch); free(d->maxq); free(d->cnt); free(d->pos); }

/* kModelInit @0x42cbe0, kModelEncode @0x42ccb0, kModelDecode @0x42d890 */
static void kmodel_init(sqo_ctx *c) {
    sm_init(c->knb, 64);
    for (int i = 0; i < 64; i++) sm_init(c->kbits[i], 2);
}
static void kmodel_encode(sqo_ctx *c, RC *rc, uint64_t v) {
    uint8_t bits[64];
    int nb = 0;
    while (v) { bits[nb++] = v & 1; v >>= 1; }
    sm_encode(c->knb, rc, (uint32_t)nb);
    for (int i = 0; i < nb; i++) sm_encode(c->kbits[i], rc, bits[i]);
}
static uint64_t kmodel_decode(sqo_ctx *c, RC *rc) {
    int nb = (int)sm_decode(c->knb, rc);
    uint64_t v = 0;
    for (int i = 0; i < nb; i++) v |= (uint64_t)sm_decode(c->kbits[i], rc) << i;
    return v;
}

/* generic "tag | size(4) | u32le count | RC bytes" emitters; return total bytes written */
static int64_t emit_small_sm(uint8_t *out, int tag, int N, const uint8_t *sym, uint32_t n, int mode) {
    /* compressDegeTip @0x424dd0 (N=2, raw flag), compressDegeCh @0x425080 (N=11, seq_val-4),
     * compressDegeMaxQual @0x425310 (N=95, maxq-33): fresh local model per call. */
    if (n == 0) return 0;
    uint8_t *p = out;
    p += put_id(p, tag);
    uint8_t *szslot = p; p += 4;
    put_u32le(p, n); p += 4;
    uint32_t m[SM_STRIDE(95)];
    sm_init(m, N);
    RC rc; rc_start_enc(&rc, p);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t s = sym[i];
        if (mode == 1) s = (uint32_t)seq_val[s] - 4;
        else if (mode == 2) s = (uint16_t)(s - 33);
        sm_encode(m, &rc, s);
    }
    rc_finish_enc(&rc);
    put_size(szslot, (uint64_t)(rc.out - p) + 4, 4);
    return rc.out - out;
}
/* Aligned-mode side streams (doAlignEncode @0x42d4c0): each is the bare range-coder output of a byte array through a
 * fresh SIMPLE_MODEL<N> -- no count in front, the counts travel in their own tags (17 AlignCount membuf+0x30,
 * 27 OrderCount +0xe8, 18 PosCount +0xf8, 19 CigalCount +0x108, 20 CigaVCount +0x138):
 *   tag  8 Order  compressOrder          @0x424b70  N = 5   array membuf+0xe0, one symbol per read (loop count membuf+0x34)
 *   tag 11 Pos    compressAlignInfo_Pos  @0x425d70  N = 2   array +0xf0, PosCount bits
 *   tag 15 Mis    compressAlignInfo_Mis  @0x425ff0  N = 8 (SeqArcParam+0x1b60 in 1..7) or 9 (== 8), array +0x110, one per aligned read (+0x118)
 *   tag 10 Rev    compressAlignInfo_Rev  @0x426480  N = 2   array +0x120, one per aligned read (+0x128)
 *   tag 12 CigaL  compressAlignInfo_CigaL @0x426700 N = 2   array +0x100, CigalCount bits
 *   tag 13 CigaV  compressAlignInfo_CigaV @0x426980 N = 4   array +0x130, CigaVCount symbols
 * What the arrays MEAN (alignment, AlignInfoToSeq @0x42e020) is outside the hot path; these two functions restate the
 * coders only and are pinned by decoding + re-encoding the streams of a reference-written aligned block. */
int64_t sqo_small_encode(int N, const uint8_t *sym, uint32_t n, uint8_t *out) {
    uint32_t m[SM_STRIDE(256)];
    if (N < 2 || N > 256) return -1;
    sm_init(m, N);
    RC rc; rc_start_enc(&rc, out);
    for (uint32_t i = 0; i < n; i++) {
        if (sym[i] >= (uint32_t)N) return -1;
        sm_encode(m, &rc, sym[i]);
    }
    rc_finish_enc(&rc);
    return rc.out - out;
}
/* `in` must be followed by at least 8 readable bytes of slack (the decoder reads ahead) */
int64_t sqo_small_decode(int N, const uint8_t *in, uint32_t n, uint8_t *sym) {
    uint32_t m[SM_STRIDE(256)];
    if (N < 2 || N > 256) return -1;
    sm_init(m, N);
    RC rc; rc_start_dec(&rc, in);
    for (uint32_t i = 0; i < n; i++) sym[i] = (uint8_t)sm_decode(m, &rc);
    return rc.in - in;
}

static int64_t emit_kmodel(sqo_ctx *c, uint8_t *out, int tag, const uint32_t *v, uint32_t n) {
    /* compressNDegeCnt @0x42d010 (tag 25), compressNDegePos @0x42d170 (tag 26) */
    if (n == 0) return 0;
    uint8_t *p = out;
    p += put_id(p, tag);
    uint8_t *szslot = p; p += 4;
    put_u32le(p, n); p += 4;
    RC rc; rc_start_enc(&rc, p);
    for (uint32_t i = 0; i < n; i++) kmodel_encode(c, &rc, v[i]);
    rc_finish_enc(&rc);
    put_size(szslot, (uint64_t)(rc.out - p) + 4, 4);
    return rc.out - out;
}

/* ------------------------------------------------------------------ MD5 (RFC 1321; Md5.cpp @0x404de0) */
static const uint32_t md5_k[64] = {
    0xd76aa478,0xe8c7b756,0x242070db,0xc1bdceee,0xf57c0faf,0x4787c62a,0xa8304613,0xfd469501,
    0x698098d8,0x8b44f7af,0xffff5bb1,0x895cd7be,0x6b901122,0xfd987193,0xa679438e,0x49b40821,
    0xf61e2562,0xc040b340,0x265e5a51,0xe9b6c7aa,0xd62f105d,0x02441453,0xd8a1e681,0xe7d3fbc8,
    0x21e1cde6,0xc33707d6,0xf4d50d87,0x455a14ed,0xa9e3e905,0xfcefa3f8,0x676f02d9,0x8d2a4c8a,
    0xfffa3942,0x8771f681,0x6d9d6122,0xfde5380c,0xa4beea44,0x4bdecfa9,0xf6bb4b60,0xbebfbc70,
    0x289b7ec6,0xeaa127fa,0xd4ef3085,0x04881d05,0xd9d4d039,0xe6db99e5,0x1fa27cf8,0xc4ac5665,
    0xf4292244,0x432aff97,0xab9423a7,0xfc93a039,0x655b59c3,0x8f0ccc92,0xffeff47d,0x85845dd1,
    0x6fa87e4f,0xfe2ce6e0,0xa3014314,0x4e0811a1,0xf7537e82,0xbd3af235,0x2ad7d2bb,0xeb86d391};
static const uint8_t md5_s[64] = {7,12,17,22,7,12,17,22,7,12,17,22,7,12,17,22,5,9,14,20,5,9,14,20,5,9,14,20,5,9,14,20,
                                  4,11,16,23,4,11,16,23,4,11,16,23,4,11,16,23,6,10,15,21,6,10,15,21,6,10,15,21,6,10,15,21};
static void md5_block(uint32_t h[4], const uint8_t *p) {
    uint32_t w[16], a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 16; i++) w[i] = get_u32le(p + 4 * i);
    for (int i = 0; i < 64; i++) {
        uint32_t f; int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        f += a + md5_k[i] + w[g];
        a = d; d = c; c = b;
        b += (f << md5_s[i]) | (f >> (32 - md5_s[i]));
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}
void sqo_md5(const uint8_t *data, uint64_t n, uint8_t digest[16]) {
    uint32_t h[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
    uint64_t i = 0;
    for (; i + 64 <= n; i += 64) md5_block(h, data + i);
    uint8_t tail[128];
    uint64_t rem = n - i;
    memset(tail, 0, sizeof tail);
    memcpy(tail, data + i, rem);
    tail[rem] = 0x80;
    uint64_t tl = rem < 56 ? 64 : 128, bits = n * 8;
    for (int k = 0; k < 8; k++) tail[tl - 8 + k] = (uint8_t)(bits >> (8 * k));
    md5_block(h, tail);
    if (tl == 128) md5_block(h, tail + 64);
    for (int k = 0; k < 4; k++) put_u32le(digest + 4 * k, h[k]);
}

/* Framing added by doAlignEncode @0x42d4c0 when a reference index is loaded but nothing mapped.
 * compressCount(tag, v) @0x422a00 = tag | 0x84 | u32le; an alignment stream with no symbols is the
 * 8 flush bytes of an untouched range coder.  tested_reads: AlignEncodeSEJob::doAlign @0x411910 /
 * AlignEncodePEJob::doAlign @0x413580 try the head of the block -- (int)(nreads * 0.05) reads, constant
 * @0x44a218 -- pushing one flag per tried read into membuf+0xe0 (count at +0xe8, written as tag 27),
 * and give up at the first ratio check when none mapped. */
static int put_count(uint8_t *p, int tag, uint32_t v) {
    int n = put_id(p, (uint64_t)tag);
    p[n] = 0x84; put_u32le(p + n + 1, v);
    return n + 5;
}
static int64_t put_side(uint8_t *p, int tag, int N, const uint8_t *sym, uint32_t n) {   /* "tag | size(4) | RC bytes" */
    int k = put_id(p, (uint64_t)tag);
    int64_t len = sqo_small_encode(N, sym, n, p + k + 4);
    put_size(p + k, (uint64_t)len, 4);
    return k + 4 + len;
}
/* Aligned mode, encode side, single-end (doAlignEncode @0x42d4c0): what the aligner found for the block the next
 * sqo_encode_block call codes -- order[nreads] (0 = not mapped), n_posbits position bits, mis[n_align], rev[n_align],
 * n_cigal offset bits, cigav[n_cigav].  The arrays' meaning: seqarc_oracle_dec.inc (align_to_seq).  n_align = 0 clears. */
void sqo_set_align(sqo_ctx *c, const uint8_t *order, const uint8_t *pos, uint32_t n_posbits, const uint8_t *mis, const uint8_t *rev,
                   uint32_t n_align, const uint8_t *cigal, uint32_t n_cigal, const uint8_t *cigav, uint32_t n_cigav) {
    c->e_order = order; c->e_pos = pos; c->e_mis = mis; c->e_rev = rev; c->e_cigal = cigal; c->e_cigav = cigav;
    c->e_posbits = n_posbits; c->e_align = n_align; c->e_ncigal = n_cigal; c->e_ncigav = n_cigav;
}
/* the side arrays of the block sqo_decode_block decoded last: which = 0 order, 1 pos bits, 2 mis, 3 rev, 4 cigal bits, 5 cigav */
uint32_t sqo_get_align(const sqo_ctx *c, int which, const uint8_t **arr) {
    switch (which) {
    case 0: *arr = c->a_order; return c->n_order;
    case 1: *arr = c->a_pos; return c->n_posbits;
    case 2: *arr = c->a_mis; return c->n_align;
    case 3: *arr = c->a_rev; return c->n_align;
    case 4: *arr = c->a_cigal; return c->n_cigal;
    case 5: *arr = c->a_cigav; return c->n_cigav;
    case 6: *arr = c->a_perel; return c->n_perel;
    case 7: *arr = 0; return c->n_insbits;
    }
    *arr = 0; return 0;
}
/* paired-end part: PERelation symbols of the pairs with both mates mapped, and PEInsrNum (membuf+0x28) */
void sqo_set_align_pe(sqo_ctx *c, const uint8_t *perel, uint32_t n_perel, uint32_t insbits) {
    c->e_perel = perel; c->e_nperel = n_perel; c->e_insbits = insbits;
}
static int put_empty_rc(uint8_t *p, int tag) {
    int n = put_id(p, (uint64_t)tag);
    put_size(p + n, 8, 4);
    memset(p + n + 4, 0, 8);
    return n + 12;
}
uint32_t sqo_tested_reads(uint32_t nreads, int pe) {
    uint32_t t = (uint32_t)(int)((double)nreads * 0.05);
    uint32_t c = pe ? 2 * (t / 2 + 1) : t + 1;
    return c < nreads ? c : nreads;
}

/* ------------------------------------------------------------------ doFqzEncode @0x42d2d0 */
int64_t sqo_encode_block(sqo_ctx *c, const sqo_block *b, uint8_t *out, uint64_t cap) {
    (void)cap;
    uint8_t md5[3][16];
    dege_t d;
    /* caller side (doTask @0x433dd0): calcBlockMd5 @0x414d90 then DegeInfoProcess, on the ORIGINAL qualities */
    if (c->p.md5) { sqo_md5(b->ids, b->idbytes, md5[0]); sqo_md5(b->seq, b->seqbytes, md5[1]); sqo_md5(b->qual, b->seqbytes, md5[2]); }
    dege_build(b, &d);

    uint8_t *p = out;
    p += put_id(p, 1);
    uint8_t *blkslot = p; p += 4;
    /* compressCount(1, nreads) @0x422a00: tag 1, 1-byte size 0x84, u32le */
    p += put_id(p, 1); *p++ = 0x84; put_u32le(p, b->nreads); p += 4;
    /* Len: tag 4 */
    { p += put_id(p, 4); uint8_t *s = p; p += 4; int64_t n = sqo_stream_len(c, b, p); put_size(s, n, 4); p += n; }
    const int aligned = c->p.ref_mode && c->e_align > 0;              /* this block holds mapped reads */
    if (aligned) {                                                    /* every read was handed to the aligner */
        p += put_count(p, 27, b->nreads);
        p += put_side(p, 8, 5, c->e_order, b->nreads);                /* compressOrder @0x424b70 */
    } else if (c->p.ref_mode) {                                       /* @0x42d547-0x42d56f */
        p += put_count(p, 27, sqo_tested_reads(b->nreads, c->p.ref_mode == 2));
        p += put_empty_rc(p, 8);
    }
    /* ID: tag 5 (compressID @0x4247c0) */
    {
        p += put_id(p, 5); uint8_t *s = p; p += 4; uint8_t *st = p;
        if (c->p.md5) { memcpy(p, md5[0], 16); p += 16; }
        if (c->p.id_rule) {   /* IDProcess::encodeIDS @0x430040: u16le len + first ID only */
            uint16_t l = b->nreads ? b->idlen[0] : 0;
            p[0] = l & 0xff; p[1] = l >> 8; memcpy(p + 2, b->ids, l); p += 2 + l;
        } else p += sqo_stream_name(c, b, p);
        put_size(s, p - st, 4);
    }
    /* Qual: tag 7 */
    {
        p += put_id(p, 7); uint8_t *s = p; p += 4; uint8_t *st = p;
        if (c->p.md5 && !c->p.lossy) { memcpy(p, md5[2], 16); p += 16; }
        p += sqo_stream_qual(c, b, p);
        put_size(s, p - st, 4);
    }
    if (aligned) {
        p += put_count(p, 17, c->e_align); p += put_count(p, 18, c->e_posbits);
        p += put_count(p, 19, c->e_ncigal); p += put_count(p, 20, c->e_ncigav);
        if (c->p.ref_mode == 2) {
            p += put_count(p, 21, c->e_insbits); p += put_count(p, 22, c->e_nperel);
            p += put_side(p, 9, 4, c->e_perel, c->e_nperel);          /* compressPERelation @0x422be0 */
        }
        p += put_side(p, 11, 2, c->e_pos, c->e_posbits);              /* compressAlignInfo_Pos @0x425d70 */
        p += put_side(p, 15, 8, c->e_mis, c->e_align);                /* ..._Mis @0x425ff0 (N = 9 when SeqArcParam+0x1b60 == 8) */
        p += put_side(p, 10, 2, c->e_rev, c->e_align);                /* ..._Rev @0x426480 */
        p += put_side(p, 12, 2, c->e_cigal, c->e_ncigal);             /* ..._CigaL @0x426700 */
        p += put_side(p, 13, 4, c->e_cigav, c->e_ncigav);             /* ..._CigaV @0x426980 */
    } else if (c->p.ref_mode) {                                       /* @0x42d5bd-0x42d6c7, PE part @0x42d830 */
        for (int t = 17; t <= 20; t++) p += put_count(p, t, 0);
        if (c->p.ref_mode == 2) {
            p += put_count(p, 21, 9);                                 /* membuf+0x28: 9 on every PE archive observed */
            p += put_count(p, 22, 0);
            p += put_empty_rc(p, 9);                                  /* compressPERelation @0x422be0 */
        }
        static const int tags[5] = {11, 15, 10, 12, 13};              /* Pos, Mis, Rev, CigaL, CigaV */
        for (int k = 0; k < 5; k++) p += put_empty_rc(p, tags[k]);
    }
    kmodel_init(c);                                                   /* @0x42d391 */
    p += emit_small_sm(p, 23, 2, d.tip, d.n_tip, 0);
    p += emit_small_sm(p, 14, 11, d.ch, d.n_ch, 1);
    p += emit_small_sm(p, 24, 95, d.maxq, d.n_maxq, 2);
    p += emit_kmodel(c, p, 25, d.cnt, d.n_cnt);
    p += emit_kmodel(c, p, 26, d.pos, d.n_pos);
    /* Seq: tag 6 */
    {
        p += put_id(p, 6); uint8_t *s = p; p += 4; uint8_t *st = p;
        if (c->p.md5) { memcpy(p, md5[1], 16); p += 16; }
        if (aligned) {                                                /* only the reads the aligner left unmapped are coded */
            RC rc; rc_start_enc(&rc, p);
            reset_seq_model(c);
            const uint8_t *q = b->seq;
            for (uint32_t i = 0; i < b->nreads; i++) { if (!c->e_order[i]) encode_seq(c, &rc, q, b->seqlen[i], 0); q += b->seqlen[i]; }
            rc_finish_enc(&rc);
            p = rc.out;
        } else
            p += sqo_stream_seq(c, b, p);
        put_size(s, p - st, 4);
    }
    put_size(blkslot, (uint64_t)(p - blkslot - 4), 4);
    dege_free(&d);
    return p - out;
}

/* ------------------------------------------------------------------ training (usemodel)
 * encodeForModel @0x4228e0: the *_formodel twins update the models without coding and WITHOUT
 * resetting between the (<=5) head blocks; SaveModelToMem @0x420c00 then snapshots them. */
void sqo_train_block(sqo_ctx *c, const sqo_block *b, int first) {
    if (first) {
        int um = c->p.usemodel; c->p.usemodel = 0;
        reset_name_models(c); reset_seq_model(c); reset_qual_models(c);
        c->p.usemodel = um;
    }
    const uint8_t *p = b->ids, *s = b->seq, *q = b->qual;
    for (uint32_t i = 0; i < b->nreads; i++) {
        if (!c->p.id_rule) encode_name(c, NULL, p, b->idlen[i], 1);
        encode_seq(c, NULL, s, b->seqlen[i], 1);
        encode_qual(c, NULL, q, b->seqlen[i], 1);
        p += b->idlen[i]; s += b->seqlen[i]; q += b->seqlen[i];
    }
}
uint64_t sqo_save_model(sqo_ctx *c, uint8_t *out) {
    uint8_t *s = out;
    if (!c->p.id_rule) {
        for (int i = 0; i < 256; i++) s = sm_pack(c->name_prefix + i * SM_STRIDE(256), 256, s);
        for (int i = 0; i < 256; i++) s = sm_pack(c->name_suffix + i * SM_STRIDE(256), 256, s);
        for (int i = 0; i < 256; i++) s = sm_pack(c->name_len + i * SM_STRIDE(256), 256, s);
        for (int i = 0; i < 8192; i++) s = sm_pack(c->name_middle + i * SM_STRIDE(128), 128, s);
    }
    memcpy(s, c->seqtab, (size_t)c->nseq * 4); s += (size_t)c->nseq * 4;
    for (uint32_t i = 0; i < c->nqual; i++) s = sm_pack(c->qual + (size_t)i * SM_STRIDE(95), 95, s);
    return (uint64_t)(s - out);
}

/* test hook: the coding triples of one stream (0 seq, 1 qual, 2 name) in stream order */
int64_t sqo_trace_stream(sqo_ctx *c, const sqo_block *b, int which, uint64_t *triples, uint8_t *scratch) {
    g_trace = triples; g_ntrace = 0;
    if (which == 0) sqo_stream_seq(c, b, scratch);
    else if (which == 1) sqo_stream_qual(c, b, scratch);
    else sqo_stream_name(c, b, scratch);
    g_trace = NULL;
    return (int64_t)g_ntrace;
}

#include "seqarc_oracle_dec.inc"
#include "seqarc_oracle_ids.inc"
