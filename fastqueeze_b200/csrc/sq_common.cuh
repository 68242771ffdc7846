// sq_common.cuh -- shared types and device helpers of the B200 SeqArc entropy path.
//
// Reference behaviour is cited as @0xADDR = virtual address in /root/reference/SeqArc-1.6
// (the reference ships as an unstripped binary only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

#define SQ_HD __host__ __device__ __forceinline__
#define SQ_D __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Streams of one block that go through the context-model + range-coder pipeline.
enum { ST_SEQ = 0, ST_QUAL = 1, ST_NAME = 2, ST_COUNT = 3 };

// Sort element: [63:40] context key | [39:32] symbol | [31:0] block-relative symbol position.
SQ_HD u64 el_make(u32 key, u32 sym, u32 pos) { return ((u64)key << 40) | ((u64)(sym & 0xff) << 32) | pos; }
SQ_HD u32 el_key(u64 e) { return (u32)(e >> 40); }
SQ_HD u32 el_sym(u64 e) { return (u32)(e >> 32) & 0xff; }
SQ_HD u32 el_pos(u64 e) { return (u32)e; }

// Coding triple handed from the model replay to the range coder: cum | freq<<16 | tot<<32.
SQ_HD u64 tr_make(u32 cum, u32 freq, u32 tot) { return (u64)cum | ((u64)freq << 16) | ((u64)tot << 32); }
// The base model's totals never exceed 253 when a base is coded (@0x421fb0), so its triples are
// packed in 32 bits (10-bit fields): half the traffic and footprint of the largest triple array.
SQ_HD u32 tr_make32(u32 cum, u32 freq, u32 tot) { return cum | (freq << 10) | (tot << 20); }
SQ_HD u64 tr_expand32(u32 t) { return tr_make(t & 0x3ffu, (t >> 10) & 0x3ffu, t >> 20); }

// Name-model key space (one sort key space for the four fqzcomp name models, @0x421070):
// prefix[256] | suffix[256] | len[256] | middle[8192].
enum { NK_PREFIX = 0, NK_SUFFIX = 256, NK_LEN = 512, NK_MIDDLE = 768, NK_TOTAL = 768 + 8192 };

// Per-block descriptor of one wave, resident in device memory.
struct BlockDesc {
    const u8 *ids, *seq, *qual;       // SeqArcMemBuf +0x80 / +0x88 / +0x90
    const u16 *idlen;                 // +0xc0
    const i32 *seqlen;                // +0xd0
    u8 *qual_work;                    // qualities the coder sees (= qual, or the R-Block smoothed copy)
    u32 nreads;
    i32 longflag;                     // membuf+0x2: the block holds a read > 65,535 bases (compressLen_long @0x423710)
    u32 rbase;                        // first read of this block in the wave-wide per-read arrays
    u64 idbytes, seqbytes;
    u64 ebase[ST_COUNT];              // first element of this block in the element / triple arrays
    u64 ebound[ST_COUNT];             // capacity (upper bound on coded symbols) per stream
    u64 obase[ST_COUNT];              // first byte of this block's range-coder scratch per stream
    u64 ocap[ST_COUNT];               // its capacity
    u64 sidebase, sidecap;            // side-stream scratch (Len, Dege*) of this block
    u8 *rec;                          // where the packed record goes (device)
    u64 reccap;
};

// Per-block results written by the kernels.
struct BlockState {
    u32 nsym[ST_COUNT];               // coded symbols per stream
    u32 olen[ST_COUNT];               // range-coder bytes per stream
    u32 side_len[6];                  // Len, DegeTip, DegeCh, DegeMaxQual, NDegeCnt, NDegePos byte counts
    u32 side_cnt[6];                  // their u32 counts (0 = stream omitted)
    u32 acgt;                         // coded bases
    u32 reclen;                       // packed record bytes
    i32 status;
    u32 pad;
    u8 md5[3][16];                    // ids, seq, qual (calcBlockMd5 @0x414d90)
};

// ---------------------------------------------------------------------------------------------
// seq_val_table @0x44b800: A/a 0, C/c 1, G/g 2, T/t 3, then N M R Y K S W H B V D = 4..14; else 4.
SQ_HD u32 seq_val(u32 c) {
    switch (c | 0x20) {
        case 'a': return 0; case 'c': return 1; case 'g': return 2; case 't': return 3;
        case 'n': return 4; case 'm': return 5; case 'r': return 6; case 'y': return 7;
        case 'k': return 8; case 's': return 9; case 'w': return 10; case 'h': return 11;
        case 'b': return 12; case 'v': return 13; case 'd': return 14;
        default: return 4;
    }
}

// ---------------------------------------------------------------------------------------------
// Serial range coder (RangeCoder, inlined in the reference: encode @0x422010-0x422085, flush
// @0x424a30, decode @0x429760-0x429880).  64-bit low, 32-bit range, carry-less renormalisation.
struct RcEnc {
    u64 low;
    u32 range;
    u8 *out;
    u8 *end;
    int overflow;
    SQ_HD void start(u8 *o, u8 *e) { low = 0; range = 0xffffffffu; out = o; end = e; overflow = 0; }
    SQ_HD void put(u8 b) { if (out < end) *out++ = b; else overflow = 1; }
    SQ_HD void renorm() {
        while (range < (1u << 24)) {
            if ((low ^ (low + range)) >> 56) range = (((u32)low) | 0xffffffu) - (u32)low;
            put((u8)(low >> 56));
            range <<= 8;
            low <<= 8;
        }
    }
    SQ_HD void encode(u32 cum, u32 freq, u32 tot) {
        range /= tot;
        low += (u32)(cum * range);
        range *= freq;
        renorm();
    }
    SQ_HD void finish() { for (int i = 0; i < 8; i++) { put((u8)(low >> 56)); low <<= 8; } }
};

struct RcDec {
    u64 low;
    u32 range, code;
    const u8 *in;
    const u8 *end;
    SQ_HD u8 get() { return in < end ? *in++ : 0; }
    SQ_HD void start(const u8 *i, const u8 *e) {
        low = 0; range = 0xffffffffu; code = 0; in = i; end = e;
        for (int k = 0; k < 8; k++) code = (code << 8) | get();
    }
    SQ_HD u32 getfreq(u32 tot) { range /= tot; return code / range; }
    SQ_HD void decode(u32 cum, u32 freq) {
        u32 t = cum * range;
        low += t;
        code -= t;
        range *= freq;
        while (range < (1u << 24)) {
            if ((low ^ (low + range)) >> 56) range = (((u32)low) | 0xffffffu) - (u32)low;
            code = (code << 8) | get();
            range <<= 8;
            low <<= 8;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// SIMPLE_MODEL<N> (ctor pattern @0x42f2b0, encodeSymbol inlined e.g. @0x4221fb-0x4222f1) on a
// plain word array: F[i] = sym | freq<<16.  tot and bub live beside it.  Used by the serial
// side-stream coders and the decoders; the bulk replay kernels keep F in shared memory instead.
template <int N>
struct SmModel {
    u32 tot, bub;
    u32 F[N];
    SQ_HD void init() {
        tot = N; bub = 0;
        for (int i = 0; i < N; i++) F[i] = (u32)i | (1u << 16);
    }
    // update after coding the symbol at index k (halving @0x424fc0, bubble swap @0x4222d3)
    SQ_HD void update(int k) {
        u32 f = (F[k] >> 16) + 8;
        F[k] = (F[k] & 0xffffu) | (f << 16);
        tot += 8;
        if (tot > 0xffe0u) {
            tot = 0;
            for (int j = 0; j < N; j++) {
                u32 g = F[j] >> 16;
                g -= g >> 1;
                F[j] = (F[j] & 0xffffu) | (g << 16);
                tot += g;
            }
        }
        if (((++bub) & 15) == 0 && k > 0 && (F[k] >> 16) > (F[k - 1] >> 16)) {
            u32 t = F[k]; F[k] = F[k - 1]; F[k - 1] = t;
        }
    }
    SQ_HD void encode(RcEnc &rc, u32 sym) {
        u32 cum = 0;
        int k = 0;
        while (k < N - 1 && (F[k] & 0xffffu) != sym) { cum += F[k] >> 16; k++; }
        rc.encode(cum, F[k] >> 16, tot);
        update(k);
    }
    SQ_HD u32 decode(RcDec &rc) {
        u32 target = rc.getfreq(tot), cum = 0;
        int k = 0;
        while (k < N - 1 && cum + (F[k] >> 16) <= target) { cum += F[k] >> 16; k++; }
        rc.decode(cum, F[k] >> 16);
        u32 s = F[k] & 0xffffu;
        update(k);
        return s;
    }
    // snapshot form (SaveModelToMem @0x420c00): u32le tot, u32le bub, N x {u8 sym, u16le freq}
    SQ_HD void unpack(const u8 *s) {
        tot = s[0] | (s[1] << 8) | (s[2] << 16) | ((u32)s[3] << 24);
        bub = s[4] | (s[5] << 8) | (s[6] << 16) | ((u32)s[7] << 24);
        s += 8;
        for (int i = 0; i < N; i++, s += 3) F[i] = (u32)s[0] | ((u32)(s[1] | (s[2] << 8)) << 16);
    }
};

// ---------------------------------------------------------------------------------------------
// Encap vints (Encap::setID @0x420720, setSize @0x420780)
SQ_HD int put_tag(u8 *p, u32 id) { p[0] = (u8)(id | 0x80); return 1; }   // every tag used here is < 127
SQ_HD void put_size4(u8 *p, u32 size) {
    u32 v = size | (1u << 28);
    p[0] = (u8)(v >> 24); p[1] = (u8)(v >> 16); p[2] = (u8)(v >> 8); p[3] = (u8)v;
}
SQ_HD void put_u32le(u8 *p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }
SQ_HD u32 get_u32le(const u8 *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((u32)p[3] << 24); }

// quality context update of encode_qual @0x422474-0x422505 (shared by encoder and decoder)
SQ_HD u32 qual_ctx(u32 q, u32 q1, u32 q2, u32 delta, int qlevel, int i) {
    u32 last = (((q1 > q2 ? q1 : q2) << 6) + q) & 0xfff;
    if (qlevel > 1) {
        last += (u32)(q1 == q2) << 12;
        last += ((delta < 56 ? delta : 56) & 0xf8) << 10;
    }
    if (qlevel > 2) last += (i <= 111) ? (u32)((i + 15) & 0x78) << 13 : 0xf0000u;
    return last;
}
