// sq_decode.cu -- CUDA kernels of the block decoder (EncapFqzComp::doDecode @0x42eb60 /
// doRealDecode @0x42e740 and the decompress* twins).
//
// On DECODE every context depends on the symbols decoded so far, so each (block, stream) is one
// strictly serial chain; parallelism is blocks x {name, qual, seq}.  One CTA per
// block: warp 0 decodes the names, warp 1 the qualities, warp 2 the Dege side streams and then the
// bases.  The base chain only needs the qualities of reads that hold non-ACGT bases (N placement,
// decode_seq @0x4296b0), so it runs concurrently with the quality chain and waits on a progress
// counter only for those reads.  Model tables live in device memory (4 MiB + 25 MB + 5 MB per
// block): the chains are L2/HBM-latency bound by construction.
#include "sq_common.cuh"
#include "sq_kernels.h"
#include "sq_md5.cuh"

#define FULL 0xffffffffu
#define QSTRIDE 97     // words per SIMPLE_MODEL<95> in qualtab  (tot, bub, F[95])
#define N256STRIDE 258
#define N128STRIDE 130

typedef SmModel<95> QModel;
typedef SmModel<256> M256;
typedef SmModel<128> M128;
static_assert(sizeof(QModel) == QSTRIDE * 4, "layout");
static_assert(sizeof(M256) == N256STRIDE * 4, "layout");
static_assert(sizeof(M128) == N128STRIDE * 4, "layout");

// ---------------------------------------------------------------------------------------------
// model table initialisation (per block; compressQual @0x426f10-0x426f84 etc. reset every block)
template <int N>
__global__ void k_init_models(DecBlock *db, int which, u32 nmodels, u32 wordoff, const u8 *snap, u32 snapstride) {
    int b = blockIdx.y;
    if (db[b].status != 0) return;
    u32 *tab = (which == 0 ? db[b].qualtab : db[b].nametab) + wordoff;
    constexpr int STRIDE = N + 2;
    if (snap) {
        u32 m = blockIdx.x * blockDim.x + threadIdx.x;
        if (m >= nmodels) return;
        ((SmModel<N> *)(tab + (u64)m * STRIDE))->unpack(snap + (u64)m * snapstride);
        return;
    }
    u64 total = (u64)nmodels * STRIDE;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        u32 w = (u32)(i % STRIDE);
        tab[i] = w == 0 ? (u32)N : w == 1 ? 0u : ((w - 2) | (1u << 16));
    }
}
__global__ void k_init_seqtab(DecBlock *db, u64 bytes, const u8 *snap) {
    int b = blockIdx.y;
    if (db[b].status != 0) return;
    u32 *t = (u32 *)db[b].seqtab;
    const u32 *s = (const u32 *)snap;
    u64 n = bytes / 4;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x)
        t[i] = s ? s[i] : 0x03030303u;
}

// ---------------------------------------------------------------------------------------------
// Len stream (decompressLen_short @0x428170 / decode_len_short @0x427ca0) + read offsets.
// The reference keeps last_len across blocks per decoder instance; a zero-length
// read therefore decodes to "whatever came before".  Here last_len starts at 0 for every block.
struct LenSmem { M256 len1, len2; SmModel<2> same; };
__global__ void __launch_bounds__(32) k_dec_len(DecBlock *db) {
    __shared__ LenSmem sm;
    DecBlock &d = db[blockIdx.x];
    if (threadIdx.x != 0 || d.status != 0) return;
    if (d.nreads > d.readcap) { d.status = SQ_ERR_CAPACITY_K; return; }
    if (d.len[4] < 8 && d.nreads) { d.status = SQ_ERR_FORMAT_K; return; }
    RcDec rc;
    rc.start(d.rec + d.off[4], d.rec + d.off[4] + d.len[4]);
    sm.len1.init(); sm.len2.init(); sm.same.init();
    i32 last = 0;
    u64 tot = 0;
    for (u32 r = 0; r < d.nreads; r++) {
        if (sm.same.decode(rc) == 0) {
            u32 l1 = sm.len1.decode(rc), l2 = sm.len2.decode(rc);
            last = (i32)(l1 | (l2 << 8));
        }
        d.seqlen[r] = last;
        d.seqoff[r] = (u32)tot;
        tot += (u32)last;
        if (tot > d.seqcap) { d.status = SQ_ERR_CAPACITY_K; return; }
    }
    d.seqoff[d.nreads] = (u32)tot;
    d.seqbytes = tot;
}

// ---------------------------------------------------------------------------------------------
// decode_name @0x428380
SQ_D int decode_name(DecBlock &d, RcDec &rc, u8 *name, u64 room, u8 *last, int &lnl, int &lpl, int &lsl) {
    M256 *prefix = (M256 *)d.nametab, *suffix = prefix + 256, *lenm = suffix + 256;
    M128 *middle = (M128 *)(lenm + 256);
    int p = (int)prefix[lpl].decode(rc);
    int s = (int)suffix[lsl].decode(rc);
    int n = (int)lenm[lnl].decode(rc);
    lpl = p; lsl = s;
    if ((u64)n > room) return -1;                                          // caller's ID buffer is full
    if (p > n || s > n - p || s > lnl) { p = p > n ? n : p; s = 0; }       // corrupt stream guard
    for (int i = 0; i < p; i++) name[i] = last[i];
    int len2 = n - s, lc2 = (p != 0), i = p, j = p, k = 0;
    while (i < len2) {
        u32 lj = j < 0 ? 0 : last[j];
        int v = ((int)lj - 32) * 2 + lc2 + k * 64;
        int ctx = v % 8192;
        if (ctx < 0) ctx = 0;
        u32 ch = middle[ctx].decode(rc);
        name[i] = (u8)ch;
#define LJ (j < 0 ? 0u : (u32)last[j])
        if (ch == ' ' && LJ != ' ') j++;
        if (ch != ' ' && LJ == ' ') j--;
        if (ch == ':' && LJ != ':') j++;
        if (ch != ':' && LJ == ':') j--;
        if (ch == ':' || ch == ' ') k = (k + 3) & ~3;
        if (j < 0) { j = 0; lc2 = 0; }
        else { lc2 = (ch == LJ); j++; }
#undef LJ
        k++; i++;
    }
    for (int t = 0; t < s; t++) name[len2 + t] = last[lnl - s + t];
    for (int t = 0; t < n; t++) last[t] = name[t];
    lnl = n;
    return n;
}

// decode_qual @0x42a750
SQ_D void decode_qual(DecBlock &d, RcDec &rc, u8 *q, int len, int qlevel) {
    QModel *tab = (QModel *)d.qualtab;
    u32 last = 0, delta = 5, q1 = 0, q2 = 0;
    for (int i = 0; i < len; i++) {
        u32 qq = tab[last].decode(rc) & 0xff;
        if (qq == 94) { for (; i < len; i++) q[i] = '#'; return; }
        q[i] = (u8)(qq + 33);
        if (q1 > qq) delta += q1 - qq;
        last = qual_ctx(qq, q1, q2, delta, qlevel, i);
        q2 = q1; q1 = qq;
    }
}

// one base (inlined four times in decode_seq @0x4296b0)
SQ_D u32 decode_base(u8 *seqtab, RcDec &rc, u32 &last, u32 mask) {
    u32 *S4 = (u32 *)(seqtab + (u64)last * 4);
    u32 x = *S4;
    u32 S0 = x & 0xff, S1 = (x >> 8) & 0xff, S2 = (x >> 16) & 0xff, S3 = x >> 24;
    u32 tot = S0 + S1 + S2 + S3;
    if (tot > 253) { S0 -= S0 >> 1; S1 -= S1 >> 1; S2 -= S2 >> 1; S3 -= S3 >> 1; tot = S0 + S1 + S2 + S3; }
    u32 target = rc.getfreq(tot);
    u32 b, cum, f;
    if (target < S0) { b = 0; cum = 0; f = S0; }
    else if (target < S0 + S1) { b = 1; cum = S0; f = S1; }
    else if (target < S0 + S1 + S2) { b = 2; cum = S0 + S1; f = S2; }
    else { b = 3; cum = S0 + S1 + S2; f = S3; }
    rc.decode(cum, f);
    if (b == 0) S0 = (S0 + 1) & 0xff; else if (b == 1) S1 = (S1 + 1) & 0xff; else if (b == 2) S2 = (S2 + 1) & 0xff; else S3 = (S3 + 1) & 0xff;
    *S4 = S0 | (S1 << 8) | (S2 << 16) | (S3 << 24);
    last = (last * 4 + b) & mask;
    return b;
}

// small side streams: "u32le count | range-coder bytes" with a fresh local model
template <int N>
SQ_D u32 dec_small(const DecBlock &d, int tag, SmModel<N> &m, u8 *arr, u32 cap, int mode) {
    if (!d.len[tag]) return 0;
    const u8 *pay = d.rec + d.off[tag];
    u32 cnt = get_u32le(pay);
    if (cnt > cap) return 0xffffffffu;
    m.init();
    RcDec rc;
    rc.start(pay + 4, pay + d.len[tag]);
    const char tbl[] = "ACGTNMRYKSWHBVD";
    for (u32 i = 0; i < cnt; i++) {
        u32 s = m.decode(rc);
        if (mode == 1) s = (u32)tbl[(s + 4) & 15];
        else if (mode == 2) s = s + 33;
        arr[i] = (u8)s;
    }
    return cnt;
}
struct KModelD {
    SmModel<64> nb;
    SmModel<2> bits[64];
    SQ_D void init() { nb.init(); for (int i = 0; i < 64; i++) bits[i].init(); }
    SQ_D u32 decode(RcDec &rc) {
        int n = (int)nb.decode(rc);
        u32 v = 0;
        for (int i = 0; i < n; i++) { u32 bit = bits[i].decode(rc); if (i < 32) v |= bit << i; }
        return v;
    }
};
struct MainSmem {
    u8 last_name[1024];
    SmModel<2> tip;
    SmModel<11> ch;
    SmModel<95> maxq;
    KModelD km;
    volatile u32 qprog;        // reads whose qualities are complete
};

__global__ void __launch_bounds__(96) k_dec_main(DecBlock *db, int qlevel, int id_rule, u32 seqmask) {
    __shared__ MainSmem sm;
    DecBlock &d = db[blockIdx.x];
    int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) sm.qprog = 0;
    __syncthreads();
    if (d.status != 0 || lane != 0) return;
    const char tbl[] = "ACGT";
    if (wid == 0) {                                            // names (decompressName @0x428ca0)
        if (id_rule) { d.status = SQ_ERR_FORMAT_K; return; }   // rule-based IDs are regenerated on the host
        for (int i = 0; i < 1024; i++) sm.last_name[i] = ' ';
        int lnl = 0, lpl = 0, lsl = 0;
        RcDec rc;
        rc.start(d.rec + d.off[5], d.rec + d.off[5] + d.len[5]);
        u64 o = 0;
        for (u32 r = 0; r < d.nreads; r++) {
            int n = decode_name(d, rc, d.ids + o, d.idcap - o, sm.last_name, lnl, lpl, lsl);
            if (n < 0) { d.status = SQ_ERR_CAPACITY_K; return; }
            d.idlen[r] = (u16)n;
            o += (u32)n;
        }
        d.idbytes = o;
    } else if (wid == 1) {                                     // qualities (decompressQual @0x42aa40)
        RcDec rc;
        rc.start(d.rec + d.off[7], d.rec + d.off[7] + d.len[7]);
        for (u32 r = 0; r < d.nreads; r++) {
            decode_qual(d, rc, d.qual + d.seqoff[r], d.seqlen[r], qlevel);
            if ((r & 15) == 15 || r + 1 == d.nreads) { __threadfence_block(); sm.qprog = r + 1; }
        }
    } else {                                                   // Dege streams, then bases (decompressSeq @0x42e390)
        u32 n_tip = dec_small<2>(d, 23, sm.tip, d.d_tip, d.nreads, 0);
        u32 n_ch = dec_small<11>(d, 14, sm.ch, d.d_ch, (u32)d.seqbytes, 1);
        u32 n_mq = dec_small<95>(d, 24, sm.maxq, d.d_maxq, d.nreads, 2);
        if (n_tip == 0xffffffffu || n_ch == 0xffffffffu || n_mq == 0xffffffffu) { d.status = SQ_ERR_FORMAT_K; return; }
        sm.km.init();                                          // kModelInit @0x42ebd2, shared by tags 25 and 26
        u32 n_cnt = 0, n_pos = 0;
        for (int k = 0; k < 2; k++) {
            int tag = 25 + k;
            if (!d.len[tag]) continue;
            const u8 *pay = d.rec + d.off[tag];
            u32 cnt = get_u32le(pay);
            if (cnt > (k == 0 ? d.nreads : (u32)d.seqbytes)) { d.status = SQ_ERR_FORMAT_K; return; }
            RcDec rc;
            rc.start(pay + 4, pay + d.len[tag]);
            u32 *arr = k == 0 ? d.d_cnt : d.d_pos;
            for (u32 i = 0; i < cnt; i++) arr[i] = sm.km.decode(rc);
            if (k == 0) n_cnt = cnt; else n_pos = cnt;
        }
        RcDec rc;
        rc.start(d.rec + d.off[6], d.rec + d.off[6] + d.len[6]);
        u32 i_mq = 0, i_ch = 0, i_pos = 0;
        for (u32 r = 0; r < d.nreads; r++) {
            u8 *seq = d.seq + d.seqoff[r];
            int len = d.seqlen[r];
            u32 last = 0x7616c7u & seqmask;
            bool flagged = r < n_tip && d.d_tip[r];
            if (!flagged) {
                for (int i = 0; i < len; i++) seq[i] = (u8)tbl[decode_base(d.seqtab, rc, last, seqmask)];
                continue;
            }
            while (sm.qprog <= r) __nanosleep(200);            // this read's qualities place its N bases
            __threadfence_block();
            const u8 *qual = d.qual + d.seqoff[r];
            if (i_mq >= n_mq) { d.status = SQ_ERR_FORMAT_K; return; }
            int mq = (int)(int8_t)d.d_maxq[i_mq];
            u32 cnt = i_mq < n_cnt ? d.d_cnt[i_mq] : 0;
            i_mq++;
            if (cnt == 0) {                                    // @0x429a20: every candidate position is non-ACGT
                for (int i = 0; i < len; i++) {
                    if (mq >= (int)(int8_t)__ldcg(qual + i)) seq[i] = i_ch < n_ch ? d.d_ch[i_ch++] : 'N';
                    else seq[i] = (u8)tbl[decode_base(d.seqtab, rc, last, seqmask)];
                }
                continue;
            }
            u32 gap = i_pos < n_pos ? d.d_pos[i_pos++] : 0, remaining = cnt - 1, run = 0;   // @0x4298cc
            for (int i = 0; i < len; i++) {
                if (mq >= (int)(int8_t)__ldcg(qual + i)) {
                    if (run == gap) {                          // @0x429b40
                        gap = 0xffffffffu;
                        if (remaining) { remaining--; gap = i_pos < n_pos ? d.d_pos[i_pos++] : 0; }
                        seq[i] = (u8)tbl[decode_base(d.seqtab, rc, last, seqmask)];
                        run = 0;
                    } else { run++; seq[i] = i_ch < n_ch ? d.d_ch[i_ch++] : 'N'; }
                } else seq[i] = (u8)tbl[decode_base(d.seqtab, rc, last, seqmask)];
            }
        }
    }
}

// blockMd5Verify @0x414e00 ("Name/Seq/Qual md5 unequal", .rodata+0x15bc)
__global__ void __launch_bounds__(32) k_dec_md5(DecBlock *db, int B) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * B) return;
    int b = t / 3, which = t % 3;
    DecBlock &d = db[b];
    if (d.status != 0 || !d.has_md5[which]) return;
    const u8 *p = which == 0 ? d.ids : which == 1 ? d.seq : d.qual;
    u64 n = which == 0 ? d.idbytes : d.seqbytes;
    u8 dig[16];
    md5_buffer(p, n, dig);
    bool ok = true;
    for (int k = 0; k < 16; k++) ok = ok && (dig[k] == d.md5[which][k]);
    if (!ok) d.status = -8 - which;      // SQ_ERR_MD5_NAME / _SEQ / _QUAL
}

static inline u32 cdivd(u64 a, u64 b) { return (u32)((a + b - 1) / b); }

cudaError_t sqk_decode(const DecWave &w, cudaStream_t st, u64 *launches) {
    int B = w.B;
    if (B == 0) return cudaSuccess;
    // per-block model reset (or snapshot restore)
    if (w.snap_qual) k_init_models<95><<<dim3(cdivd(w.nqual, 128), B), 128, 0, st>>>(w.db, 0, w.nqual, 0, w.snap_qual, 293);
    else k_init_models<95><<<dim3(w.nsm * 2, B), 256, 0, st>>>(w.db, 0, w.nqual, 0, nullptr, 0);
    k_init_seqtab<<<dim3(w.nsm, B), 256, 0, st>>>(w.db, (u64)w.nseq * 4, w.snap_seq);
    *launches += 2;
    if (!w.id_rule) {
        // nametab: 768 x SIMPLE_MODEL<256> then 8192 x SIMPLE_MODEL<128>
        if (w.snap_name) {
            k_init_models<256><<<dim3(cdivd(768, 128), B), 128, 0, st>>>(w.db, 1, 768, 0, w.snap_name, 776);
            k_init_models<128><<<dim3(cdivd(8192, 128), B), 128, 0, st>>>(w.db, 1, 8192, 768 * N256STRIDE, w.snap_name + 3ull * 256 * 776, 392);
        } else {
            k_init_models<256><<<dim3(8, B), 256, 0, st>>>(w.db, 1, 768, 0, nullptr, 0);
            k_init_models<128><<<dim3(32, B), 256, 0, st>>>(w.db, 1, 8192, 768 * N256STRIDE, nullptr, 0);
        }
        *launches += 2;
    }
    k_dec_len<<<B, 32, 0, st>>>(w.db);
    k_dec_main<<<B, 96, 0, st>>>(w.db, w.qlevel, w.id_rule, w.seqmask);
    *launches += 2;
    if (w.md5) { k_dec_md5<<<cdivd(3ull * B, 32), 32, 0, st>>>(w.db, B); *launches += 1; }
    return cudaGetLastError();
}
