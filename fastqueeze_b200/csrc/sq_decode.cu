// sq_decode.cu -- CUDA kernels of the block decoder (EncapFqzComp::doDecode @0x42eb60 /
// doRealDecode @0x42e740 and the decompress* twins).
//
// On DECODE every context depends on the symbols decoded so far, so each (block, stream) is one
// strictly serial chain; parallelism is blocks x {name, qual, seq}.  One CTA per
// block: warp 0 decodes the names, warp 1 the qualities, warp 2 the Dege side streams and then the
// bases.  The SIMPLE_MODELs are "warp models" (see WM below): the 32 lanes of the decoding warp
// hold one model, so a symbol costs one coalesced fetch, one division and one ballot.  The base chain only needs the qualities of reads that hold non-ACGT bases (N placement,
// decode_seq @0x4296b0), so it runs concurrently with the quality chain and waits on a progress
// counter only for those reads.  Model tables live in device memory (4 MiB + 25 MB + 5 MB per
// block): the chains are L2/HBM-latency bound by construction.
#include "sq_common.cuh"
#include "sq_kernels.h"
#include "sq_md5.cuh"

#define FULL 0xffffffffu

// ---------------------------------------------------------------------------------------------
// Decoder-private model layout ("warp model").  A SIMPLE_MODEL<N> is spread over the 32 lanes of
// the decoding warp, EPL consecutive entries per lane (lane l owns F[l*EPL .. l*EPL+EPL)):
//   row i   (i < EPL), word l : entry i of lane l = sym | bub << 8 (row 0 only) | freq << 16
//   row EPL,           word l : base_l | tot << 16, base_l = sum of the frequencies of all lower lanes
// Every word is only ever read and written by its own lane, every row is one coalesced 128-byte
// access, and a symbol is found with EPL multiply-compares per lane + one ballot (no second division:
// target = code / r  is replaced by comparing code with cum * r, which is the same test on integers).
template <int N, int EPL> struct WM {
    static constexpr int ROWS = EPL + 1;
    static constexpr int STRIDE = ROWS * 32;          // words per model
};
#define WQ_STRIDE (WM<95, 3>::STRIDE)
#define W256_STRIDE (WM<256, 8>::STRIDE)
#define W128_STRIDE (WM<128, 4>::STRIDE)

SQ_D u32 warp_incl_scan_d(u32 v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}

// per-block model reset (compressQual @0x426f10-0x426f84 etc. reset every block) or snapshot
// restore (unpack loops @0x42705c-0x42712e): one warp per model
template <int N, int EPL>
__global__ void __launch_bounds__(256) k_init_wm(DecBlock *db, int which, u32 nmodels, u32 wordoff, const u8 *snap, u32 snapstride) {
    int b = blockIdx.y;
    if (db[b].status != 0) return;
    u32 *tab = (which == 0 ? db[b].qualtab : db[b].nametab) + wordoff;
    int lane = threadIdx.x & 31;
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 m = warp; m < nmodels; m += nwarps) {
        u32 *mp = tab + (u64)m * WM<N, EPL>::STRIDE;
        u32 w[EPL], T = 0, tot = N, bub = 0;
        const u8 *s = snap ? snap + (u64)m * snapstride : nullptr;
        if (s) { tot = get_u32le(s); bub = get_u32le(s + 4) & 0xff; }
#pragma unroll
        for (int i = 0; i < EPL; i++) {
            int slot = lane * EPL + i;
            u32 sym = (u32)slot & 0xff, f = slot < N ? 1 : 0;
            if (s && slot < N) { const u8 *q = s + 8 + 3 * slot; sym = q[0]; f = (u32)(q[1] | (q[2] << 8)); }
            w[i] = sym | (f << 16);
            T += f;
        }
        u32 base = warp_incl_scan_d(T, lane) - T;
        w[0] |= bub << 8;
#pragma unroll
        for (int i = 0; i < EPL; i++) mp[i * 32 + lane] = w[i];
        mp[EPL * 32 + lane] = base | (tot << 16);
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

// decodeSymbol (inlined e.g. @0x42a7d0-0x42a960) on a warp model; every lane returns the symbol.
// rc is warp-uniform (all lanes carry the same coder state).
template <int N, int EPL>
SQ_D u32 wm_decode(u32 *mp, RcDec &rc, int lane) {
    u32 w[EPL];
#pragma unroll
    for (int i = 0; i < EPL; i++) w[i] = mp[i * 32 + lane];
    u32 bt = mp[EPL * 32 + lane];
    u32 base = bt & 0xffffu, tot = bt >> 16;
    u32 bub = (w[0] >> 8) & 0xff;
    w[0] &= 0xffff00ffu;
    rc.range /= tot;
    u32 r = rc.range, code = rc.code;
    // lane-local search: first entry whose inclusive cumulative frequency * r exceeds code
    int li = -1;
    u32 c = base, cex = 0;
#pragma unroll
    for (int i = 0; i < EPL; i++) {
        u32 cn = c + (w[i] >> 16);
        if (li < 0 && code < cn * r) { li = i; cex = c; }
        c = cn;
    }
    u32 bal = __ballot_sync(FULL, li >= 0);
    int kl;
    if (bal) kl = __ffs(bal) - 1;
    else {                                   // code beyond tot * r: corrupt stream; take the last symbol
        kl = (N - 1) / EPL;
        if (lane == kl) { li = (N - 1) % EPL; cex = c - (w[(N - 1) % EPL] >> 16); }
    }
    u32 wsel = 0;
#pragma unroll
    for (int i = 0; i < EPL; i++) if (i == li) wsel = w[i];
    wsel = __shfl_sync(FULL, wsel, kl);
    cex = __shfl_sync(FULL, cex, kl);
    u32 sym = wsel & 0xff, freq = wsel >> 16;
    rc.decode(cex, freq);
    // model update (Freq += 8, TotFreq += 8, halve above 0xffe0, bubble swap every 16th symbol)
    bool mine = lane == kl;
    if (mine) {
#pragma unroll
        for (int i = 0; i < EPL; i++) if (i == li) w[i] += 8u << 16;
    }
    if (lane > kl) base += 8;
    tot += 8;
    bool all = false;
    if (tot > 0xffe0u) {
        u32 T = 0;
#pragma unroll
        for (int i = 0; i < EPL; i++) { u32 f = w[i] >> 16; f -= f >> 1; w[i] = (w[i] & 0xffffu) | (f << 16); T += f; }
        u32 inc = warp_incl_scan_d(T, lane);
        base = inc - T;
        tot = __shfl_sync(FULL, inc, 31);
        all = true;
    }
    bub = (bub + 1) & 0xff;
    bool prevlane = false;
    if ((bub & 15) == 0) {
        u32 up = __shfl_up_sync(FULL, w[EPL - 1], 1);
        u32 cur = 0, prv = up;
#pragma unroll
        for (int i = 0; i < EPL; i++) {
            if (i == li) cur = w[i];
            if (i + 1 == li) prv = w[i];
        }
        bool dosw = mine && (kl * EPL + li > 0) && (cur >> 16) > (prv >> 16);
        u32 swb = __ballot_sync(FULL, dosw);
        if (swb) {
            int lik = __shfl_sync(FULL, li, kl);
            u32 cw = __shfl_sync(FULL, cur, kl), pw = __shfl_sync(FULL, prv, kl);
            if (lik > 0) {
                if (mine) {
#pragma unroll
                    for (int i = 1; i < EPL; i++) if (i == li) { w[i] = pw; w[i - 1] = cw; }
                }
            } else {
                if (mine) { w[0] = pw; base += (cw >> 16) - (pw >> 16); }
                if (lane == kl - 1) { w[EPL - 1] = cw; prevlane = true; }
            }
        }
    }
    // write back: row 0 carries bub and the base row carries tot, so both are rewritten by every lane
    mp[lane] = w[0] | (bub << 8);
    mp[EPL * 32 + lane] = base | (tot << 16);
    if (mine || prevlane || all) {
#pragma unroll
        for (int i = 1; i < EPL; i++) mp[i * 32 + lane] = w[i];
    }
    return sym;
}

// ---------------------------------------------------------------------------------------------
// Len stream (decompressLen_short @0x428170 / decode_len_short @0x427ca0) + read offsets.
// The reference keeps last_len across blocks per decoder instance; a zero-length
// read therefore decodes to "whatever came before".  Here last_len starts at 0 for every block.
struct LenSmem { SmModel<256> len1, len2; SmModel<2> same; };
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
// decode_name @0x428380 -- executed by the whole name warp with uniform control flow; the models
// are warp models, last_name / the name being built live in shared memory.
struct NameSmem { u8 last[1024]; u8 cur[256]; };

SQ_D int decode_name(DecBlock &d, RcDec &rc, u8 *name_out, u64 room, NameSmem &ns, int &lnl, int &lpl, int &lsl, int lane) {
    u32 *prefix = d.nametab, *suffix = prefix + 256 * W256_STRIDE, *lenm = suffix + 256 * W256_STRIDE;
    u32 *middle = lenm + 256 * W256_STRIDE;
    int p = (int)wm_decode<256, 8>(prefix + (u64)lpl * W256_STRIDE, rc, lane);
    int s = (int)wm_decode<256, 8>(suffix + (u64)lsl * W256_STRIDE, rc, lane);
    int n = (int)wm_decode<256, 8>(lenm + (u64)lnl * W256_STRIDE, rc, lane);
    lpl = p; lsl = s;
    if ((u64)n > room) return -1;                                          // caller's ID buffer is full
    if (p > n || s > n - p || s > lnl) { p = p > n ? n : p; s = 0; }       // corrupt stream guard
    u8 *last = ns.last, *cur = ns.cur;
    for (int i = lane; i < p; i += 32) cur[i] = last[i];
    int len2 = n - s, lc2 = (p != 0), i = p, j = p, k = 0;
    while (i < len2) {
        u32 lj = j < 0 ? 0 : last[j];
        int v = ((int)lj - 32) * 2 + lc2 + k * 64;
        int ctx = v % 8192;
        if (ctx < 0) ctx = 0;
        u32 ch = wm_decode<128, 4>(middle + (u64)ctx * W128_STRIDE, rc, lane);
        if (lane == 0) cur[i] = (u8)ch;
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
    for (int t = lane; t < s; t += 32) cur[len2 + t] = last[lnl - s + t];
    __syncwarp();
    for (int t = lane; t < n; t += 32) { u8 c = cur[t]; last[t] = c; name_out[t] = c; }
    __syncwarp();
    lnl = n;
    return n;
}

// decode_qual @0x42a750 (whole warp, uniform)
SQ_D void decode_qual(DecBlock &d, RcDec &rc, u8 *q, int len, int qlevel, int lane) {
    u32 *tab = d.qualtab;
    u32 last = 0, delta = 5, q1 = 0, q2 = 0;
    for (int i = 0; i < len; i++) {
        u32 qq = wm_decode<95, 3>(tab + (u64)last * WQ_STRIDE, rc, lane);
        if (qq == 94) { for (int t = i + lane; t < len; t += 32) q[t] = '#'; return; }
        if (lane == 0) q[i] = (u8)(qq + 33);
        if (q1 > qq) delta += q1 - qq;
        last = qual_ctx(qq, q1, q2, delta, qlevel, i);
        q2 = q1; q1 = qq;
    }
}

// one base (inlined four times in decode_seq @0x4296b0); whole warp, uniform.  Every lane stores
// the updated counters (same value, same address) so that each lane later reads its own store.
// The 256 contexts reachable in four more bases are one contiguous KiB: lanes 0..7 prefetch it.
SQ_D u32 decode_base(u8 *seqtab, RcDec &rc, u32 &last, u32 mask, int lane) {
    if (lane < 8) {
        const u8 *pf = seqtab + ((u64)((last << 8) & mask) << 2) + lane * 128;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
    }
    u32 *S4 = (u32 *)(seqtab + (u64)last * 4);
    u32 x = *S4;
    u32 S0 = x & 0xff, S1 = (x >> 8) & 0xff, S2 = (x >> 16) & 0xff, S3 = x >> 24;
    u32 tot = S0 + S1 + S2 + S3;
    if (tot > 253) { S0 -= S0 >> 1; S1 -= S1 >> 1; S2 -= S2 >> 1; S3 -= S3 >> 1; tot = S0 + S1 + S2 + S3; }
    rc.range /= tot;
    u32 r = rc.range, code = rc.code;
    u32 b, cum, f;
    if (code < S0 * r) { b = 0; cum = 0; f = S0; }
    else if (code < (S0 + S1) * r) { b = 1; cum = S0; f = S1; }
    else if (code < (S0 + S1 + S2) * r) { b = 2; cum = S0 + S1; f = S2; }
    else { b = 3; cum = S0 + S1 + S2; f = S3; }
    rc.decode(cum, f);
    if (b == 0) S0 = (S0 + 1) & 0xff; else if (b == 1) S1 = (S1 + 1) & 0xff; else if (b == 2) S2 = (S2 + 1) & 0xff; else S3 = (S3 + 1) & 0xff;
    *S4 = S0 | (S1 << 8) | (S2 << 16) | (S3 << 24);
    last = (last * 4 + b) & mask;
    return b;
}

// small side streams: "u32le count | range-coder bytes" with a fresh local model (lane 0 only)
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
    NameSmem ns;
    SmModel<2> tip;
    SmModel<11> ch;
    SmModel<95> maxq;
    KModelD km;
    u32 cnts[5];               // n_tip, n_ch, n_mq, n_cnt, n_pos (published by lane 0 of the seq warp)
    int side_err;
    volatile u32 qprog;        // reads whose qualities are complete
};

__global__ void __launch_bounds__(96) k_dec_main(DecBlock *db, int qlevel, int id_rule, u32 seqmask) {
    __shared__ MainSmem sm;
    DecBlock &d = db[blockIdx.x];
    int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { sm.qprog = 0; sm.side_err = 0; }
    __syncthreads();
    if (d.status != 0) return;
    const char tbl[] = "ACGT";
    if (wid == 0) {                                            // names (decompressName @0x428ca0)
        if (id_rule) { if (lane == 0) d.status = SQ_ERR_FORMAT_K; return; }   // rule-based IDs: regenerated on the host ("next" row)
        for (int i = lane; i < 1024; i += 32) sm.ns.last[i] = ' ';
        __syncwarp();
        int lnl = 0, lpl = 0, lsl = 0;
        RcDec rc;
        rc.start(d.rec + d.off[5], d.rec + d.off[5] + d.len[5]);
        u64 o = 0;
        for (u32 r = 0; r < d.nreads; r++) {
            int n = decode_name(d, rc, d.ids + o, d.idcap - o, sm.ns, lnl, lpl, lsl, lane);
            if (n < 0) { if (lane == 0) d.status = SQ_ERR_CAPACITY_K; return; }
            if (lane == 0) d.idlen[r] = (u16)n;
            o += (u32)n;
        }
        if (lane == 0) d.idbytes = o;
    } else if (wid == 1) {                                     // qualities (decompressQual @0x42aa40)
        RcDec rc;
        rc.start(d.rec + d.off[7], d.rec + d.off[7] + d.len[7]);
        for (u32 r = 0; r < d.nreads; r++) {
            decode_qual(d, rc, d.qual + d.seqoff[r], d.seqlen[r], qlevel, lane);
            if ((r & 15) == 15 || r + 1 == d.nreads) {
                __syncwarp();
                __threadfence_block();
                if (lane == 0) sm.qprog = r + 1;
            }
        }
    } else {                                                   // Dege streams, then bases (decompressSeq @0x42e390)
        if (lane == 0) {
            u32 n_tip = dec_small<2>(d, 23, sm.tip, d.d_tip, d.nreads, 0);
            u32 n_ch = dec_small<11>(d, 14, sm.ch, d.d_ch, (u32)d.seqbytes, 1);
            u32 n_mq = dec_small<95>(d, 24, sm.maxq, d.d_maxq, d.nreads, 2);
            if (n_tip == 0xffffffffu || n_ch == 0xffffffffu || n_mq == 0xffffffffu) sm.side_err = 1;
            sm.km.init();                                      // kModelInit @0x42ebd2, shared by tags 25 and 26
            u32 n_cnt = 0, n_pos = 0;
            for (int k = 0; k < 2 && !sm.side_err; k++) {
                int tag = 25 + k;
                if (!d.len[tag]) continue;
                const u8 *pay = d.rec + d.off[tag];
                u32 cnt = get_u32le(pay);
                if (cnt > (k == 0 ? d.nreads : (u32)d.seqbytes)) { sm.side_err = 1; break; }
                RcDec rc;
                rc.start(pay + 4, pay + d.len[tag]);
                u32 *arr = k == 0 ? d.d_cnt : d.d_pos;
                for (u32 i = 0; i < cnt; i++) arr[i] = sm.km.decode(rc);
                if (k == 0) n_cnt = cnt; else n_pos = cnt;
            }
            sm.cnts[0] = n_tip; sm.cnts[1] = n_ch; sm.cnts[2] = n_mq; sm.cnts[3] = n_cnt; sm.cnts[4] = n_pos;
            __threadfence_block();
        }
        __syncwarp();
        if (sm.side_err) { if (lane == 0) d.status = SQ_ERR_FORMAT_K; return; }
        u32 n_tip = sm.cnts[0], n_ch = sm.cnts[1], n_mq = sm.cnts[2], n_cnt = sm.cnts[3], n_pos = sm.cnts[4];
        // the Dege arrays were written by lane 0: the other lanes read them through L2
        RcDec rc;
        rc.start(d.rec + d.off[6], d.rec + d.off[6] + d.len[6]);
        u32 i_mq = 0, i_ch = 0, i_pos = 0;
        for (u32 r = 0; r < d.nreads; r++) {
            u8 *seq = d.seq + d.seqoff[r];
            int len = d.seqlen[r];
            u32 last = 0x7616c7u & seqmask;
            bool flagged = r < n_tip && __ldcg(d.d_tip + r);
            if (!flagged) {
                for (int i = 0; i < len; i++) {
                    u32 bb = decode_base(d.seqtab, rc, last, seqmask, lane);
                    if (lane == 0) seq[i] = (u8)tbl[bb];
                }
                continue;
            }
            while (sm.qprog <= r) __nanosleep(200);            // this read's qualities place its N bases
            __threadfence_block();
            const u8 *qual = d.qual + d.seqoff[r];
            if (i_mq >= n_mq) { if (lane == 0) d.status = SQ_ERR_FORMAT_K; return; }
            int mq = (int)(int8_t)__ldcg(d.d_maxq + i_mq);
            u32 cnt = i_mq < n_cnt ? __ldcg(d.d_cnt + i_mq) : 0;
            i_mq++;
            if (cnt == 0) {                                    // @0x429a20: every candidate position is non-ACGT
                for (int i = 0; i < len; i++) {
                    u32 c;
                    if (mq >= (int)(int8_t)__ldcg(qual + i)) { c = i_ch < n_ch ? __ldcg(d.d_ch + i_ch) : 'N'; i_ch++; }
                    else c = (u8)tbl[decode_base(d.seqtab, rc, last, seqmask, lane)];
                    if (lane == 0) seq[i] = (u8)c;
                }
                continue;
            }
            u32 gap = i_pos < n_pos ? __ldcg(d.d_pos + i_pos) : 0, remaining = cnt - 1, run = 0;   // @0x4298cc
            i_pos++;
            for (int i = 0; i < len; i++) {
                u32 c;
                if (mq >= (int)(int8_t)__ldcg(qual + i)) {
                    if (run == gap) {                          // @0x429b40
                        gap = 0xffffffffu;
                        if (remaining) { remaining--; gap = i_pos < n_pos ? __ldcg(d.d_pos + i_pos) : 0; i_pos++; }
                        c = (u8)tbl[decode_base(d.seqtab, rc, last, seqmask, lane)];
                        run = 0;
                    } else { run++; c = i_ch < n_ch ? __ldcg(d.d_ch + i_ch) : 'N'; i_ch++; }
                } else c = (u8)tbl[decode_base(d.seqtab, rc, last, seqmask, lane)];
                if (lane == 0) seq[i] = (u8)c;
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
    // per-block model reset (or snapshot restore) into the decoder's warp-model tables
    k_init_wm<95, 3><<<dim3(w.nsm * 2, B), 256, 0, st>>>(w.db, 0, w.nqual, 0, w.snap_qual, 293);
    k_init_seqtab<<<dim3(w.nsm, B), 256, 0, st>>>(w.db, (u64)w.nseq * 4, w.snap_seq);
    *launches += 2;
    if (!w.id_rule) {
        // nametab: prefix | suffix | len (256 x SIMPLE_MODEL<256> each) then 8192 x SIMPLE_MODEL<128>
        k_init_wm<256, 8><<<dim3(12, B), 256, 0, st>>>(w.db, 1, 768, 0, w.snap_name, 776);
        k_init_wm<128, 4><<<dim3(32, B), 256, 0, st>>>(w.db, 1, 8192, 768 * W256_STRIDE,
                                                     w.snap_name ? w.snap_name + 3ull * 256 * 776 : nullptr, 392);
        *launches += 2;
    }
    k_dec_len<<<B, 32, 0, st>>>(w.db);
    k_dec_main<<<B, 96, 0, st>>>(w.db, w.qlevel, w.id_rule, w.seqmask);
    *launches += 2;
    if (w.md5) { k_dec_md5<<<cdivd(3ull * B, 32), 32, 0, st>>>(w.db, B); *launches += 1; }
    return cudaGetLastError();
}
