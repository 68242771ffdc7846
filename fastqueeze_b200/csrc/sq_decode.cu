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
#include <cstdio>

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
static_assert(WQ_STRIDE == SQ_WQ_STRIDE, "layout");
static_assert(768ull * W256_STRIDE + 8192ull * W128_STRIDE == SQ_WNAME_WORDS, "layout");

SQ_D u32 warp_incl_scan_d(u32 v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}
// Reciprocal estimate e ~ 2^32 / d (never above it), the seed the compiler's own u32 division uses:
// q' = mulhi(n, e) is within [floor(n / d) - 2, floor(n / d)] for every 32-bit n, so a division by
// a divisor whose estimate is at hand is one multiply and at most two fix-up steps.
SQ_D u32 recip_est(u32 d) {
    float fr;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(fr) : "f"(__uint2float_ru(d)));
    u32 e0 = __float2uint_rz(__uint_as_float(__float_as_uint(fr) + 0x0ffffffeu));
    return e0 + __umulhi(e0, 0u - e0 * d);
}
SQ_D u32 div_est(u32 n, u32 d, u32 e) {
    u32 q = __umulhi(n, e);
    u32 r = n - q * d;
    if (r >= d) { q++; r -= d; }
    if (r >= d) { q++; r -= d; }
    if (r >= d) q += r / d;              // never taken for a valid estimate; keeps the result exact regardless
    return q;
}

// Range decoder with a byte window: the next bytes of the stream sit in a register and are
// refilled four at a time well before they are needed, so no load is on the symbol chain.
struct RcDecW {
    u64 low;
    u32 range, code;
    const u8 *in, *end;
    u64 win;
    int avail;
    SQ_D void refill() {
        u32 v;
        if (in + 4 <= end) v = ((u32)in[0] << 24) | ((u32)in[1] << 16) | ((u32)in[2] << 8) | in[3];
        else {
            v = 0;
            for (int k = 0; k < 4; k++) v = (v << 8) | (in + k < end ? in[k] : 0);
        }
        in += 4;
        win |= (u64)v << (32 - 8 * avail);
        avail += 4;
    }
    SQ_D u32 get() {
        u32 b = (u32)(win >> 56);
        win <<= 8;
        avail--;
        if (avail <= 4) refill();
        return b;
    }
    SQ_D void start(const u8 *i, const u8 *e) {
        low = 0; range = 0xffffffffu; code = 0; in = i; end = e; win = 0; avail = 0;
        refill(); refill();
        for (int k = 0; k < 8; k++) code = (code << 8) | get();
    }
    // range already divided by tot.  A symbol renormalises by at most two bytes (range / tot >= 2^8),
    // so the common path shifts by 0 / 8 / 16 bits without a loop; the carry-less range adjustment
    // (@0x429760-0x429880) needs bits 32..47 of low all ones, which sends us to the exact loop.
    SQ_D void decode(u32 cum, u32 freq) {
        u32 t = cum * range;
        low += t;
        code -= t;
        range *= freq;
        if (range >= (1u << 24)) return;
        if ((~(u32)(low >> 32) & 0xffffu) == 0) {
            while (range < (1u << 24)) {
                if ((low ^ (low + range)) >> 56) range = (((u32)low) | 0xffffffu) - (u32)low;
                code = (code << 8) | get();
                range <<= 8;
                low <<= 8;
            }
            return;
        }
        u32 sh = range < (1u << 16) ? 16 : 8;
        code = __funnelshift_l((u32)(win >> 32), code, sh);
        win <<= sh;
        avail -= (int)(sh >> 3);
        range <<= sh;
        low <<= sh;
        if (avail <= 4) refill();
    }
};

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
// decodeSymbol (inlined e.g. @0x42a7d0-0x42a960) on a warp model, in three phases so that the
// caller can issue the NEXT model's loads between the search and the write-back of this one.
template <int EPL> struct WReg { u32 w[EPL]; u32 bt; };

// L2 residency hints.  The warp-model tables (hot set: a few hundred KB per block) must survive in L2
// beside the base-context tables, whose accesses are uniformly random over 4 MB per block and would
// otherwise evict them: models are accessed evict_last, base-context sectors evict_first.
SQ_D u64 l2_policy_keep() { u64 p; asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
SQ_D u64 l2_policy_stream() { u64 p; asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
#ifndef SQ_L2_HINTS
#define SQ_L2_HINTS 2          // 0: none, 1: everywhere, 2: base-context accesses only (measured best)
#endif
template <bool HINT = (SQ_L2_HINTS == 1)>
SQ_D u32 ld_hint(const u32 *p, u64 pol) {
    if (!HINT) return *p;
    u32 v;
    asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
template <bool HINT = (SQ_L2_HINTS == 1)>
SQ_D void st_hint(u32 *p, u32 v, u64 pol) {
    if (!HINT) { *p = v; return; }
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

template <int EPL>
SQ_D void wm_load(WReg<EPL> &R, const u32 *mp, int lane, u64 pol) {
#pragma unroll
    for (int i = 0; i < EPL; i++) R.w[i] = ld_hint(mp + i * 32 + lane, pol);
    R.bt = ld_hint(mp + EPL * 32 + lane, pol);
}

// search + coder step; rc is warp-uniform.  Returns the symbol on every lane; kl / li locate it.
// wm_find leaves the coder step (rc.decode(cex, freq)) to the caller, which can then request the
// next model -- it only depends on the symbol -- before renormalising.
template <int N, int EPL>
SQ_D u32 wm_find(const WReg<EPL> &R, RcDecW &rc, int lane, int &kl, int &li, u32 &cex_out, u32 &freq_out) {
    u32 base = R.bt & 0xffffu, tot = R.bt >> 16;
    u32 q = div_est(rc.range, tot, recip_est(tot));
    rc.range = q;
    u32 code = rc.code;
    li = -1;
    u32 c = base, cex = 0, wsel = 0;
#pragma unroll
    for (int i = 0; i < EPL; i++) {
        u32 cn = c + (R.w[i] >> 16);
        bool h = (li < 0) && (code < cn * q);
        if (h) { li = i; cex = c; wsel = R.w[i]; }
        c = cn;
    }
    u32 bal = __ballot_sync(FULL, li >= 0);
    // code beyond tot * r only happens on a corrupt stream: take the last symbol then
    bool lastl = (bal == 0) && (lane == (N - 1) / EPL);
    li = lastl ? (N - 1) % EPL : li;
    wsel = lastl ? R.w[(N - 1) % EPL] : wsel;
    cex = lastl ? c - (R.w[(N - 1) % EPL] >> 16) : cex;
    kl = bal ? (__ffs(bal) - 1) : (N - 1) / EPL;
    wsel = __shfl_sync(FULL, wsel, kl);
    cex = __shfl_sync(FULL, cex, kl);
    cex_out = cex;
    freq_out = wsel >> 16;
    return wsel & 0xff;
}
template <int N, int EPL>
SQ_D u32 wm_search(const WReg<EPL> &R, RcDecW &rc, int lane, int &kl, int &li) {
    u32 cex, freq;
    u32 sym = wm_find<N, EPL>(R, rc, lane, kl, li, cex, freq);
    rc.decode(cex, freq);
    return sym;
}

// model update (Freq += 8, TotFreq += 8, halve above 0xffe0, bubble swap every 16th symbol) and
// write-back; R is left holding the updated model.
template <int N, int EPL>
SQ_D void wm_update_store(u32 *mp, WReg<EPL> &R, int kl, int li, int lane, u64 pol) {
    u32 base = R.bt & 0xffffu, tot = R.bt >> 16;
    u32 bub = (R.w[0] >> 8) & 0xff;
    R.w[0] &= 0xffff00ffu;
    bool mine = lane == kl;
#pragma unroll
    for (int i = 0; i < EPL; i++) R.w[i] += (mine && i == li) ? (8u << 16) : 0u;
    base += (lane > kl) ? 8u : 0u;
    tot += 8;
    bool all = false;
    if (tot > 0xffe0u) {
        u32 T = 0;
#pragma unroll
        for (int i = 0; i < EPL; i++) { u32 f = R.w[i] >> 16; f -= f >> 1; R.w[i] = (R.w[i] & 0xffffu) | (f << 16); T += f; }
        u32 inc = warp_incl_scan_d(T, lane);
        base = inc - T;
        tot = __shfl_sync(FULL, inc, 31);
        all = true;
    }
    bub = (bub + 1) & 0xff;
    bool prevlane = false;
    if ((bub & 15) == 0) {
        u32 up = __shfl_up_sync(FULL, R.w[EPL - 1], 1);
        u32 cur = 0, prv = up;
#pragma unroll
        for (int i = 0; i < EPL; i++) {
            if (i == li) cur = R.w[i];
            if (i + 1 == li) prv = R.w[i];
        }
        bool dosw = mine && (kl * EPL + li > 0) && (cur >> 16) > (prv >> 16);
        u32 swb = __ballot_sync(FULL, dosw);
        if (swb) {
            int lik = __shfl_sync(FULL, li, kl);
            u32 cw = __shfl_sync(FULL, cur, kl), pw = __shfl_sync(FULL, prv, kl);
            if (lik > 0) {
#pragma unroll
                for (int i = 1; i < EPL; i++) if (mine && i == li) { R.w[i] = pw; R.w[i - 1] = cw; }
            } else {
                if (mine) { R.w[0] = pw; base += (cw >> 16) - (pw >> 16); }
                if (lane == kl - 1) { R.w[EPL - 1] = cw; prevlane = true; }
            }
        }
    }
    R.w[0] |= bub << 8;
    R.bt = base | (tot << 16);
    // row 0 carries bub and the base row tot: both are rewritten by every lane
    st_hint(mp + lane, R.w[0], pol);
    st_hint(mp + EPL * 32 + lane, R.bt, pol);
    if (mine || prevlane || all) {
#pragma unroll
        for (int i = 1; i < EPL; i++) st_hint(mp + i * 32 + lane, R.w[i], pol);
    }
}

template <int N, int EPL>
SQ_D u32 wm_decode(u32 *mp, RcDecW &rc, int lane, u64 pol) {
    WReg<EPL> R;
    wm_load(R, mp, lane, pol);
    int kl, li;
    u32 sym = wm_search<N, EPL>(R, rc, lane, kl, li);
    wm_update_store<N, EPL>(mp, R, kl, li, lane, pol);
    return sym;
}

// ---------------------------------------------------------------------------------------------
// Quality models of the decoder: "Q4" layout.  A SIMPLE_MODEL<95> is 32 lanes x 16 bytes (512 B, one
// 128-bit load / store per lane); lane l owns entries 3l, 3l+1, 3l+2 (slot 95 is a dummy of frequency 0):
//   x = c1 | c2 << 16      inclusive cumulative frequencies through the lane's 1st / 2nd entry
//   y = c3 | tot << 16     ... through its 3rd entry; TotFreq
//   z = floor((2^32 - 1) / tot)   the reciprocal the division range / tot needs -- kept in the model so that
//                                 no int->float->rcp->int conversion sits on the symbol chain
//   w = s0 | s1 << 8 | s2 << 16 | (BubCnt & 15) << 24
// Cumulative sums are absolute (they include the lower lanes), so the search needs no prefix: lane l "hits"
// when code < c3_l * r, hits are monotone in l, and the symbol's lane is the number of lanes that miss.
// The update adds 8 to every cumulative sum at or above the symbol -- one packed add per word.
// In memory the words from SQ_QF on (the TAIL) hold their sums relative to the total of the words before them: see the
// front / tail split below.
struct RcWin;
SQ_D u32 recip32(u32 d) { return 0xffffffffu / (d ? d : 1u); }     // in [2^32/d - 1, 2^32/d]: mulhi(n, .) is floor(n/d) or one less

template <int N>
__global__ void __launch_bounds__(256) k_init_q4(DecBlock *db, u32 nmodels, const u8 *snap, u32 snapstride, u32 nsnap) {
    static_assert(N <= 96, "three entries per lane");
    int b = blockIdx.y;
    if (db[b].status != 0) return;
    uint4 *tab = (uint4 *)db[b].qualtab;
    int lane = threadIdx.x & 31;
    u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 m = warp; m < nmodels; m += nwarps) {
        const u8 *s = (snap && m < nsnap) ? snap + (u64)m * snapstride : nullptr;
        u32 tot = N, bub = 0;
        if (s) { tot = get_u32le(s) & 0xffffu; bub = get_u32le(s + 4) & 15u; }
        u32 f[3], sy[3], T = 0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            int slot = lane * 3 + i;
            sy[i] = (u32)slot & 0xff; f[i] = slot < N ? 1 : 0;
            if (s && slot < N) { const u8 *q = s + 8 + 3 * slot; sy[i] = q[0]; f[i] = (u32)(q[1] | (q[2] << 8)); }
            T += f[i];
        }
        u32 base = warp_incl_scan_d(T, lane) - T;
        u32 c1 = (base + f[0]) & 0xffffu, c2 = (base + f[0] + f[1]) & 0xffffu, c3 = (base + T) & 0xffffu;
        if (SQ_QF < 32) {                                  // tail lanes hold their sums relative to the front's total
            const u32 F = __shfl_sync(FULL, c3, SQ_QF - 1);
            if (lane >= SQ_QF) { c1 = (c1 - F) & 0xffffu; c2 = (c2 - F) & 0xffffu; c3 = (c3 - F) & 0xffffu; }
        }
        uint4 v;
        v.x = c1 | (c2 << 16);
        v.y = c3 | (tot << 16);
        v.z = recip32(tot);
        v.w = sy[0] | (sy[1] << 8) | (sy[2] << 16) | (bub << 24);
        tab[(u64)m * 32 + lane] = v;
    }
}

// Range decoder with a byte window fed by aligned 32-bit loads issued one refill ahead: no load and no
// byte assembly sits on the symbol chain (RangeCoder decode @0x429760-0x429880).
struct RcWin {
    u64 low;
    u32 range, code;
    u64 win;                 // the next bytes of the stream, first byte on top
    int avail;               // valid bytes in win
    const u32 *wp, *wend;    // next aligned word to fetch / first word past the stream
    u32 nxt;                 // the word fetched ahead
    SQ_D void fetch() { nxt = wp < wend ? __ldg(wp) : 0u; wp++; }
    // the same under a predicate, without a branch (keeps the pipelined loop body one basic block)
    SQ_D void fetch_if(bool c) {
        u32 ok = (c && wp < wend) ? 1u : 0u, v = nxt;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}" : "+r"(v) : "l"(wp), "r"(ok));
        nxt = c ? (ok ? v : 0u) : nxt;
        wp += c ? 1 : 0;
    }
    SQ_D void refill() {     // needs avail <= 4
        win |= (u64)__byte_perm(nxt, 0, 0x0123) << (32 - 8 * avail);
        avail += 4;
        fetch();
    }
    SQ_D u32 get() {
        u32 b = (u32)(win >> 56);
        win <<= 8;
        avail--;
        if (avail <= 4) refill();
        return b;
    }
    SQ_D void start(const u8 *i, const u8 *e) {
        low = 0; range = 0xffffffffu; code = 0;
        uintptr_t a = (uintptr_t)i;
        int skip = (int)(a & 3);
        wp = (const u32 *)(a - skip);
        wend = (const u32 *)(((uintptr_t)e + 3) & ~(uintptr_t)3);
        fetch();
        win = (u64)__byte_perm(nxt, 0, 0x0123) << (32 + 8 * skip);
        avail = 4 - skip;
        fetch();
        refill();
        for (int k = 0; k < 8; k++) code = (code << 8) | get();
    }
    // `range` already holds range / tot.  A symbol renormalises by at most two bytes in the common case
    // (range / tot >= 2^8); the carry-less range adjustment needs bits 32..47 of low all ones, which sends
    // us to the exact loop.
    SQ_D void decode(u32 cum, u32 freq) {
        u32 t = cum * range;
        low += t;
        code -= t;
        range *= freq;
        if (range >= (1u << 24)) return;
        if ((~(u32)(low >> 32) & 0xffffu) == 0 || range < (1u << 8)) {
            for (int guard = 0; guard < 16 && range < (1u << 24); guard++) {
                if ((low ^ (low + range)) >> 56) range = (((u32)low) | 0xffffffu) - (u32)low;
                code = (code << 8) | get();
                range <<= 8;
                low <<= 8;
            }
            return;
        }
        u32 sh = range < (1u << 16) ? 16 : 8;
        code = __funnelshift_l((u32)(win >> 32), code, sh);
        win <<= sh;
        avail -= (int)(sh >> 3);
        range <<= sh;
        low <<= sh;
        if (avail <= 4) refill();
    }
    // The same step without a branch, for the software-pipelined quality loop: returns true when the exact
    // loop has to run instead (carry-less adjustment possible, or more than two bytes) -- the state is then
    // left un-renormalised for renorm_exact().
    SQ_D bool decode_fast(u32 cum, u32 freq) {
        u32 t = cum * range;
        low += t;
        code -= t;
        range *= freq;
        const bool slow = range < (1u << 24) && ((~(u32)(low >> 32) & 0xffffu) == 0 || range < (1u << 8));
        u32 sh = (range < (1u << 24) ? 8u : 0u) + (range < (1u << 16) ? 8u : 0u);
        sh = slow ? 0u : sh;
        code = __funnelshift_l((u32)(win >> 32), code, sh);
        win <<= sh;
        range <<= sh;
        low <<= sh;
        avail -= (int)(sh >> 3);
        const bool rf = avail <= 4;
        const u64 add = (u64)__byte_perm(nxt, 0, 0x0123) << ((32 - 8 * avail) & 63);
        win |= rf ? add : 0ull;
        avail += rf ? 4 : 0;
        fetch_if(rf);
        return slow;
    }
    SQ_D void renorm_exact() {
        for (int guard = 0; guard < 16 && range < (1u << 24); guard++) {
            if ((low ^ (low + range)) >> 56) range = (((u32)low) | 0xffffffu) - (u32)low;
            code = (code << 8) | get();
            range <<= 8;
            low <<= 8;
        }
    }
};

// L2 hints in the two long chains (measured on the full job, 274 blocks resident): the base-context ring copies and
// stores marked evict_first: decode 7.37 -> 7.04 s (they stream 1.5 TB through the L2 the quality models live in);
// the quality models marked evict_last on top of that: no change.  The hint rides in the instruction's descriptor:
// no instruction is added.
#ifndef SQ_POL_QUAL
#define SQ_POL_QUAL 0
#endif
#ifndef SQ_POL_BASE
#define SQ_POL_BASE 1
#endif
#if SQ_POL_QUAL
#define SQ_QPOL_ASM ".L2::cache_hint"
#define SQ_QPOL_ARG(n) ", %" #n
#define SQ_QPOL_IN , "l"(l2_policy_keep())
#else
#define SQ_QPOL_ASM ""
#define SQ_QPOL_ARG(n) ""
#define SQ_QPOL_IN
#endif
SQ_D uint4 q4_ld(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global" SQ_QPOL_ASM ".v4.u32 {%0,%1,%2,%3}, [%4]" SQ_QPOL_ARG(5) ";" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) SQ_QPOL_IN : "memory");
    return v;
}
SQ_D void q4_st(uint4 *p, const uint4 &v) {
    asm volatile("st.global" SQ_QPOL_ASM ".v4.u32 [%0], {%1,%2,%3,%4}" SQ_QPOL_ARG(5) ";" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) SQ_QPOL_IN : "memory");
}
SQ_D u32 shl_clamp(u32 v, u32 n) { u32 r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n)); return r; }

// Front / tail split of a model.  The bubble swaps keep a context's frequent symbols at the head of its table, and the
// symbols a data set never uses (55 of 95 for Illumina qualities) stay behind them with frequency 1.  Per symbol only the
// FRONT is touched: the first SQ_QF 16-byte words = 3 SQ_QF entries (2 cache lines instead of 4 for SQ_QF = 16).  Lanes
// >= SQ_QF load the SAME word as lane - SQ_QF (no extra lines): every lane then has tot / the reciprocal / the bubble
// count and all conditions stay warp-uniform; their hit predicate is forced off and they never store.  The remaining
// words (the TAIL) hold their cumulative sums RELATIVE to the front's total F: a front update adds 8 to F and to every
// absolute sum behind it, so the relative form does not change.  A symbol in the tail (all front lanes miss) and the
// halving of a model take the whole model, one word per lane, straight from memory (q4_full_*): registers and memory
// agree at those points.  The tail test is NOT a branch on the symbol chain: the step runs on speculatively (a branch
// after the ballot cost 40 of 360 cycles) with a coder update that is the identity, and the step's closing branch redoes
// it on the whole model.  Measured on the full job: the quality chain 12.5 -> 11.4 G cycles (fewer L1 / L2 misses with
// two blocks per SM); a layout with every front word replicated on two neighbouring lanes (no predicates at all) was
// 30 % slower.
static_assert(SQ_QF == 32 || SQ_QF == 16 || SQ_QF == 8, "front words per model");
SQ_D int q4_flane(int lane) { return lane & (SQ_QF - 1); }
// store of the front words: only the lanes below SQ_QF own one
SQ_D void q4_stf(uint4 *p, const uint4 &v, int lane) {
    if (SQ_QF >= 32) { q4_st(p, v); return; }
    const u32 front = lane < SQ_QF ? 1u : 0u;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.global" SQ_QPOL_ASM ".v4.u32 [%0], {%1,%2,%3,%4}" SQ_QPOL_ARG(6) ";\n\t}"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(front) SQ_QPOL_IN : "memory");
}

// rare: TotFreq passed 0xffe0 (@0x4222b7): halve every frequency, rebuild the cumulative sums (natural layout)
SQ_D void q4_halve(uint4 &M, int lane) {
    u32 c1 = M.x & 0xffffu, c2 = M.x >> 16, c3 = M.y & 0xffffu;
    u32 c0 = __shfl_up_sync(FULL, c3, 1);
    c0 = lane ? c0 : 0u;
    u32 f0 = (c1 - c0) & 0xffffu, f1 = (c2 - c1) & 0xffffu, f2 = (c3 - c2) & 0xffffu;
    f0 -= f0 >> 1; f1 -= f1 >> 1; f2 -= f2 >> 1;
    u32 T = f0 + f1 + f2;
    u32 inc = warp_incl_scan_d(T, lane);
    u32 base = inc - T;
    u32 tot = __shfl_sync(FULL, inc, 31);
    M.x = ((base + f0) & 0xffffu) | ((base + f0 + f1) << 16);
    M.y = (inc & 0xffffu) | (tot << 16);
}
// every 16th symbol of a model (@0x4222d3-0x4222f1): swap the coded entry with its predecessor when its frequency is
// strictly greater.  kl / li locate the coded entry (li is only meaningful on lane kl).
SQ_D void q4_bubble(uint4 &M, int kl, int li, int lane) {
    u32 c1 = M.x & 0xffffu, c2 = M.x >> 16, c3 = M.y & 0xffffu;
    u32 c0 = __shfl_up_sync(FULL, c3, 1);
    c0 = lane ? c0 : 0u;
    u32 up_f2 = __shfl_up_sync(FULL, c3 - c2, 1);          // frequency of the previous lane's last entry
    u32 up_s2 = __shfl_up_sync(FULL, (M.w >> 16) & 0xffu, 1);
    u32 fk = li == 0 ? c1 - c0 : (li == 1 ? c2 - c1 : c3 - c2);
    u32 fp = li == 0 ? up_f2 : (li == 1 ? c1 - c0 : c2 - c1);
    bool mine = lane == kl;
    bool dosw = mine && (kl * 3 + li > 0) && fk > fp;
    u32 swb = __ballot_sync(FULL, dosw);
    if (!swb) return;
    int lik = __shfl_sync(FULL, li, kl);
    u32 fkb = __shfl_sync(FULL, fk, kl);
    u32 s0b = __shfl_sync(FULL, M.w & 0xffu, kl);
    if (lik == 1) {
        if (mine) { c1 = c0 + fk; M.x = c1 | (c2 << 16); M.w = __byte_perm(M.w, 0, 0x3201); }
    } else if (lik == 2) {
        if (mine) { c2 = c1 + fk; M.x = c1 | (c2 << 16); M.w = __byte_perm(M.w, 0, 0x3120); }
    } else {
        // the predecessor is the last entry of lane kl - 1: its inclusive sum becomes c2 + f_k there
        if (mine) M.w = (M.w & 0xffffff00u) | up_s2;
        if (lane == kl - 1) { M.y = (M.y & 0xffff0000u) | ((c2 + fkb) & 0xffffu); M.w = (M.w & 0xff00ffffu) | (s0b << 16); }
    }
}

// the whole model `ctx` in its natural layout (lane l = word l), tail sums made absolute; tabf = table + q4_flane(lane)
SQ_D uint4 q4_full_in(const uint4 *tabf, u32 ctx, int lane) {
    __syncwarp();
    uint4 N = q4_ld(tabf + (u64)ctx * 32 + (lane - q4_flane(lane)));
    if (SQ_QF < 32) {
        const u32 F = __shfl_sync(FULL, N.y & 0xffffu, SQ_QF - 1);
        const u32 y0 = __shfl_sync(FULL, N.y, 0), z0 = __shfl_sync(FULL, N.z, 0), w0 = __shfl_sync(FULL, N.w, 0);
        if (lane >= SQ_QF) {
            N.x += F * 0x10001u;
            N.y = (y0 & 0xffff0000u) | ((N.y + F) & 0xffffu);
            N.z = z0;
            N.w = (w0 & 0xff000000u) | (N.w & 0x00ffffffu);
        }
    }
    return N;
}
SQ_D void q4_full_out(uint4 *tabf, u32 ctx, const uint4 &N, int lane) {
    uint4 V = N;
    if (SQ_QF < 32) {
        const u32 F = __shfl_sync(FULL, N.y & 0xffffu, SQ_QF - 1);
        if (lane >= SQ_QF) { V.x = N.x - F * 0x10001u; V.y = (N.y & 0xffff0000u) | ((N.y - F) & 0xffffu); }
    }
    q4_st(tabf + (u64)ctx * 32 + (lane - q4_flane(lane)), V);
    __syncwarp();
}

// context of the next symbol (encode_qual @0x422474-0x422505); 0 opens the next read.
// SAT: delta has reached 56 in this read (it only grows), so its context bits are constant.
template <int QLEVEL, bool SAT>
SQ_D u32 q4_next_ctx(u32 qq, u32 q1, u32 q2, u32 &delta, bool more, int i) {
    const u32 amax = q1 > q2 ? q1 : q2;
    u32 K = 0;
    if (QLEVEL > 1) K |= (u32)(q1 == q2) << 12;
    if (QLEVEL > 2) K |= (i <= 111) ? (u32)((i + 15) & 0x78) << 13 : 0xf0000u;
    if (SAT || QLEVEL == 1) {
        if (QLEVEL > 1) K |= (56u & 0xf8u) << 10;
        const u32 mk = more ? 0xfffu : 0u;
        K = more ? K : 0u;
        return ((amax * 64 + qq) & mk) | K;
    }
    int dd = (int)q1 - (int)qq;
    u32 d2 = delta + (u32)(dd > 0 ? dd : 0);
    delta = d2;
    return (((amax * 64 + qq) & 0xfffu) | K | (((d2 < 56 ? d2 : 56u) & 0xf8u) << 10)) & (more ? 0xffffffffu : 0u);
}

// Software-pipelined quality step (decode_qual @0x42a750; decodeSymbol inlined @0x42a7d0-0x42a960).  A warp issues in
// order and at best one instruction every other cycle, so what a symbol costs is (a) its instruction COUNT and (b) how
// well independent chains interleave.  The next model's load is issued as soon as the symbol is known.  The common-case
// update of model i (none of it on the symbol chain) is DEFERRED: it is carried as a pending record into the next step
// and sits in the same basic block as that step's search.  Cases the deferred form cannot take are completed on the spot
// (ONE uniform branch at the end of the step): the next symbol uses the same context, the update halves or runs the
// 16-symbol bubble check, the total is too small for the one-step Newton refresh of the reciprocal, or the symbol was
// not among the front entries (then the step so far was a no-op on the coder and is redone on the whole model).
//   * the +8 masks come straight from the search predicates: lanes above the symbol's compare "code < c1 r" (add to all
//     three sums), lanes below compare "code >= c3 r" (add nothing), the symbol's own pick by p1 / p2;
//   * reciprocal of tot + 8 from the reciprocal of tot: e1 = e - mulhi(e, e (tot + 8) - 2^32) is ceil(E (1 - d^2)),
//     d <= 8 / tot, i.e. floor(E) or ceil(E) once 2^38 / tot^3 < 1: one downward fix-up (SQ_Q_NEWTON_MIN).
struct QPend { uint4 M; u32 ctx, addx, addy; };
#define SQ_Q_NEWTON_MIN 8192u

SQ_D void q4_pend_none(QPend &P, u32 dummy_ctx) { P.M = make_uint4(0, 0x20000000u, 0x20000u, 0); P.ctx = dummy_ctx; P.addx = 0; P.addy = 0; }

SQ_D void q4_apply(uint4 *tab, const QPend &P, int lane) {
    uint4 Q = P.M;
    Q.x += P.addx;
    Q.y += P.addy;
    Q.w += 1u << 24;                                  // the bubble count does not wrap here (checked when deferring)
    const u32 d = Q.y >> 16;
    u32 e = Q.z;
    e -= __umulhi(e, e * d);                          // e * d - 2^32 fits 32 bits: the old reciprocal overshoots the new total
    if ((i32)(e * d) >= 0) e--;                       // e * d wrapped past 2^32 - 1
    Q.z = e;
    q4_stf(tab + (u64)P.ctx * 32, Q, lane);
}

// One symbol on the WHOLE model of context `last` (memory holds its current state): the symbol lies behind the front
// entries.  Search over the natural layout, update, halving (@0x4222b7), bubble check (@0x4222d3), reciprocal by
// division, front and tail written back; returns with the next model's front loaded.  r = range / tot.
template <int QLEVEL, bool SAT>
SQ_D u32 q4_full_step(uint4 &M, u32 &last, QPend &P, RcWin &rc, uint4 *tab, u8 *qp, bool more, int i, u32 &q1, u32 &q2, u32 &delta,
                      int lane, u32 dummy_ctx, u32 r) {
    uint4 N = q4_full_in(tab, last, lane);
    const u32 tot = N.y >> 16;
    const u32 c1 = N.x & 0xffffu, c2 = N.x >> 16, c3 = N.y & 0xffffu;
    u32 c0 = __shfl_up_sync(FULL, c3, 1);
    c0 = lane ? c0 : 0u;
    const u32 code = rc.code;
    const bool p1 = code >= c1 * r, p2 = code >= c2 * r, p3 = code >= c3 * r;
    const int kl = __popc(__ballot_sync(FULL, p3));
    const int li = (p1 ? 1 : 0) + (p2 ? 1 : 0);
    const u32 cex = p2 ? c2 : (p1 ? c1 : c0);
    const u32 cnx = p2 ? c3 : (p1 ? c2 : c1);
    u32 A = ((cnx - cex) << 16) + cex;
    const u32 sy = __byte_perm(N.w, 0, 0x4440u + (u32)li);
    const u32 qq = __shfl_sync(FULL, sy, kl);
    A = __shfl_sync(FULL, A, kl);
    const u32 nl = q4_next_ctx<QLEVEL, SAT>(qq, q1, q2, delta, more, i);
    rc.range = r;
    rc.decode(A & 0xffffu, A >> 16);
    if (lane == 0) *qp = (u8)(qq + 33);
    N.x += p2 ? 0u : (p1 ? 0x00080000u : 0x00080008u);
    N.y += p3 ? 0x00080000u : 0x00080008u;
    const bool bubdue = ((N.w + (1u << 24)) & (15u << 24)) == 0;
    N.w = (N.w & 0xf0ffffffu) | ((N.w + (1u << 24)) & (15u << 24));
    if (tot + 8 > 0xffe0u) q4_halve(N, lane);
    if (bubdue) q4_bubble(N, kl, li, lane);
    N.z = recip32(N.y >> 16);
    q4_full_out(tab, last, N, lane);
    M = q4_ld(tab + (u64)nl * 32);
    q4_pend_none(P, dummy_ctx);
    last = nl;
    q2 = q1; q1 = qq;
    return qq;
}

// the model of `ctx` (memory holds its state after this symbol's update) passed 0xffe0: halve all of it, then the
// bubble check of the coded entry (kn = its word, li = its index in the word), new reciprocal, written back
SQ_D void q4_full_halve(uint4 *tab, u32 ctx, bool bubdue, int kn, int li, int lane) {
    uint4 N = q4_full_in(tab, ctx, lane);
    q4_halve(N, lane);
    if (bubdue) q4_bubble(N, kn, li, lane);
    N.z = recip32(N.y >> 16);
    q4_full_out(tab, ctx, N, lane);
}

// tab = table + q4_flane(lane)
template <int QLEVEL, bool SAT>
SQ_D u32 q4_pstep(uint4 &M, u32 &last, QPend &P, RcWin &rc, uint4 *tab, u8 *qp, bool more, int i, u32 &q1, u32 &q2, u32 &delta,
                  int lane, u32 dummy_ctx) {
    const u32 tot = M.y >> 16;
    const u32 c1 = M.x & 0xffffu, c2 = M.x >> 16, c3 = M.y & 0xffffu;
    u32 c0 = __shfl_up_sync(FULL, c3, 1);
    c0 = lane ? c0 : 0u;
    q4_apply(tab, P, lane);                           // previous symbol's model (or the dummy slot): interleaves with the search below
    u32 r = __umulhi(rc.range, M.z);
    r += (rc.range - r * tot >= tot) ? 1u : 0u;
    const u32 code = rc.code;
    const bool p1 = code >= c1 * r, p2 = code >= c2 * r;                        // monotone
    const bool p3 = SQ_QF >= 32 ? code >= c3 * r : (lane < SQ_QF && code >= c3 * r);
    const u32 miss = __ballot_sync(FULL, p3);
    const int kl = __popc(miss);                      // hits are monotone in the lane index
    const bool tail = SQ_QF < 32 && kl >= SQ_QF;      // not among the front entries
    const u32 cex = p2 ? c2 : (p1 ? c1 : c0);
    const u32 cnx = p2 ? c3 : (p1 ? c2 : c1);
    u32 A = ((cnx - cex) << 16) + cex;
    const u32 sy = __byte_perm(M.w, 0, 0x4440u + (p1 ? 1u : 0u) + (p2 ? 1u : 0u));
    const u32 qq = __shfl_sync(FULL, sy, kl);
    A = __shfl_sync(FULL, A, kl);
    const u32 addx = p2 ? 0u : (p1 ? 0x00080000u : 0x00080008u);
    const u32 addy = p3 ? 0x00080000u : 0x00080008u;
    const u32 delta0 = delta;
    const u32 nl = q4_next_ctx<QLEVEL, SAT>(qq, q1, q2, delta, more, i);
    const uint4 Mn = q4_ld(tab + (u64)nl * 32);
    // tail: (cum 0, freq 1) on the unchanged range leaves the coder as it is; the step is redone below
    const u32 range0 = rc.range;
    rc.range = tail ? range0 : r;
    rc.decode(tail ? 0u : (A & 0xffffu), tail ? 1u : (A >> 16));
    if (lane == 0) *qp = (u8)(qq + 33);
    // deferred unless: same context next, halving (tot + 8 > 0xffe0), total below the Newton range, bubble check due
    const bool special = (M.y - (SQ_Q_NEWTON_MIN << 16)) > (((0xffe0u - 8u - SQ_Q_NEWTON_MIN) << 16) | 0xffffu);
    const bool bubdue = ((M.w + (1u << 24)) & (15u << 24)) == 0;
    const bool odd = special || bubdue || tail;
    if (nl == last || odd) {
        if (tail) {
            delta = delta0;
            return q4_full_step<QLEVEL, SAT>(M, last, P, rc, tab, qp, more, i, q1, q2, delta, lane, dummy_ctx, r);
        }
        M.x += addx;
        M.y += addy;
        M.w = (M.w & 0xf0ffffffu) | ((M.w + (1u << 24)) & (15u << 24));
        if (!odd) {                                   // the same context again (runs of one quality value): Newton refresh as in q4_apply
            const u32 d = M.y >> 16;
            u32 e = M.z;
            e -= __umulhi(e, e * d);
            if ((i32)(e * d) >= 0) e--;
            M.z = e;
            q4_stf(tab + (u64)last * 32, M, lane);
        } else if (tot + 8 > 0xffe0u) {               // the whole model halves
            q4_stf(tab + (u64)last * 32, M, lane);
            q4_full_halve(tab, last, bubdue, kl, (p1 ? 1 : 0) + (p2 ? 1 : 0), lane);
            M = q4_ld(tab + (u64)last * 32);
        } else {
            if (bubdue) q4_bubble(M, kl, (p1 ? 1 : 0) + (p2 ? 1 : 0), lane);
            M.z = recip32(M.y >> 16);
            q4_stf(tab + (u64)last * 32, M, lane);
        }
        if (nl != last) M = Mn;
        q4_pend_none(P, dummy_ctx);
    } else {
        P.M = M; P.ctx = last; P.addx = addx; P.addy = addy;
        M = Mn;
    }
    last = nl;
    q2 = q1; q1 = qq;
    return qq;
}

// ---------------------------------------------------------------------------------------------
// Len stream (decompressLen_short @0x428170 / decode_len_short @0x427ca0) + read offsets.
// The reference keeps last_len across blocks per decoder instance; a zero-length
// read therefore decodes to "whatever came before".  Here last_len starts at 0 for every block.
struct LenSmem { SmModel<256> len1, len2, len3, len4; SmModel<2> same; };
__global__ void __launch_bounds__(32) k_dec_len(DecBlock *db) {
    __shared__ LenSmem sm;
    DecBlock &d = db[blockIdx.x];
    if (threadIdx.x != 0 || d.status != 0) return;
    if (d.nreads > d.readcap) { d.status = SQ_ERR_CAPACITY_K; return; }
    if (d.len[4] < 8 && d.nreads) { d.status = SQ_ERR_FORMAT_K; return; }
    RcDec rc;
    rc.start(d.rec + d.off[4], d.rec + d.off[4] + d.len[4]);
    sm.len1.init(); sm.len2.init(); sm.same.init();
    const bool lng = d.longflag != 0;                 // decode_len_long @0x427190: four length bytes
    if (lng) { sm.len3.init(); sm.len4.init(); }
    i32 last = 0;
    u64 tot = 0;
    for (u32 r = 0; r < d.nreads; r++) {
        if (sm.same.decode(rc) == 0) {
            u32 l1 = sm.len1.decode(rc), l2 = sm.len2.decode(rc);
            u32 l3 = lng ? sm.len3.decode(rc) : 0, l4 = lng ? sm.len4.decode(rc) : 0;
            last = (i32)(l1 | (l2 << 8) | (l3 << 16) | (l4 << 24));
            if (last < 0) { d.status = SQ_ERR_FORMAT_K; return; }
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

SQ_D int decode_name(u32 *nametab, RcDecW &rc, u8 *name_out, u64 room, NameSmem &ns, int &lnl, int &lpl, int &lsl, int lane, u64 pol) {
    u32 *prefix = nametab, *suffix = prefix + 256 * W256_STRIDE, *lenm = suffix + 256 * W256_STRIDE;
    u32 *middle = lenm + 256 * W256_STRIDE;
    // the three header models of a read are known up front: fetch them together
    WReg<8> Rp, Rs, Rl;
    u32 *mpp = prefix + (u64)lpl * W256_STRIDE, *mps = suffix + (u64)lsl * W256_STRIDE, *mpl = lenm + (u64)lnl * W256_STRIDE;
    wm_load(Rp, mpp, lane, pol); wm_load(Rs, mps, lane, pol); wm_load(Rl, mpl, lane, pol);
    int kl, li;
    int p = (int)wm_search<256, 8>(Rp, rc, lane, kl, li); wm_update_store<256, 8>(mpp, Rp, kl, li, lane, pol);
    int s = (int)wm_search<256, 8>(Rs, rc, lane, kl, li); wm_update_store<256, 8>(mps, Rs, kl, li, lane, pol);
    int n = (int)wm_search<256, 8>(Rl, rc, lane, kl, li); wm_update_store<256, 8>(mpl, Rl, kl, li, lane, pol);
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
        u32 ch = wm_decode<128, 4>(middle + (u64)ctx * W128_STRIDE, rc, lane, pol);
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

// ---------------------------------------------------------------------------------------------
// Base model of the decoder (BASE_MODEL<uint8_t>, decode inlined four times in decode_seq @0x4296b0).
// Decoder-private context word: c1 | c2 << 8 | c3 << 16 | tot << 24 -- the inclusive cumulative counts of
// A, A+C, A+C+G and the total (<= 254: a context is halved when its total passes 253 BEFORE a base is coded,
// @0x421fb0, then one counter is incremented).  Thresholds, the coded interval and the update (+1 on every
// byte at or above the base) come straight out of the word without unpacking four counters.
// The 4^SQ_LOOK contexts reachable SQ_LOOK bases ahead are contiguous in the table: every step the warp
// copies them asynchronously (cp.async, L2 -> shared, bypassing L1) into an 8-slot ring and reads the word it
// needs NOW from the slot filled SQ_LOOK steps ago, so the table's HBM latency (the table is 4 MiB per block,
// random access: it never stays in cache) is off the symbol chain.  Words updated inside the look-ahead
// window are forwarded from a register log.  Every lane stores the updated word (same value, same address):
// each lane's later copies are ordered after its own stores.
#ifndef SQ_LOOK
#define SQ_LOOK 3
#endif
#define SQ_LOOK_N (1 << (2 * SQ_LOOK))          // contexts per ring slot
#define SQ_LOG 4                                // forwarded updates (>= SQ_LOOK + 1)
static_assert(SQ_LOOK >= 2 && SQ_LOOK <= 3, "ring copy uses 16-byte pieces on up to 16 lanes");
struct SeqPipe {
    u32 step;
    u32 lc0, lc1, lc2, lc3, lv0, lv1, lv2, lv3;      // last four updated (context, word), 0 = newest
    SQ_D void init() { step = 0; lc0 = lc1 = lc2 = lc3 = 0xffffffffu; lv0 = lv1 = lv2 = lv3 = 0; }
};

SQ_D u32 base_halve(u32 x);
__global__ void k_init_seqtab(DecBlock *db, u64 nctx, const u8 *snap) {
    int b = blockIdx.y;
    if (db[b].status != 0) return;
    u32 *t = (u32 *)db[b].seqtab;
    const u32 *s = (const u32 *)snap;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nctx; i += (u64)gridDim.x * blockDim.x) {
        u32 v = s ? s[i] : 0x03030303u;
        u32 s0 = v & 0xff, s1 = (v >> 8) & 0xff, s2 = (v >> 16) & 0xff, s3 = v >> 24;
        u32 c1 = s0, c2 = c1 + s1, c3 = c2 + s2, tot = c3 + s3;     // <= 254 for every snapshot sq_load_model accepts
        u32 w = (c1 & 0xff) | ((c2 & 0xff) << 8) | ((c3 & 0xff) << 16) | ((tot & 0xff) << 24);
        t[i] = (w >> 24) > 253 ? base_halve(w) : w;
    }
}
// a snapshot whose base contexts sum above 254 cannot come from the reference's training (same update rule)
__global__ void k_check_seqsnap(const u32 *snap, u64 nctx, int *bad) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nctx; i += (u64)gridDim.x * blockDim.x) {
        u32 v = snap[i];
        if ((v & 0xff) + ((v >> 8) & 0xff) + ((v >> 16) & 0xff) + (v >> 24) > 254u) *bad = 1;
    }
}

// The reference halves a context whose total has passed 253 when it next CODES with it (@0x421fb0).  Nothing
// else reads the context in between, so the decoder halves when it WRITES the word (and when it builds the
// table): the test leaves the symbol chain.  S -= S >> 1 on the four counters behind the cumulative bytes.
SQ_D u32 base_halve(u32 x) {
    u32 tot = x >> 24;
    u32 s0 = x & 0xff, s1 = ((x >> 8) - x) & 0xff, s2 = ((x >> 16) - (x >> 8)) & 0xff, s3 = (tot - (x >> 16)) & 0xff;
    s0 -= s0 >> 1; s1 -= s1 >> 1; s2 -= s2 >> 1; s3 -= s3 >> 1;
    u32 c2 = s0 + s1, c3 = c2 + s2;
    tot = (c3 + s3) & 0xff;
    return s0 | ((c2 & 0xff) << 8) | ((c3 & 0xff) << 16) | (tot << 24);
}

SQ_D u32 decode_base(u8 *seqtab, u32 *ring, const u32 *rlut, RcWin &rc, u32 &last, u32 mask, u32 nb, SeqPipe &P, int lane, u64 pol) {
    if (lane < SQ_LOOK_N / 4) {
        u32 basec = (last << (2 * SQ_LOOK)) & mask;
        const u8 *src = seqtab + ((u64)basec << 2) + lane * 16;
        u32 dst = (u32)__cvta_generic_to_shared(ring + ((P.step + SQ_LOOK) & 7) * SQ_LOOK_N + lane * 4);
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "l"(pol) : "memory");
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    u32 *S4 = (u32 *)(seqtab + (u64)last * 4);
    u32 x;
    if (nb >= SQ_LOOK) {
        asm volatile("cp.async.wait_group %0;\n" ::"n"(SQ_LOOK) : "memory");
        __syncwarp();
        x = ring[(P.step & 7) * SQ_LOOK_N + (last & (SQ_LOOK_N - 1))];
    } else {
        asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(x) : "l"(S4), "l"(pol) : "memory");
    }
    // forwarding: resolved while the ring read is in flight
    const bool h0 = P.lc0 == last, h1 = P.lc1 == last, h2 = P.lc2 == last, h3 = P.lc3 == last;
    const u32 fv = h0 ? P.lv0 : (h1 ? P.lv1 : (h2 ? P.lv2 : P.lv3));
    x = (h0 || h1 || h2 || h3) ? fv : x;
    const u32 tot = x >> 24;                         // <= 253: contexts are halved when they are written (base_halve)
    u32 r = __umulhi(rc.range, rlut[tot]);
    u32 rem = rc.range - r * tot;
    r += (rem >= tot) ? 1u : 0u;
    const u32 code = rc.code;
    const u32 c1 = x & 0xff, c2 = (x >> 8) & 0xff, c3 = (x >> 16) & 0xff;
    const u32 b = (u32)(code >= c1 * r) + (u32)(code >= c2 * r) + (u32)(code >= c3 * r);
    const u32 sh = 8 * b;
    const u32 cum = ((x << 8) >> sh) & 0xff, cnx = (x >> sh) & 0xff;
    rc.range = r;
    rc.decode(cum, (cnx - cum) & 0xff);
    u32 nv = x + (0x01010101u << sh);                // S[b]++: every cumulative byte at or above b
    if ((nv >> 24) > 253) nv = base_halve(nv);
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(S4), "r"(nv), "l"(pol) : "memory");
    __syncwarp();
    P.lc3 = P.lc2; P.lv3 = P.lv2; P.lc2 = P.lc1; P.lv2 = P.lv1; P.lc1 = P.lc0; P.lv1 = P.lv0; P.lc0 = last; P.lv0 = nv;
    P.step++;
    last = (last * 4 + b) & mask;
    return b;
}

// The same step for the bulk of a read (its first SQ_LOOK bases go through decode_base, which reads the table
// directly): no look-ahead test, bytes picked
// with PRMT.  A warp issues at best one instruction every other cycle, so the count is what matters here.
SQ_D u32 lds32(u32 saddr) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory"); return v; }
// ring_s / rlut_s: shared-window addresses of the ring and the reciprocal table (computed once per warp);
// cp_src: seqtab + lane * 16, cp_dst: ring_s + lane * 16 (lanes below SQ_LOOK_N / 4 copy)
SQ_D u32 base_step_fast(u8 *seqtab, const u8 *cp_src, u32 cp_dst, u32 ring_s, u32 rlut_s, RcWin &rc, u32 &last, u32 mask, SeqPipe &P, int lane) {
    if (lane < SQ_LOOK_N / 4) {
        const u32 basec = (last << (2 * SQ_LOOK)) & mask;
#if SQ_POL_BASE
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(cp_dst + (((P.step + SQ_LOOK) & 7) * SQ_LOOK_N * 4)), "l"(cp_src + ((u64)basec << 2)), "l"(l2_policy_stream()) : "memory");
#else
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(cp_dst + (((P.step + SQ_LOOK) & 7) * SQ_LOOK_N * 4)), "l"(cp_src + ((u64)basec << 2)) : "memory");
#endif
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group %0;\n" ::"n"(SQ_LOOK) : "memory");
    __syncwarp();
    u32 x = lds32(ring_s + ((P.step & 7) * SQ_LOOK_N + (last & (SQ_LOOK_N - 1))) * 4);
    const bool h0 = P.lc0 == last, h1 = P.lc1 == last, h2 = P.lc2 == last, h3 = P.lc3 == last;
    const u32 fv = h0 ? P.lv0 : (h1 ? P.lv1 : (h2 ? P.lv2 : P.lv3));
    x = (h0 || h1 || h2 || h3) ? fv : x;
    const u32 tot = x >> 24;
    u32 r = __umulhi(rc.range, lds32(rlut_s + tot * 4));
    r += (rc.range - r * tot >= tot) ? 1u : 0u;
    const u32 code = rc.code;
    const u32 c1 = __byte_perm(x, 0, 0x4440), c2 = __byte_perm(x, 0, 0x4441), c3 = __byte_perm(x, 0, 0x4442);
    const bool p1 = code >= c1 * r, p2 = code >= c2 * r, p3 = code >= c3 * r;
    const u32 b = (p1 ? 1u : 0u) + (p2 ? 1u : 0u) + (p3 ? 1u : 0u);
    const u32 cum = p3 ? c3 : (p2 ? c2 : (p1 ? c1 : 0u));
    const u32 cnx = p3 ? tot : (p2 ? c3 : (p1 ? c2 : c1));
    rc.range = r;
    rc.decode(cum, cnx - cum);
    u32 nv = x + (0x01010101u << (8 * b));
    if ((nv >> 24) > 253) nv = base_halve(nv);
#if SQ_POL_BASE
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(seqtab + (u64)last * 4), "r"(nv), "l"(l2_policy_stream()) : "memory");
#else
    asm volatile("st.global.u32 [%0], %1;" ::"l"(seqtab + (u64)last * 4), "r"(nv) : "memory");
#endif
    P.lc3 = P.lc2; P.lv3 = P.lv2; P.lc2 = P.lc1; P.lv2 = P.lv1; P.lc1 = P.lc0; P.lv1 = P.lv0; P.lc0 = last; P.lv0 = nv;
    P.step++;
    last = (last * 4 + b) & mask;
    return b;
}

// base code -> letter without a table in local memory
SQ_D u32 base_char(u32 b) { return (0x54474341u >> (8 * b)) & 0xffu; }      // "ACGT"

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
    u32 ring[8 * SQ_LOOK_N];   // base-model candidate ring of the seq warp (8 slots x 4^SQ_LOOK contexts)
    u32 rlut[256];             // reciprocal of every possible base-model total (<= 255)
    NameSmem ns;
    SmModel<2> tip;
    SmModel<11> ch;
    SmModel<95> maxq;
    KModelD km;
    u32 cnts[5];               // n_tip, n_ch, n_mq, n_cnt, n_pos (published by lane 0 of the seq warp)
    int side_err;
    volatile u32 qprog;        // reads whose qualities are complete
    volatile u32 sprog;        // reads whose bases are complete (only read by the in-kernel MD5 chaser)
};

// ---------------------------------------------------------------------------------------------
// blockMd5Verify @0x414e00 inside the decode kernel: the names warp is done after ~10 % of the kernel's time;
// its lane 0 then digests the ID buffer and CHASES the quality and base warps, hashing their output 64 bytes at a
// time as whole reads are published (qprog / sprog), so that the digests are ready when the last read is.
struct Md5Run {
    u32 h[4];
    u64 pos;                     // bytes digested (multiple of 64 until the end)
    SQ_D void init() { h[0] = 0x67452301u; h[1] = 0xefcdab89u; h[2] = 0x98badcfeu; h[3] = 0x10325476u; pos = 0; }
    // digest whole chunks of p[pos, avail), at most `maxchunks`; returns true if it did any work
    SQ_D bool feed(const u8 *p, u64 avail, int maxchunks) {
        bool any = false;
        u32 w[16];
        const bool al = (((uintptr_t)p) & 15) == 0;
        while (pos + 64 <= avail && maxchunks-- > 0) {
            if (al) {
                const uint4 *p4 = (const uint4 *)(p + pos);
                uint4 x0 = __ldcg(p4), x1 = __ldcg(p4 + 1), x2 = __ldcg(p4 + 2), x3 = __ldcg(p4 + 3);
                w[0] = x0.x; w[1] = x0.y; w[2] = x0.z; w[3] = x0.w; w[4] = x1.x; w[5] = x1.y; w[6] = x1.z; w[7] = x1.w;
                w[8] = x2.x; w[9] = x2.y; w[10] = x2.z; w[11] = x2.w; w[12] = x3.x; w[13] = x3.y; w[14] = x3.z; w[15] = x3.w;
            } else {
                for (int k = 0; k < 16; k++) {
                    const u8 *b = p + pos + 4 * k;
                    w[k] = (u32)__ldcg(b) | ((u32)__ldcg(b + 1) << 8) | ((u32)__ldcg(b + 2) << 16) | ((u32)__ldcg(b + 3) << 24);
                }
            }
            md5_chunk(h, w);
            pos += 64;
            any = true;
        }
        return any;
    }
    // the last (< 64) bytes + padding; `total` = length of the whole buffer
    SQ_D void finish(const u8 *p, u64 total, u8 out[16]) {
        u8 tail[128];
        u32 rem = (u32)(total - pos);
        for (int k = 0; k < 128; k++) tail[k] = 0;
        for (u32 k = 0; k < rem; k++) tail[k] = __ldcg(p + pos + k);
        tail[rem] = 0x80;
        u32 tl = rem < 56 ? 64 : 128;
        u64 bits = total * 8;
        for (int k = 0; k < 8; k++) tail[tl - 8 + k] = (u8)(bits >> (8 * k));
        u32 w[16];
        for (u32 o = 0; o < tl; o += 64) {
            for (int k = 0; k < 16; k++) w[k] = get_u32le(tail + o + 4 * k);
            md5_chunk(h, w);
        }
        for (int k = 0; k < 4; k++) put_u32le(out + 4 * k, h[k]);
    }
};
SQ_D bool md5_equal(const u8 *a, const u8 *b) {
    bool ok = true;
    for (int k = 0; k < 16; k++) ok = ok && (a[k] == b[k]);
    return ok;
}
// lane 0 of the names warp, after the names are done.  Never waits on a warp that has given up: a block whose
// status turned non-zero is left alone.
SQ_D void md5_chase(DecBlock &d, volatile u32 *qprog, volatile u32 *sprog, const u32 *seqoff, u32 nreads, int id_rule) {
    volatile i32 *status = (volatile i32 *)&d.status;
    u8 dig[16];
    if (!id_rule && d.has_md5[0]) {
        md5_buffer(d.ids, d.idbytes, dig);
        if (!md5_equal(dig, d.md5[0]) && *status == 0) *status = -8;          // SQ_ERR_MD5_NAME
    }
    if (!d.has_md5[1] && !d.has_md5[2]) return;
    const u64 total = d.seqbytes;
    Md5Run rs, rq;
    rs.init(); rq.init();
    bool want_s = d.has_md5[1] != 0, want_q = d.has_md5[2] != 0;
    u64 as = 0, aq = 0;                               // bytes of reads published so far
    for (;;) {
        bool done_s = !want_s || (as == total && rs.pos + 64 > total);
        bool done_q = !want_q || (aq == total && rq.pos + 64 > total);
        if (done_s && done_q) break;
        if (*status != 0) return;
        u32 ps = *sprog, pq = *qprog;
        __threadfence_block();
        as = seqoff[ps < nreads ? ps : nreads];
        aq = seqoff[pq < nreads ? pq : nreads];
        bool any = false;
        if (want_s) any |= rs.feed(d.seq, as, 64);
        if (want_q) any |= rq.feed(d.qual, aq, 64);
        if (!any) __nanosleep(2000);
    }
    if (want_s) { rs.finish(d.seq, total, dig); if (!md5_equal(dig, d.md5[1]) && *status == 0) *status = -9; }
    if (want_q) { rq.finish(d.qual, total, dig); if (!md5_equal(dig, d.md5[2]) && *status == 0) *status = -10; }
}

#define SQ_DEC_THREADS 96
template <int QLEVEL>
__global__ void __launch_bounds__(SQ_DEC_THREADS) k_dec_main(DecBlock *db, int id_rule, u32 seqmask, int inline_md5, u32 qdummy) {
    constexpr int qlevel = QLEVEL;
    __shared__ __align__(16) MainSmem sm;
    DecBlock &d = db[blockIdx.x];
    int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { sm.qprog = 0; sm.sprog = 0; sm.side_err = 0; }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm.rlut[i] = recip32(i);
    __syncthreads();
    const u64 pol = wid == 2 ? l2_policy_stream() : l2_policy_keep();
    if (d.status != 0) return;
#ifdef SQ_DEBUG_TIMING
    long long t_begin = clock64();
#define SQ_TICK(name) if (blockIdx.x == 0 && lane == 0) printf("dec warp %d %s: %lld cycles\n", wid, name, clock64() - t_begin)
#else
#define SQ_TICK(name)
#endif
    const u32 nreads = d.nreads;
#ifdef SQ_DBG_SKIP                                               // timing experiments only: drop a role (the output is then wrong)
    if (wid == 2 && (SQ_DBG_SKIP & 1)) { if (lane == 0) sm.sprog = nreads; return; }
    if (wid == 0 && (SQ_DBG_SKIP & 2)) return;
#endif
    const u8 *rec = d.rec;
    const u32 *seqoff = d.seqoff;
    const i32 *seqlen = d.seqlen;
    if (wid == 0) {                                            // names (decompressName @0x428ca0)
        if (id_rule) {                                         // rule-based IDs (IDProcess::decodeIDS) are regenerated by the host layer from the tag-5 payload
            if (lane == 0) { d.idbytes = 0; if (inline_md5) md5_chase(d, &sm.qprog, &sm.sprog, seqoff, nreads, id_rule); }
            return;
        }
        for (int i = lane; i < 1024; i += 32) sm.ns.last[i] = ' ';
        __syncwarp();
        int lnl = 0, lpl = 0, lsl = 0;
        RcDecW rc;
        rc.start(rec + d.off[5], rec + d.off[5] + d.len[5]);
        u64 o = 0;
        u32 *nametab = d.nametab;
        u8 *ids = d.ids;
        u16 *idlen = d.idlen;
        const u64 idcap = d.idcap;
        for (u32 r = 0; r < nreads; r++) {
            int n = decode_name(nametab, rc, ids + o, idcap - o, sm.ns, lnl, lpl, lsl, lane, pol);
            if (n < 0) { if (lane == 0) d.status = SQ_ERR_CAPACITY_K; return; }
            if (lane == 0) idlen[r] = (u16)n;
            o += (u32)n;
        }
        if (lane == 0) d.idbytes = o;
        SQ_TICK("names");
        if (inline_md5 && lane == 0) md5_chase(d, &sm.qprog, &sm.sprog, seqoff, nreads, id_rule);
    } else if (wid == 1) {                                     // qualities (decompressQual @0x42aa40, decode_qual @0x42a750)
        RcWin rc;
        rc.start(rec + d.off[7], rec + d.off[7] + d.len[7]);
        uint4 *tab = (uint4 *)d.qualtab + q4_flane(lane);            // front words: see q4_pstep
        u8 *qual = d.qual;
        const u32 dummy_ctx = qdummy;                          // spare model slot behind the table: target of "no pending update"
        u32 last = 0;
        uint4 M = q4_ld(tab);                                  // context 0 opens every read
        QPend P;
        q4_pend_none(P, dummy_ctx);
        u32 off_n = nreads ? seqoff[0] : 0;
        int len_n = nreads ? seqlen[0] : 0;
        for (u32 r = 0; r < nreads; r++) {
            u8 *q = qual + off_n;
            const int len = len_n;
            if (r + 1 < nreads) { off_n = seqoff[r + 1]; len_n = seqlen[r + 1]; }     // fetched a read ahead
            u32 delta = 5, q1 = 0, q2 = 0, qq = 0;
            int i = 0;
            if (qlevel > 1)
                for (; i < len && delta < 56 && qq != 94; i++)
                    qq = q4_pstep<QLEVEL, false>(M, last, P, rc, tab, q + i, i + 1 < len, i, q1, q2, delta, lane, dummy_ctx);
            for (; i < len && qq != 94; i++)
                qq = q4_pstep<QLEVEL, true>(M, last, P, rc, tab, q + i, i + 1 < len, i, q1, q2, delta, lane, dummy_ctx);
            if (qq == 94) {                                    // the rest of the read is '#' (@0x42a9a0); context 0 opens the next read
                for (int t = i - 1 + lane; t < len; t += 32) q[t] = '#';
                q4_apply(tab, P, lane);
                q4_pend_none(P, dummy_ctx);
                if (last != 0) { M = q4_ld(tab); last = 0; }
            }
            if ((r & 15) == 15 || r + 1 == nreads) {
                __syncwarp();
                __threadfence_block();
                if (lane == 0) sm.qprog = r + 1;
            }
        }
        SQ_TICK("qual");
    } else if (wid == 2) {                                     // Dege streams, then bases (decompressSeq @0x42e390)
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
        SQ_TICK("dege");
        u32 n_tip = sm.cnts[0], n_ch = sm.cnts[1], n_mq = sm.cnts[2], n_cnt = sm.cnts[3], n_pos = sm.cnts[4];
        // the Dege arrays were written by lane 0: the other lanes read them through L2
        const u8 *d_tip = d.d_tip, *d_ch = d.d_ch, *d_maxq = d.d_maxq;
        const u32 *d_cnt = d.d_cnt, *d_pos = d.d_pos;
        u8 *seqtab = d.seqtab, *seqout = d.seq;
        const u8 *qualbuf = d.qual;
        RcWin rc;
        rc.start(rec + d.off[6], rec + d.off[6] + d.len[6]);
        SeqPipe P;
        P.init();
        const u32 ring_s = (u32)__cvta_generic_to_shared(sm.ring), rlut_s = (u32)__cvta_generic_to_shared(sm.rlut);
        const u8 *cp_src = seqtab + lane * 16;
        u32 smask = seqmask;
        asm volatile("mov.u32 %0, %0;" : "+r"(smask));           // keep the mask in a register (else it is re-read from the parameter bank every base)
        const u32 cp_dst = ring_s + lane * 16;
        u32 i_mq = 0, i_ch = 0, i_pos = 0;
        for (u32 r = 0; r < nreads; r++) {
            if (inline_md5 && r && (r & 15) == 0) {             // reads [0, r) are final: let the MD5 chaser have them
                __threadfence_block();
                if (lane == 0) sm.sprog = r;
            }
            u8 *seq = seqout + seqoff[r];
            int len = seqlen[r];
            u32 last = 0x7616c7u & seqmask;
            u32 nb = 0;
            bool flagged = r < n_tip && __ldcg(d_tip + r);
            if (!flagged) {
                int i = 0;
                for (; i < len && i < SQ_LOOK; i++) {
                    u32 bb = decode_base(seqtab, sm.ring, sm.rlut, rc, last, seqmask, nb++, P, lane, pol);
                    if (lane == 0) seq[i] = base_char(bb);
                }
                for (; i < len; i++) {
                    u32 bb = base_step_fast(seqtab, cp_src, cp_dst, ring_s, rlut_s, rc, last, smask, P, lane);
                    if (lane == 0) seq[i] = base_char(bb);
                }
                continue;
            }
            while (sm.qprog <= r) __nanosleep(200);            // this read's qualities place its N bases
            __threadfence_block();
            const u8 *qual = qualbuf + seqoff[r];
            if (i_mq >= n_mq) { if (lane == 0) d.status = SQ_ERR_FORMAT_K; return; }
            int mq = (int)(int8_t)__ldcg(d_maxq + i_mq);
            u32 cnt = i_mq < n_cnt ? __ldcg(d_cnt + i_mq) : 0;
            i_mq++;
            if (cnt == 0) {                                    // @0x429a20: every candidate position is non-ACGT
                for (int i = 0; i < len; i++) {
                    u32 c;
                    if (mq >= (int)(int8_t)__ldcg(qual + i)) { c = i_ch < n_ch ? __ldcg(d_ch + i_ch) : 'N'; i_ch++; }
                    else c = base_char(decode_base(seqtab, sm.ring, sm.rlut, rc, last, seqmask, nb++, P, lane, pol));
                    if (lane == 0) seq[i] = (u8)c;
                }
                continue;
            }
            u32 gap = i_pos < n_pos ? __ldcg(d_pos + i_pos) : 0, remaining = cnt - 1, run = 0;   // @0x4298cc
            i_pos++;
            for (int i = 0; i < len; i++) {
                u32 c;
                if (mq >= (int)(int8_t)__ldcg(qual + i)) {
                    if (run == gap) {                          // @0x429b40
                        gap = 0xffffffffu;
                        if (remaining) { remaining--; gap = i_pos < n_pos ? __ldcg(d_pos + i_pos) : 0; i_pos++; }
                        c = base_char(decode_base(seqtab, sm.ring, sm.rlut, rc, last, seqmask, nb++, P, lane, pol));
                        run = 0;
                    } else { run++; c = i_ch < n_ch ? __ldcg(d_ch + i_ch) : 'N'; i_ch++; }
                } else c = base_char(decode_base(seqtab, sm.ring, sm.rlut, rc, last, seqmask, nb++, P, lane, pol));
                if (lane == 0) seq[i] = (u8)c;
            }
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        if (inline_md5) {
            __threadfence_block();
            if (lane == 0) sm.sprog = nreads;
        }
        SQ_TICK("seq");
    }
}

// blockMd5Verify @0x414e00 ("Name/Seq/Qual md5 unequal", .rodata+0x15bc)
__global__ void __launch_bounds__(32) k_dec_md5(DecBlock *db, int B, int id_rule) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * B) return;
    int b = t / 3, which = t % 3;
    if (which == 0 && id_rule) return;           // names are regenerated (and checked) on the host in this mode
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

// sq_load_model: refuse base-context tables the decoder's 8-bit cumulative form cannot hold
cudaError_t sqk_check_seqsnap(const u8 *snap_seq, u64 nctx, int *d_flag, int *bad, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(d_flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    k_check_seqsnap<<<256, 256, 0, st>>>((const u32 *)snap_seq, nctx, d_flag);
    e = cudaMemcpyAsync(bad, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

cudaError_t sqk_decode(const DecWave &w, cudaStream_t st, u64 *launches) {
    int B = w.B;
    if (B == 0) return cudaSuccess;
    // per-block model reset (or snapshot restore) into the decoder's warp-model tables
    k_init_q4<95><<<dim3(w.nsm * 2, B), 256, 0, st>>>(w.db, w.nqual + 1, w.snap_qual, 293, w.nqual);      // + the spare slot (QPend)
    k_init_seqtab<<<dim3(w.nsm, B), 256, 0, st>>>(w.db, (u64)w.nseq, w.snap_seq);
    *launches += 2;
    if (!w.id_rule) {
        // nametab: prefix | suffix | len (256 x SIMPLE_MODEL<256> each) then 8192 x SIMPLE_MODEL<128>
        k_init_wm<256, 8><<<dim3(12, B), 256, 0, st>>>(w.db, 1, 768, 0, w.snap_name, 776);
        k_init_wm<128, 4><<<dim3(32, B), 256, 0, st>>>(w.db, 1, 8192, 768 * W256_STRIDE,
                                                     w.snap_name ? w.snap_name + 3ull * 256 * 776 : nullptr, 392);
        *launches += 2;
    }
    {
        // The quality / name models are re-read through L1 (measured: bypassing it costs 20 %): ask for
        // the smallest shared-memory carve-out that still holds this launch's co-resident CTAs.
        int dev = 0, smem_sm = 228 * 1024;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        int ctas = (B + w.nsm - 1) / (w.nsm > 0 ? w.nsm : 1);
        int need = ctas * ((int)sizeof(MainSmem) + 1024 + 256);
        int carve = (int)((100ll * need + smem_sm - 1) / smem_sm);
        if (carve > 100) carve = 100;
        {
            cudaFuncSetAttribute(k_dec_main<1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(k_dec_main<2>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            cudaFuncSetAttribute(k_dec_main<3>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        }
    }
    k_dec_len<<<B, 32, 0, st>>>(w.db);
    // The MD5s are verified inside k_dec_main (lane 0 of the names warp chases the other two warps); k_dec_md5 is the
    // separate pass kept for builds with -DSQ_DEC_SEPARATE_MD5.
#ifdef SQ_DEC_SEPARATE_MD5
    const int inline_md5 = 0;
#else
    const int inline_md5 = w.md5 ? 1 : 0;
#endif
    if (w.qlevel <= 1) k_dec_main<1><<<B, SQ_DEC_THREADS, 0, st>>>(w.db, w.id_rule, w.seqmask, inline_md5, w.nqual);
    else if (w.qlevel == 2) k_dec_main<2><<<B, SQ_DEC_THREADS, 0, st>>>(w.db, w.id_rule, w.seqmask, inline_md5, w.nqual);
    else k_dec_main<3><<<B, SQ_DEC_THREADS, 0, st>>>(w.db, w.id_rule, w.seqmask, inline_md5, w.nqual);
    *launches += 2;
    if (w.md5 && !inline_md5) { k_dec_md5<<<cdivd(3ull * B, 32), 32, 0, st>>>(w.db, B, w.id_rule); *launches += 1; }
    return cudaGetLastError();
}
