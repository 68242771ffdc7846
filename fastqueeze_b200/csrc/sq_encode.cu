// sq_encode.cu -- CUDA kernels of the block-united encoder (EncapFqzComp::doFqzEncode @0x42d2d0).
//
// Design: on ENCODE every context index is a pure function of the input symbols
// (seq @0x422090, qual @0x422474-0x422505, name @0x421578-0x4215a4), never of coder state, so a
// block's encode is decomposed into
//   1. context kernels      : (context key, symbol, stream position) per coded symbol, coalesced
//   2. stable radix sort    : group symbols by context inside each block, stream order kept
//   3. model replay         : one thread (bases) or one warp (SIMPLE_MODEL) per non-empty context
//                             replays that context's adaptive model over its own symbols and emits
//                             (cum, freq, tot) per symbol
//   4. range-coder pass     : the only serial part -- one warp per (block, stream) runs the
//                             64-bit low / 32-bit range recurrence over the triples
//   5. pack                 : tags, sizes, MD5s and side streams are laid out as the block record
// All kernels work on a WAVE of independent blocks at once (models reset per block).
#include "sq_common.cuh"
#include "sq_kernels.h"
#include "sq_md5.cuh"

#define FULL 0xffffffffu

// =============================================================================================
// helpers
SQ_D int find_block(const u32 *rbase, int B, u32 r) {       // rbase[0..B]: first read of each block
    int lo = 0, hi = B;                                      // invariant rbase[lo] <= r < rbase[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (rbase[mid] <= r) lo = mid; else hi = mid;
    }
    return lo;
}
SQ_D u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
SQ_D u32 warp_incl_scan(u32 v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}

// =============================================================================================
// 0. per-read preparation: lengths -> u32 arrays (scanned in place afterwards), validation
__global__ void k_read_prep(const BlockDesc *bd, BlockState *bs, const u32 *rbase, int B, u32 R,
                            u32 *seqoff, u32 *idoff) {
    u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= R) return;
    int b = find_block(rbase, B, w);
    u32 r = w - rbase[b];
    i32 sl = bd[b].seqlen[r];
    u32 il = bd[b].idlen[r];
    if (sl < 0 || (sl > 65535 && !bd[b].longflag) || il > 255) { bs[b].status = SQ_ERR_INPUT_K; sl = 0; il = 0; }
    seqoff[w] = (u32)sl;
    idoff[w] = il;
}

// the lengths must add up to the buffer sizes the caller gave; a block that fails is skipped by
// every later kernel (status != 0) so that no kernel reads outside the caller's buffers
__global__ void k_check_totals(const BlockDesc *bd, BlockState *bs, const u32 *tot_seq, const u32 *tot_id, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (tot_seq[b] != bd[b].seqbytes || tot_id[b] != bd[b].idbytes) bs[b].status = SQ_ERR_INPUT_K;
}

// In-place exclusive scan of arr[rbase[b] .. rbase[b+1]) for every block b (grid.x = B, grid.y =
// number of arrays); the block total goes to totals[y * B + b].
struct ScanArgs { u32 *arr[6]; };
__global__ void __launch_bounds__(1024) k_excl_scan(ScanArgs a, const u32 *rbase, u32 *totals, int B) {
    __shared__ u32 wsum[32];
    __shared__ u32 carry_s;
    int b = blockIdx.x;
    u32 *arr = a.arr[blockIdx.y] + rbase[b];
    u32 n = rbase[b + 1] - rbase[b];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int ITEMS = 8;
    for (u32 base = 0; base < n; base += 1024 * ITEMS) {
        u32 v[ITEMS], s = 0;
        u32 i0 = base + threadIdx.x * ITEMS;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) { v[k] = (i0 + k < n) ? arr[i0 + k] : 0; s += v[k]; }
        u32 inc = warp_incl_scan(s, lane);
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) { u32 t = wsum[lane]; u32 ti = warp_incl_scan(t, lane); wsum[lane] = ti - t; }
        __syncthreads();
        u32 excl = carry_s + wsum[wid] + inc - s;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) { if (i0 + k < n) arr[i0 + k] = excl; excl += v[k]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.y * B + b] = carry_s;
}

// =============================================================================================
// 1a. sequence: count coded (ACGT) bases per read / emit (context, base) elements.
// encode_seq @0x421f30: last = 0x7616c7 & mask at each read start; non-ACGT bases are skipped
// (@0x421f6d-0x421f7d) and do not enter the context; last = (last*4 + b) & mask (@0x422090).
template <bool EMIT>
__global__ void k_seq_pass(const BlockDesc *bd, BlockState *bs, const u32 *rbase, int B, u32 R,
                           const u32 *seqoff, u32 *acgt, u8 *tip, u64 *elem, u32 mask) {
    u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= R) return;
    int b = find_block(rbase, B, w);
    if (bs[b].status != 0) return;
    u32 r = w - rbase[b];
    const u8 *s = bd[b].seq + seqoff[w];
    u32 len = (u32)bd[b].seqlen[r];
    u64 prev = 0x7616c7u & mask;
    u32 done = 0;
    u32 base = EMIT ? acgt[w] : 0;
    u64 *dst = EMIT ? elem + bd[b].ebase[ST_SEQ] : nullptr;
    u32 lt = lanemask_lt();
    for (u32 i0 = 0; i0 < len; i0 += 32) {
        u32 i = i0 + lane;
        u32 v = 4;
        if (i < len) v = seq_val(s[i]);
        bool valid = v < 4;
        u32 bal = __ballot_sync(FULL, valid);
        u32 cnt = __popc(bal);
        if (EMIT && cnt) {
            u32 rk = __popc(bal & lt);
            // W: the chunk's coded bases, earliest at the top (base with rank j at bits 63-2j..62-2j)
            u64 x = valid ? ((u64)v << (62 - 2 * rk)) : 0;
            u32 whi = __reduce_or_sync(FULL, (u32)(x >> 32));
            u32 wlo = __reduce_or_sync(FULL, (u32)x);
            u64 W = ((u64)whi << 32) | wlo;
            if (valid) {
                u64 h = rk ? ((prev << (2 * rk)) | (W >> (64 - 2 * rk))) : prev;
                u32 pos = base + done + rk;
                dst[pos] = el_make((u32)h & mask, v, pos);
            }
            prev = (cnt == 32) ? W : ((prev << (2 * cnt)) | (W >> (64 - 2 * cnt)));
        }
        done += cnt;
    }
    if (!EMIT && lane == 0) {
        acgt[w] = done;
        tip[w] = (done != len);
        if (done != len) atomicAdd(&bs[b].side_cnt[3], 1u);      // flagged reads (= DegeMaxQual count)
    }
}

// 1b. quality: encode_qual @0x422180.  COUNT: strip trailing '#' (len2), symbols = len2 + (len2 != len).
// EMIT: symbol i is coded in context ctx_i = f(q[i-1], q[i-2], q[i-3], delta[i-1]); the optional
// terminator (symbol 94) is position len2.
template <bool EMIT>
__global__ void k_qual_pass(const BlockDesc *bd, BlockState *bs, const u32 *rbase, int B, u32 R,
                            const u32 *seqoff, u32 *qlen2, u32 *qoff, u64 *elem, int qlevel) {
    u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= R) return;
    int b = find_block(rbase, B, w);
    if (bs[b].status != 0) return;
    u32 r = w - rbase[b];
    const u8 *q = bd[b].qual_work + seqoff[w];
    u32 len = (u32)bd[b].seqlen[r];
    if (!EMIT) {
        u32 len2 = 0;
        bool bad = false;
        for (u32 i0 = 0; i0 < len; i0 += 32) {
            u32 i = i0 + lane;
            u32 c = (i < len) ? q[i] : (u32)'#';
            if (i < len && (u8)(c - 33) > 93) bad = true;      // '!' .. '~': symbol 94 is the read terminator
            u32 m = __ballot_sync(FULL, c != '#');
            if (m) len2 = i0 + 32 - __clz(m);
        }
        if (__any_sync(FULL, bad) && lane == 0) bs[b].status = SQ_ERR_INPUT_K;
        if (lane == 0) { qlen2[w] = len2; qoff[w] = len2 + (len2 != len); }
        return;
    }
    u32 len2 = qlen2[w];
    u32 n2 = len2 + (len2 != len);
    u32 base = qoff[w];
    u64 *dst = elem + bd[b].ebase[ST_QUAL];
    u32 pq1 = 0, pq2 = 0, pq3 = 0;      // q[i0-1], q[i0-2], q[i0-3]
    u32 pdelta = 5;                      // delta after symbol i0-1
    for (u32 i0 = 0; i0 < n2; i0 += 32) {
        u32 i = i0 + lane;
        u32 sym = 0;
        if (i < len2) sym = (u8)(q[i] - 33);
        else if (i < n2) sym = 94;
        u32 a1 = __shfl_up_sync(FULL, sym, 1), a2 = __shfl_up_sync(FULL, sym, 2), a3 = __shfl_up_sync(FULL, sym, 3);
        u32 q1 = lane >= 1 ? a1 : pq1;                       // q[i-1]
        u32 q2 = lane >= 2 ? a2 : (lane == 1 ? pq1 : pq2);   // q[i-2]
        u32 q3 = lane >= 3 ? a3 : (lane == 2 ? pq1 : (lane == 1 ? pq2 : pq3));   // q[i-3]
        u32 d = (i < len2 && q1 > sym) ? q1 - sym : 0;       // delta += (q1 > q) * (q1 - q)
        u32 dinc = warp_incl_scan(d, lane);
        u32 delta_i = pdelta + dinc;                         // delta after symbol i
        u32 dprev = __shfl_up_sync(FULL, delta_i, 1);
        if (lane == 0) dprev = pdelta;                       // delta after symbol i-1
        if (i < n2) {
            u32 ctx = (i == 0) ? 0 : qual_ctx(q1, q2, q3, dprev, qlevel, (int)i - 1);
            dst[base + i] = el_make(ctx, sym, base + i);
        }
        pq3 = __shfl_sync(FULL, sym, 29); pq2 = __shfl_sync(FULL, sym, 30); pq1 = __shfl_sync(FULL, sym, 31);
        pdelta = __shfl_sync(FULL, delta_i, 31);
    }
}

// 1c. names: encode_name @0x421070 (fqzcomp nlevel-1 name model).  The model
// contexts depend on the previous name only, so every read is independent given the input.
// last_name[] is never cleared in the reference (memcpy of the new name over the old one,
// @0x42452c): index j >= the previous length still holds the most recent earlier name that was
// longer than j, or ' ' from the per-block reset.
struct NameView {
    const u8 *ids; const u16 *idlen; const u32 *idoff;   // block-relative arrays of this block
    u32 r;                                               // current read
    SQ_D u32 last(int j) const {
        if (j < 0) return 0;                             // byte in front of last_name[] (@+0x1087) is 0
        for (i64 rr = (i64)r - 1; rr >= 0; rr--)
            if ((int)idlen[rr] > j) return ids[idoff[rr] + j];
        return ' ';
    }
};

template <bool EMIT>
__global__ void k_name_pass(const BlockDesc *bd, BlockState *bs, const u32 *rbase, int B, u32 R,
                            const u32 *idoff, u8 *np, u8 *ns, u32 *noff, u64 *elem) {
    u32 w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= R) return;
    int b = find_block(rbase, B, w);
    if (bs[b].status != 0) return;
    u32 r = w - rbase[b];
    NameView nv{bd[b].ids, bd[b].idlen, idoff + rbase[b], r};
    const u8 *name = nv.ids + nv.idoff[r];
    int n = nv.idlen[r];
    int lnl = r ? (int)nv.idlen[r - 1] : 0;
    int p, s;
    if (!EMIT) {
        p = 0; s = 0;
        if (n > 0 && lnl > 0) {
            const u8 *last = nv.ids + nv.idoff[r - 1];
            if (name[0] == last[0]) { p = 1; while (p < n && p < lnl && name[p] == last[p]) p++; }
            int i = n - 1, j = lnl - 1;
            while (i >= 0 && j >= 0 && name[i] == last[j]) { i--; j--; }
            s = n - 1 - i;
        }
        if (n - s - p < 0) s = n - p;
        np[w] = (u8)p; ns[w] = (u8)s;
        noff[w] = 3 + (u32)(n - s - p);
        return;
    }
    p = np[w]; s = ns[w];
    u32 o = noff[w];
    u64 *dst = elem + bd[b].ebase[ST_NAME];
    u32 lpl = r ? np[w - 1] : 0, lsl = r ? ns[w - 1] : 0;
    dst[o] = el_make(NK_PREFIX + lpl, (u32)p, o); o++;
    dst[o] = el_make(NK_SUFFIX + lsl, (u32)s, o); o++;
    dst[o] = el_make(NK_LEN + (u32)lnl, (u32)n, o); o++;
    int len2 = n - s, lc2 = (p != 0), i = p, j = p, k = 0;
    bool bad = false;
    while (i < len2) {
        u32 lj = nv.last(j);
        int v = ((int)lj - 32) * 2 + lc2 + k * 64;
        int ctx = v % 8192;
        if (ctx < 0) { bad = true; ctx = 0; }
        u32 ch = name[i];
        if (ch & 0x80) bad = true;                      // the name model codes ch & 0x7f (@0x4215b0): bit 7 would not survive
        dst[o] = el_make(NK_MIDDLE + (u32)ctx, ch & 0x7f, o); o++;
        if (ch == ' ' && nv.last(j) != ' ') j++;
        if (ch != ' ' && nv.last(j) == ' ') j--;
        if (ch == ':' && nv.last(j) != ':') j++;
        if (ch != ':' && nv.last(j) == ':') j--;
        if (ch == ':' || ch == ' ') k = (k + 3) & ~3;
        if (j < 0) { j = 0; lc2 = 0; }
        else { lc2 = (ch == nv.last(j)); j++; }
        k++; i++;
    }
    if (bad) bs[b].status = SQ_ERR_INPUT_K;
}

// record the per-stream symbol counts (totals of the scans) in BlockState
__global__ void k_set_counts(BlockState *bs, const u32 *totals, int B, int with_names) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // totals layout: [y * B + b]; y = 0 acgt, 1 qual symbols, 2 name symbols
    bool okb = bs[b].status == 0;
    bs[b].nsym[ST_SEQ] = okb ? totals[0 * B + b] : 0;
    bs[b].acgt = bs[b].nsym[ST_SEQ];
    bs[b].nsym[ST_QUAL] = okb ? totals[1 * B + b] : 0;
    bs[b].nsym[ST_NAME] = (okb && with_names) ? totals[2 * B + b] : 0;
}

// =============================================================================================
// 2. stable LSD radix sort by context key, segmented per block (BlockDesc::ebase / BlockState::nsym).
// One CTA of 512 threads owns one tile of SQ_TILE = 8192 consecutive elements.  A pass is
//   k_sort_hist     per-tile digit counts                           -> hist[block][digit][tile]
//   k_scan_*        exclusive scan of each block's [digit][tile] table (three small parallel kernels)
//   k_sort_scatter  the tile is ranked (warp: __match_any_sync, so key skew costs nothing; CTA: per-warp counters
//                   scanned per digit), written to SHARED memory in sorted order, and copied out from there:
//                   consecutive threads write consecutive elements of one digit's run, so the global stores are
//                   coalesced runs (8 elements = 64 B on average for 10-bit digits of the uniformly distributed
//                   base contexts, 32 elements for 8-bit digits) instead of one 8-byte store per sector.
// hist layout per block: [digit][tile] (tile-minor), block stride = D * maxtiles.
#define SQ_SORT_WARPS (SQ_SORT_THREADS / 32)
#define SQ_SORT_ITEMS (SQ_TILE / SQ_SORT_THREADS)            // elements per thread (16)
static_assert(SQ_TILE == SQ_SORT_THREADS * SQ_SORT_ITEMS && SQ_SORT_ITEMS == 16, "tile shape");

template <int BITS>
__global__ void __launch_bounds__(SQ_SORT_THREADS) k_sort_hist(const u64 *src, u32 *hist, const BlockDesc *bd,
                                                               const BlockState *bs, int stream, int shift, u32 maxtiles) {
    constexpr int D = 1 << BITS;
    __shared__ u32 cnt[D];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const u32 tile = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += SQ_SORT_THREADS) cnt[d] = 0;
    __syncthreads();
    const u32 n = bs[b].nsym[stream];
    const u64 t0 = (u64)tile * SQ_TILE;
    if (t0 < n) {
        const u32 m = (u32)min((u64)SQ_TILE, (u64)n - t0);
        const u64 *p = src + bd[b].ebase[stream] + t0;
        const u32 lt = lanemask_lt();
#pragma unroll 4
        for (int it = 0; it < SQ_SORT_ITEMS; it++) {
            const u32 i = (u32)wid * (SQ_TILE / SQ_SORT_WARPS) + it * 32 + lane;
            const bool valid = i < m;
            const u32 act = __ballot_sync(FULL, valid);
            if (valid) {
                const u32 dg = (u32)(p[i] >> (40 + shift)) & (D - 1);
                const u32 peers = __match_any_sync(act, dg);
                if ((peers & lt) == 0) atomicAdd(&cnt[dg], (u32)__popc(peers));
            }
        }
    }
    __syncthreads();
    u32 *h = hist + (u64)b * D * maxtiles + tile;
    for (int d = threadIdx.x; d < D; d += SQ_SORT_THREADS) h[(u64)d * maxtiles] = cnt[d];
}

// exclusive scan of every block's [digit][tile] table, entries_per_block values each, in chunks of SQ_SCAN_CHUNK
#define SQ_SCAN_CHUNK 4096
__global__ void __launch_bounds__(256) k_scan_partial(const u32 *hist, u64 entries_per_block, u32 nchunks, u32 *part) {
    __shared__ u32 ws[8];
    const u32 *h = hist + (u64)blockIdx.y * entries_per_block;
    const u64 lo = (u64)blockIdx.x * SQ_SCAN_CHUNK;
    u32 s = 0;
    for (int k = 0; k < SQ_SCAN_CHUNK / 256; k++) { u64 i = lo + (u64)k * 256 + threadIdx.x; if (i < entries_per_block) s += h[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { u32 t = 0; for (int k = 0; k < 8; k++) t += ws[k]; part[(u64)blockIdx.y * nchunks + blockIdx.x] = t; }
}
__global__ void __launch_bounds__(1024) k_scan_mid(u32 *part, u32 nchunks) {
    __shared__ u32 wsum[32];
    __shared__ u32 carry;
    u32 *p = part + (u64)blockIdx.x * nchunks;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < nchunks; base += 1024) {
        const u32 i = base + threadIdx.x;
        const u32 v = i < nchunks ? p[i] : 0;
        const u32 inc = warp_incl_scan(v, lane);
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) { u32 t = wsum[lane]; u32 ti = warp_incl_scan(t, lane); wsum[lane] = ti - t; }
        __syncthreads();
        const u32 excl = carry + wsum[wid] + inc - v;
        if (i < nchunks) p[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_scan_final(u32 *hist, u64 entries_per_block, u32 nchunks, const u32 *part) {
    __shared__ u32 wsum[8];
    u32 *h = hist + (u64)blockIdx.y * entries_per_block;
    const u64 lo = (u64)blockIdx.x * SQ_SCAN_CHUNK + (u64)threadIdx.x * (SQ_SCAN_CHUNK / 256);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 v[SQ_SCAN_CHUNK / 256], s = 0;
#pragma unroll
    for (int k = 0; k < SQ_SCAN_CHUNK / 256; k++) { v[k] = lo + k < entries_per_block ? h[lo + k] : 0; s += v[k]; }
    const u32 inc = warp_incl_scan(s, lane);
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    u32 wbase = 0;
    for (int k = 0; k < wid; k++) wbase += wsum[k];
    u32 run = part[(u64)blockIdx.y * nchunks + blockIdx.x] + wbase + inc - s;
#pragma unroll
    for (int k = 0; k < SQ_SCAN_CHUNK / 256; k++) { if (lo + k < entries_per_block) h[lo + k] = run; run += v[k]; }
}

template <int BITS>
__global__ void __launch_bounds__(SQ_SORT_THREADS, 1) k_sort_scatter(const u64 *src, u64 *dst, const u32 *hist, const BlockDesc *bd,
                                                                     const BlockState *bs, int stream, int shift, u32 maxtiles) {
    constexpr int D = 1 << BITS;
    extern __shared__ __align__(16) u8 smem_raw[];
    u64 *sorted = (u64 *)smem_raw;                                   // [SQ_TILE]
    u32 *ls = (u32 *)(sorted + SQ_TILE);                              // [D] tile-local start of each digit's run
    u32 *gb = ls + D;                                                 // [D] where that run goes in the block's array
    u16 *wcnt = (u16 *)(gb + D);                                      // [SQ_SORT_WARPS][D] per-warp counts -> per-warp bases
    __shared__ u32 wsum[SQ_SORT_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const u32 tile = blockIdx.x;
    const u32 n = bs[b].nsym[stream];
    const u64 t0 = (u64)tile * SQ_TILE;
    if (t0 >= n) return;
    const u32 m = (u32)min((u64)SQ_TILE, (u64)n - t0);
    const u64 *p = src + bd[b].ebase[stream] + t0;
    u64 *q = dst + bd[b].ebase[stream];
    for (int d = threadIdx.x; d < D * SQ_SORT_WARPS / 2; d += SQ_SORT_THREADS) ((u32 *)wcnt)[d] = 0;
    const u32 *h = hist + (u64)b * D * maxtiles + tile;
    for (int d = threadIdx.x; d < D; d += SQ_SORT_THREADS) gb[d] = h[(u64)d * maxtiles];
    __syncthreads();
    // 1. rank inside the warp's 512 consecutive elements (stream order: iteration, then lane)
    u64 e[SQ_SORT_ITEMS];
    u16 rk[SQ_SORT_ITEMS];
    u16 *mycnt = wcnt + wid * D;
    const u32 lt = lanemask_lt();
#pragma unroll
    for (int it = 0; it < SQ_SORT_ITEMS; it++) {
        const u32 i = (u32)wid * (SQ_TILE / SQ_SORT_WARPS) + it * 32 + lane;
        e[it] = i < m ? p[i] : 0;
    }
#pragma unroll
    for (int it = 0; it < SQ_SORT_ITEMS; it++) {
        const u32 i = (u32)wid * (SQ_TILE / SQ_SORT_WARPS) + it * 32 + lane;
        const bool valid = i < m;
        const u32 act = __ballot_sync(FULL, valid);
        rk[it] = 0;
        if (valid) {
            const u32 dg = (u32)(e[it] >> (40 + shift)) & (D - 1);
            const u32 peers = __match_any_sync(act, dg);
            const u32 r = __popc(peers & lt);
            const u32 base = mycnt[dg];
            __syncwarp(act);
            if (r == 0) mycnt[dg] = (u16)(base + __popc(peers));
            rk[it] = (u16)(base + r);
        }
        __syncwarp();
    }
    __syncthreads();
    // 2. per digit: exclusive prefix over the warps, tile total; then the exclusive scan of the totals over the digits
    u32 mine[D / SQ_SORT_THREADS > 0 ? D / SQ_SORT_THREADS : 1];
    u32 tsum = 0;
    constexpr int DPT = D / SQ_SORT_THREADS > 0 ? D / SQ_SORT_THREADS : 1;      // digits per thread (2 for D = 1024; D = 256: threads < 256 take one)
#pragma unroll
    for (int k = 0; k < DPT; k++) {
        const int d = threadIdx.x * DPT + k;
        u32 acc = 0;
        if (d < D) {
            for (int w = 0; w < SQ_SORT_WARPS; w++) { u32 t = wcnt[w * D + d]; wcnt[w * D + d] = (u16)acc; acc += t; }
        }
        mine[k] = acc;
        tsum += acc;
    }
    const u32 inc = warp_incl_scan(tsum, lane);
    if (lane == 31) wsum[wid] = inc;
    __syncthreads();
    u32 wbase = 0;
    for (int k = 0; k < wid; k++) wbase += wsum[k];
    u32 run = wbase + inc - tsum;
#pragma unroll
    for (int k = 0; k < DPT; k++) {
        const int d = threadIdx.x * DPT + k;
        if (d < D) ls[d] = run;
        run += mine[k];
    }
    __syncthreads();
    // 3. the tile in sorted order, in shared memory
#pragma unroll
    for (int it = 0; it < SQ_SORT_ITEMS; it++) {
        const u32 i = (u32)wid * (SQ_TILE / SQ_SORT_WARPS) + it * 32 + lane;
        if (i < m) {
            const u32 dg = (u32)(e[it] >> (40 + shift)) & (D - 1);
            sorted[ls[dg] + mycnt[dg] + rk[it]] = e[it];
        }
    }
    __syncthreads();
    // 4. out: position k of the sorted tile belongs to its digit's run at gb[digit] + (k - ls[digit])
    for (u32 k = threadIdx.x; k < m; k += SQ_SORT_THREADS) {
        const u64 x = sorted[k];
        const u32 dg = (u32)(x >> (40 + shift)) & (D - 1);
        q[gb[dg] + (k - ls[dg])] = x;
    }
}

// =============================================================================================
// 3a. list of context segments (heads) of the sorted element array
__global__ void k_seglist(const u64 *src, const BlockDesc *bd, const BlockState *bs, int stream,
                          u64 *seglist, u32 *nseg, u32 segcap) {
    int b = blockIdx.y;
    u32 n = bs[b].nsym[stream];
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31;
    if ((i & ~31u) >= n) return;
    u64 g = bd[b].ebase[stream] + i;
    bool head = false;
    if (i < n) head = (i == 0) || (el_key(src[g - 1]) != el_key(src[g]));
    u32 bal = __ballot_sync(FULL, head);
    if (!bal) return;
    u32 base = 0;
    if (lane == 0) base = atomicAdd(nseg, (u32)__popc(bal));
    base = __shfl_sync(FULL, base, 0);
    if (head) {
        u32 slot = base + __popc(bal & lanemask_lt());
        if (slot < segcap) seglist[slot] = ((u64)b << 40) | g;
    }
}

// warp-granular dynamic work distribution over the segment list
SQ_D u32 grab32(u32 *cursor, int lane) {
    u32 base = 0;
    if (lane == 0) base = atomicAdd(cursor, 32u);
    return __shfl_sync(FULL, base, 0);
}

// 3b. base model replay (fqzcomp BASE_MODEL<uint8_t>, @0x421f7f-0x422087): 4 byte counters per
// context, init 3 (or the snapshot's table), halve when the sum exceeds 253 BEFORE coding.
// One thread per context; the element stream of a context is contiguous after the sort, so the
// next element is fetched one step ahead and the line after next is prefetched into L1.
__global__ void __launch_bounds__(128) k_replay_seq(const u64 *src, const BlockDesc *bd, const BlockState *bs,
                                                    const u64 *seglist, const u32 *nseg, u32 *cursor,
                                                    u32 *triples, const u8 *snap_seq, u8 *train_out) {
    int lane = threadIdx.x & 31;
    u32 total = *nseg;
    for (;;) {
        u32 base = grab32(cursor, lane);
        if (base >= total) break;
        u32 idx = base + lane;
        if (idx >= total) continue;
        u64 ent = seglist[idx];
        int b = (int)(ent >> 40);
        u64 g = ent & (((u64)1 << 40) - 1);
        u64 eb = bd[b].ebase[ST_SEQ];
        u64 end = eb + bs[b].nsym[ST_SEQ];
        u64 e = src[g];
        u32 key = el_key(e);
        u32 S0 = 3, S1 = 3, S2 = 3, S3 = 3;
        if (snap_seq) { const u8 *t = snap_seq + (u64)key * 4; S0 = t[0]; S1 = t[1]; S2 = t[2]; S3 = t[3]; }
        u32 *out = triples + eb;
        bool hn = g + 1 < end;
        u64 en = hn ? src[g + 1] : 0;
        for (;;) {
            if ((g & 15) == 0 && g + 32 < end) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + g + 32));
            u32 sym = el_sym(e);
            u32 tot = S0 + S1 + S2 + S3;
            if (tot > 253) {
                S0 -= S0 >> 1; S1 -= S1 >> 1; S2 -= S2 >> 1; S3 -= S3 >> 1;
                tot = S0 + S1 + S2 + S3;
            }
            u32 cum, f;
            if (sym == 0) { cum = 0; f = S0; S0 = (S0 + 1) & 0xff; }
            else if (sym == 1) { cum = S0; f = S1; S1 = (S1 + 1) & 0xff; }
            else if (sym == 2) { cum = S0 + S1; f = S2; S2 = (S2 + 1) & 0xff; }
            else { cum = S0 + S1 + S2; f = S3; S3 = (S3 + 1) & 0xff; }
            if (!train_out) out[el_pos(e)] = tr_make32(cum, f, tot);
            if (!hn) break;
            e = en;
            g++;
            if (el_key(e) != key) break;
            hn = g + 1 < end;
            if (hn) en = src[g + 1];
        }
        if (train_out) {          // encode_seq_formodel @0x4220c0: the counters after the last sampled base
            u8 *t = train_out + (u64)key * 4;
            t[0] = (u8)S0; t[1] = (u8)S1; t[2] = (u8)S2; t[3] = (u8)S3;
        }
    }
}

// 3c. SIMPLE_MODEL<N> replay (encodeSymbol @0x4221fb-0x4222f1), one WARP per context, 16 symbols
// per step.  Between two bubble-swap checks (every 16th symbol of a context, @0x4222d3) the order of
// the table is fixed and every coded symbol adds exactly 8 to its own frequency, so for the symbols
// s_0..s_{n-1} (n <= 16) of such a batch, with k_t the table slot of s_t and (F0, C0, tot0) the
// frequencies / exclusive prefix sums / total at the start of the batch:
//     cum_t  = C0[k_t] + 8 * #{u < t : k_u <  k_t}
//     freq_t = F0[k_t] + 8 * #{u < t : k_u == k_t}
//     tot_t  = tot0 + 8 t
// Lane t computes step t: one table lookup and 16 shuffles instead of a serial scan + update per
// symbol.  The model lives in shared memory (order -> symbol, symbol -> slot, F, C).  A batch in which
// TotFreq would pass 0xffe0 (the halving, @0x4222b7, about every 4096th symbol) is replayed serially.
// keylo/keyhi select the contexts this launch handles; snap (may be null) + snapstride give the
// packed snapshot model of key keylo + k at snap + k * snapstride.
template <int N>
SQ_D u32 rb_prefix(const u32 *F, u32 *Cm, int lane) {          // C[k] = sum of F[0..k); returns the total
    constexpr int NP = (N + 31) / 32;
    u32 v[NP], sum = 0;
#pragma unroll
    for (int i = 0; i < NP; i++) { int k = lane * NP + i; v[i] = k < N ? F[k] : 0; sum += v[i]; }
    u32 inc = warp_incl_scan(sum, lane);
    u32 run = inc - sum;
#pragma unroll
    for (int i = 0; i < NP; i++) { int k = lane * NP + i; if (k < N) Cm[k] = run; run += v[i]; }
    return __shfl_sync(FULL, inc, 31);
}

template <int N>
__global__ void __launch_bounds__(128) k_replay_batch(const u64 *src, const BlockDesc *bd, const BlockState *bs, int stream,
                                                      const u64 *seglist, const u32 *nseg, u32 *cursor, u64 *triples,
                                                      u32 keylo, u32 keyhi, const u8 *snap, u32 snapstride, u8 *train_out) {
    __shared__ u32 s_f[4][256], s_c[4][256];
    __shared__ u8 s_pos[4][256], s_ord[4][256];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 *F = s_f[wid], *Cm = s_c[wid];
    u8 *pos = s_pos[wid], *ord = s_ord[wid];
    u32 total = *nseg;
    for (;;) {
        u32 idx = 0;
        if (lane == 0) idx = atomicAdd(cursor, 1u);
        idx = __shfl_sync(FULL, idx, 0);
        if (idx >= total) break;
        u64 ent = seglist[idx];
        int b = (int)(ent >> 40);
        u64 g = ent & (((u64)1 << 40) - 1);
        u32 key = el_key(src[g]);
        if (key < keylo || key >= keyhi) continue;
        u64 eb = bd[b].ebase[stream];
        u64 end = eb + bs[b].nsym[stream];
        u64 *out = triples + eb;
        u32 tot, bub;
        __syncwarp();
        if (snap) {
            const u8 *sp = snap + (u64)(key - keylo) * snapstride;
            tot = get_u32le(sp); bub = get_u32le(sp + 4);
            for (int k = lane; k < N; k += 32) {
                const u8 *q = sp + 8 + 3 * k;
                u32 sym = q[0];
                ord[k] = (u8)sym; pos[sym] = (u8)k; F[k] = (u32)(q[1] | (q[2] << 8));
            }
        } else {
            tot = N; bub = 0;
            for (int k = lane; k < N; k += 32) { ord[k] = (u8)k; pos[k] = (u8)k; F[k] = 1; }
        }
        __syncwarp();
        rb_prefix<N>(F, Cm, lane);
        __syncwarp();
        u64 en = (g + lane < end) ? src[g + lane] : ~(u64)0;         // lane i holds element g + i
        for (;;) {
            int room = 16 - (int)(bub & 15);                         // symbols up to and including the next swap check
            u64 gi = g + lane;
            u64 e = en;
            bool valid = lane < room && gi < end && el_key(e) == key;
            u32 vb = __ballot_sync(FULL, valid);
            int cnt = __ffs(~vb) - 1;                                // valid lanes form a prefix; room <= 16
            if (cnt == 0) break;
            {                                                        // next batch's elements, in flight during this batch
                u64 gn = g + (u64)cnt + lane;
                en = gn < end ? src[gn] : ~(u64)0;
            }
            u32 s = el_sym(e);
            if (tot + 8u * (u32)cnt <= 0xffe0u) {
                u32 k = 0, c0 = 0, f0 = 0;
                if (lane < cnt) { k = pos[s]; c0 = Cm[k]; f0 = F[k]; }
                u32 lt = 0, eq = 0;
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    u32 ku = __shfl_sync(FULL, k, u);
                    bool before = u < lane && u < cnt;
                    lt += (before && ku < k) ? 1u : 0u;
                    eq += (before && ku == k) ? 1u : 0u;
                }
                if (lane < cnt && !train_out) out[el_pos(e)] = tr_make(c0 + 8 * lt, f0 + 8 * eq, tot + 8 * (u32)lane);
                __syncwarp();
                if (lane < cnt) atomicAdd(&F[k], 8u);
                tot += 8u * (u32)cnt;
                bub += (u32)cnt;
                __syncwarp();
                rb_prefix<N>(F, Cm, lane);
                __syncwarp();
            } else {                                                 // the halving falls inside this batch: one symbol at a time
                for (int t = 0; t < cnt; t++) {
                    u32 st = __shfl_sync(FULL, s, t);
                    u32 k = pos[st];
                    u32 c0 = Cm[k], f0 = F[k];
                    if (lane == t && !train_out) out[el_pos(e)] = tr_make(c0, f0, tot);
                    __syncwarp();
                    if (lane == 0) F[k] = f0 + 8;
                    for (int j = (int)k + 1 + lane; j < N; j += 32) Cm[j] += 8;
                    tot += 8;
                    __syncwarp();
                    if (tot > 0xffe0u) {
                        for (int j = lane; j < N; j += 32) { u32 x = F[j]; x -= x >> 1; F[j] = x; }
                        __syncwarp();
                        tot = rb_prefix<N>(F, Cm, lane);
                        __syncwarp();
                    }
                    bub++;
                }
            }
            if ((bub & 15) == 0) {                                   // bubble swap with the predecessor if strictly greater
                u32 sl = __shfl_sync(FULL, s, cnt - 1);
                if (lane == 0) {
                    u32 k = pos[sl];
                    if (k > 0 && F[k] > F[k - 1]) {
                        u32 fk = F[k], fp = F[k - 1];
                        u8 ok = ord[k], op = ord[k - 1];
                        ord[k] = op; ord[k - 1] = ok; pos[op] = (u8)k; pos[ok] = (u8)(k - 1);
                        F[k] = fp; F[k - 1] = fk;
                        Cm[k] = Cm[k - 1] + fk;
                    }
                }
                __syncwarp();
            }
            g += (u64)cnt;
            if (cnt < room) break;
        }
        if (train_out) {          // SaveModelToMem @0x420c00: u32le Tot, u32le Bub, N x {u8 sym, u16le freq} in table order
            u8 *t = train_out + (u64)(key - keylo) * snapstride;
            if (lane == 0) { put_u32le(t, tot); put_u32le(t + 4, bub); }
            for (int k = lane; k < N; k += 32) { u8 *q = t + 8 + 3 * k; u32 f = F[k]; q[0] = ord[k]; q[1] = (u8)f; q[2] = (u8)(f >> 8); }
        }
        __syncwarp();
    }
}

// default (constructor-state) snapshot: what SaveModelToMem would write for untouched models
__global__ void k_snap_init(u8 *blob, u64 name_bytes, u32 nseq, u32 nqual) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 nname = name_bytes ? 768 + 8192 : 0;
    u64 total = nname + nseq + nqual;
    if (i >= total) return;
    if (i < nname) {
        int N = i < 768 ? 256 : 128;
        u8 *t = blob + (i < 768 ? i * 776 : 768ull * 776 + (i - 768) * 392);
        put_u32le(t, (u32)N); put_u32le(t + 4, 0);
        for (int k = 0; k < N; k++) { t[8 + 3 * k] = (u8)k; t[9 + 3 * k] = 1; t[10 + 3 * k] = 0; }
    } else if (i < nname + nseq) {
        u8 *t = blob + name_bytes + (i - nname) * 4;
        t[0] = t[1] = t[2] = t[3] = 3;
    } else {
        u8 *t = blob + name_bytes + (u64)nseq * 4 + (i - nname - nseq) * 293;
        put_u32le(t, 95); put_u32le(t + 4, 0);
        for (int k = 0; k < 95; k++) { t[8 + 3 * k] = (u8)k; t[9 + 3 * k] = 1; t[10 + 3 * k] = 0; }
    }
}

// =============================================================================================
// 4. range-coder pass (RangeCoder::Encode @0x422010-0x422085, flush @0x424a30): the only strictly
// serial part of the encoder.  One CTA of two warps per (block, stream):
//   warp 1 (producer) streams the triples in, 32 per step and coalesced, and turns each `tot`
//          into the 64-bit multiplier M = ceil(2^64 / tot) (floor(r / tot) = hi64(r * M), exact for
//          r < 2^32, tot <= 2^16), filling a small shared-memory ring;
//   warp 0 lane 0 (consumer) runs the recurrence.  The common path is branch free: at most two
//          renormalisation shifts per symbol, two bytes always stored, the cursor advanced by the
//          shift count.  The carry-less range adjustment (@0x422040-0x422058) fires about once
//          per 2^32 renormalisations; it is DETECTED on the fast path, and the 32-symbol group is
//          then rolled back and re-run with the exact loop.
#ifndef RC_NBUF
#define RC_NBUF 8
#endif
#ifndef RC_SLEEP
#define RC_SLEEP 0
#endif
struct RcState { u64 low; u32 range; u32 nout; };

SQ_D void rc_exact_step(RcState &s, u8 *out, u32 cum, u32 freq, u32 tot) {
    s.range /= tot;
    s.low += (u32)(cum * s.range);
    s.range *= freq;
    while (s.range < (1u << 24)) {
        if ((s.low ^ (s.low + s.range)) >> 56) s.range = (((u32)s.low) | 0xffffffu) - (u32)s.low;
        out[s.nout++] = (u8)(s.low >> 56);
        s.range <<= 8;
        s.low <<= 8;
    }
}

// Fast step.  The coder's range is kept as (r0, sh): range = r0 << sh with sh in {0, 8, 16} the
// pending renormalisation shift, so that the shift is folded into the next division:
//   floor((r0 << sh) / tot) = bits [64 - sh, 96 - sh) of r0 * M   (r0 << sh < 2^32, no truncation)
// and the serial dependency per symbol is IMAD.HI -> IMAD.WIDE -> funnel shift -> IMAD.
// `evm` tracks min over the group of (~low >> 32) & 0xffff: it is 0 iff bits 32..47 of low were all
// ones at some step, a NECESSARY condition for the carry-less adjustment (a carry out of the low 32
// bits must ripple through them to change byte 7 or byte 6) -- a conservative rollback trigger.
#define RC_FAST_STEP(rr)                                                                     \
    {                                                                                        \
        u32 t_ = __umulhi(r0, rr.x);                                                         \
        u64 P_ = (u64)r0 * rr.y + t_;                                                        \
        u32 q_ = __funnelshift_l((u32)P_, (u32)(P_ >> 32), sh);                              \
        low += (u32)((rr.z & 0xffffu) * q_);                                                 \
        r0 = q_ * rr.w;                                                                      \
        evm = min(evm, ~(u32)(low >> 32) & 0xffffu);                                         \
        op[0] = (u8)(low >> 56);                                                             \
        op[1] = (u8)(low >> 48);                                                             \
        u32 nb_ = (r0 < (1u << 24)) + (r0 < (1u << 16));                                     \
        op += nb_;                                                                           \
        sh = nb_ * 8;                                                                        \
        low <<= sh;                                                                          \
    }

__global__ void __launch_bounds__(64) k_rc(const BlockDesc *bd, BlockState *bs, const u64 *tr_seq, const u64 *tr_qual,
                                           const u64 *tr_name, u8 *obuf, int skip_name, int st) {
    __shared__ uint4 rec[RC_NBUF][32];
    __shared__ volatile u32 produced, consumed;
    int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (st == ST_NAME && skip_name) {          // rule-based IDs: no name stream
        if (threadIdx.x == 0) bs[b].olen[st] = 0;
        return;
    }
    const u64 *tr = (st == ST_QUAL ? tr_qual : tr_name) + bd[b].ebase[st];
    const u32 *tr32 = (const u32 *)tr_seq + bd[b].ebase[st];           // the base stream's triples are packed in 32 bits
    const bool s32 = st == ST_SEQ;
    u32 n = bs[b].nsym[st];
    u32 ngroups = (n + 31) / 32;
    if (threadIdx.x == 0) { produced = 0; consumed = 0; }
    __syncthreads();
    if (warp == 1) {
        // triples are fetched three groups ahead of the one being prepared
#define RC_LOAD(i) ((u64)(i) < n ? (s32 ? tr_expand32(tr32[i]) : tr[i]) : 0)
        u64 t0 = RC_LOAD(lane);
        u64 t1 = RC_LOAD(32 + lane);
        u64 t2 = RC_LOAD(64 + lane);
        for (u32 g = 0; g < ngroups; g++) {
            u64 nx = (u64)(g + 3) * 32 + lane;
            u64 t3 = RC_LOAD(nx);
            while (g >= consumed + RC_NBUF) { if (RC_SLEEP) __nanosleep(RC_SLEEP); }
            u64 t = t0;
            u32 tot = (u32)(t >> 32) & 0xffffu;
            if (tot < 2) tot = 2;
            u64 M = ~(u64)0 / tot + 1;
            // x = M lo, y = M hi, z = cum | tot << 16, w = freq
            rec[g % RC_NBUF][lane] = make_uint4((u32)M, (u32)(M >> 32), ((u32)t & 0xffffu) | (tot << 16), (u32)(t >> 16) & 0xffffu);
            t0 = t1; t1 = t2; t2 = t3;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) produced = g + 1;
        }
        return;
    }
    if (lane != 0) return;
    u8 *out = obuf + bd[b].obase[st];
    u8 *op = out;
    u8 *lim = out + bd[b].ocap[st] - 96;       // a group writes at most 65 bytes, the flush 8
    u64 low = 0;
    u32 r0 = 0xffffffffu, sh = 0;
    for (u32 g = 0; g < ngroups; g++) {
        if (op > lim) {                        // incompressible beyond the scratch bound: give up on this block
            bs[b].status = SQ_ERR_CAPACITY_K;
            consumed = ngroups + RC_NBUF;      // releases the producer
            bs[b].olen[st] = 0;
            return;
        }
        while (produced <= g) { }
        __threadfence_block();
        const uint4 *rp = rec[g % RC_NBUF];
        u32 cnt = min(32u, n - g * 32);
        RcState save{low, r0 << sh, (u32)(op - out)};
        u32 evm = 0xffffu;
        if (cnt == 32) {
#pragma unroll 8
            for (int j = 0; j < 32; j++) { uint4 rr = rp[j]; RC_FAST_STEP(rr) }
        } else {
            for (u32 j = 0; j < cnt; j++) { uint4 rr = rp[j]; RC_FAST_STEP(rr) }
        }
        if (evm == 0) {                 // a carry-less adjustment may have happened in this group: redo it exactly
            RcState s = save;
            for (u32 j = 0; j < cnt; j++) { uint4 rr = rp[j]; rc_exact_step(s, out, rr.z & 0xffffu, rr.w, rr.z >> 16); }
            low = s.low; r0 = s.range; sh = 0; op = out + s.nout;
        }
        __threadfence_block();
        consumed = g + 1;
    }
    for (int i = 0; i < 8; i++) { *op++ = (u8)(low >> 56); low <<= 8; }
    bs[b].olen[st] = (u32)(op - out);
}

// =============================================================================================
// MD5 of the three input buffers of a block (calcBlockMd5 @0x414d90): one thread per buffer.
__global__ void __launch_bounds__(32) k_md5(const BlockDesc *bd, BlockState *bs, int B) {
    int t = blockIdx.x;                       // one CTA (lane 0 of one warp) per buffer: chains must not share an issue slot
    if (t >= 3 * B || threadIdx.x != 0) return;
    int b = t / 3, which = t % 3;
    const u8 *p = which == 0 ? bd[b].ids : which == 1 ? bd[b].seq : bd[b].qual;
    u64 n = which == 0 ? bd[b].idbytes : bd[b].seqbytes;
    md5_buffer(p, n, bs[b].md5[which]);
}

// =============================================================================================
// R-Block lossy smoothing (EncapFqzComp::rblock @0x426c10): greedy runs over the
// block's concatenated quality bytes, end of data = byte count.  Serial per
// block; doubles exactly as the reference (IEEE sqrt / div, round half up).
SQ_D double rnd_half_up(double x) { return floor(x + 0.5); }

// One warp per block.  The doubles only ever see pairs of byte values, so the warp first tabulates, with the
// reference's own arithmetic, g(lo, hi) = round(sqrt(lo * hi)) and ok(lo, hi) = (r > g / lo && r > hi / g) for
// every pair below 128 (the "extend upwards" test uses (min, c), the "extend downwards" test (c, max): the same
// function of the ordered pair).  The scan itself is serial -- every run starts where the previous one closed --
// and runs on lane 0 over 4-KiB tiles staged in shared memory by the whole warp; a run is filled in the tile
// when it closes, the part of it that lies in earlier tiles by all lanes in global memory.
#define RB_TILE 4096
struct RbSmem {
    u8 g[128 * 128];             // fill value of the ordered pair (lo, hi)
    u32 ok[128 * 128 / 32];      // bit (lo * 128 + hi)
    u32 tile[RB_TILE / 4];
};
__global__ void __launch_bounds__(32) k_rblock(const BlockDesc *bd, int B, double r) {
    __shared__ RbSmem sm;
    int b = blockIdx.x, lane = threadIdx.x;
    if (b >= B) return;
    u8 *q = bd[b].qual_work;
    const u64 n = bd[b].seqbytes;
    if (n == 0) return;
    for (int w = lane; w < 128 * 128 / 32; w += 32) {
        u32 bits = 0;
        for (int k = 0; k < 32; k++) {
            int idx = w * 32 + k, lo = idx >> 7, hi = idx & 127;
            u8 gv = 0;
            if (lo > 0 && hi >= lo) {
                double g = rnd_half_up(sqrt((double)hi * lo));
                gv = (u8)g;
                if (r > g / lo && r > hi / g) bits |= 1u << k;
            }
            sm.g[idx] = gv;
        }
        sm.ok[w] = bits;
    }
    __syncwarp();
    u8 *tb = (u8 *)sm.tile;
    // Scan state (warp-uniform): the open run [start, ...) with its min / max.  The 32 lanes test 32 bytes at a time
    // against [mn, mx]; only the first byte outside the range is an event (extend the range, or close the run and
    // open the next one with that byte), after which the scan resumes right behind it with the new range.  Bytes
    // inside the range cost one ballot per 32, so the scan runs at the speed of its events, not of its bytes.
    u64 start = 0;
    int mn = q[0] & 127, mx = mn;
    for (u64 t0 = 0; t0 < n; t0 += RB_TILE) {
        const u32 tn = (u32)(n - t0 < RB_TILE ? n - t0 : RB_TILE);
        for (u32 k = lane; k < tn; k += 32) tb[k] = q[t0 + k];
        __syncwarp();
        u32 p = t0 == 0 ? 1u : 0u;                            // byte 0 opened the first run
        while (p < tn) {
            const u32 i = p + lane;
            const int ch = i < tn ? (tb[i] & 127) : mn;       // bytes past the tile never trigger
            const u32 m = __ballot_sync(FULL, ch < mn || ch > mx);
            if (m == 0) { p += 32; continue; }
            const int j = __ffs(m) - 1;
            const int c = __shfl_sync(FULL, ch, j);
            const u32 e = p + (u32)j;
            p = e + 1;
            if (c > mx) { const int idx = mn * 128 + c; if ((sm.ok[idx >> 5] >> (idx & 31)) & 1) { mx = c; continue; } }
            else { const int idx = c * 128 + mx; if ((sm.ok[idx >> 5] >> (idx & 31)) & 1) { mn = c; continue; } }
            // close [start, t0 + e): the part inside this tile in shared memory, an earlier part straight in global memory
            const u8 v = sm.g[mn * 128 + mx];
            const u32 from = start >= t0 ? (u32)(start - t0) : 0;
            __syncwarp();
            for (u32 x = from + lane; x < e; x += 32) tb[x] = v;
            if (start < t0) for (u64 x = start + lane; x < t0; x += 32) q[x] = v;
            __syncwarp();
            start = t0 + e; mn = mx = c;
        }
        if (t0 + tn >= n) {                                   // end of data closes the last run
            const u8 v = sm.g[mn * 128 + mx];
            const u32 from = start >= t0 ? (u32)(start - t0) : 0;
            __syncwarp();
            for (u32 x = from + lane; x < tn; x += 32) tb[x] = v;
            if (start < t0) for (u64 x = start + lane; x < t0; x += 32) q[x] = v;
            __syncwarp();
            start = n;
        }
        // write back what is final: everything before the run that is still open
        const u32 fin = start >= t0 + tn ? tn : (start > t0 ? (u32)(start - t0) : 0);
        __syncwarp();
        for (u32 k = lane; k < fin; k += 32) q[t0 + k] = tb[k];
        __syncwarp();
    }
}

// =============================================================================================
// Side streams: Len (compressLen @0x424120 / encode_len_short @0x4239a0), DegeTip / DegeCh /
// DegeMaxQual (@0x424dd0 / @0x425080 / @0x425310, arrays of DegeInfoProcess @0x433a10) and
// NDegeCnt / NDegePos (kModel @0x42cbe0 / @0x42ccb0, @0x42d010 / @0x42d170).  KB-scale, serial:
// one warp (lane 0) per stream, five warps per block.
struct KModel {
    SmModel<64> nb;
    SmModel<2> bits[64];
    SQ_D void init() { nb.init(); for (int i = 0; i < 64; i++) bits[i].init(); }
    SQ_D void encode(RcEnc &rc, u32 v) {
        int n = 32 - __clz(v | 0) ; if (v == 0) n = 0;
        nb.encode(rc, (u32)n);
        for (int i = 0; i < n; i++) bits[i].encode(rc, (v >> i) & 1);
    }
};
// bytes of side scratch reserved for the Len stream: 2 B per coded symbol, 3 (5 for long-read blocks) symbols per read
SQ_D u64 len_region(const BlockDesc &d) { return (d.longflag ? 10ull : 6ull) * d.nreads + 16; }
struct SideSmem {
    SmModel<256> len1, len2, len3, len4;
    SmModel<2> same, tip;
    SmModel<11> ch;
    SmModel<95> maxq;
    KModel km;
    u64 off[7];
};

__global__ void __launch_bounds__(160) k_side(const BlockDesc *bd, BlockState *bs, const u32 *rbase,
                                              const u32 *seqoff, const u8 *tip, u8 *sidebuf) {
    __shared__ SideSmem sm;
    int b = blockIdx.x, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const BlockDesc &d = bd[b];
    if (bs[b].status != 0) return;
    u32 n = d.nreads;
    u32 nch = (u32)d.seqbytes - bs[b].acgt;
    u32 nf = bs[b].side_cnt[3];
    if (threadIdx.x == 0) {
        // scratch regions: Len, Tip, Ch, MaxQ, Cnt, Pos.  Len / Tip / MaxQ get their worst case (2 B per
        // symbol); Ch gets 1 B per non-ACGT base (an 11-symbol alphabet needs < 0.5 B), Pos the rest
        // (>= 1 B per candidate base).  Overflow is detected by the coder, never assumed away.
        u64 o = 0;
        sm.off[0] = o; o += len_region(d);
        sm.off[1] = o; o += 2ull * n + 16;
        sm.off[2] = o; o += (u64)nch + 64;
        sm.off[3] = o; o += 2ull * nf + 16;
        sm.off[4] = o; o += 8ull * nf + 16;
        sm.off[5] = o;
        sm.off[6] = d.sidecap;
        if (o + 16 > d.sidecap) { bs[b].status = SQ_ERR_CAPACITY_K; for (int k = 0; k < 6; k++) sm.off[k] = 0; sm.off[6] = 0; }
    }
    __syncthreads();
    if (lane != 0) return;
    u8 *base = sidebuf + d.sidebase;
    const u32 *so = seqoff + rbase[b];
    const u8 *tp = tip + rbase[b];
    RcEnc rc;
    int ovf = 0;
    if (wid == 0) {                                             // Len
        rc.start(base + sm.off[0], base + sm.off[1]);
        sm.len1.init(); sm.len2.init(); sm.same.init();
        const bool lng = d.longflag != 0;                       // compressLen_long @0x423710: four length bytes
        if (lng) { sm.len3.init(); sm.len4.init(); }
        for (u32 r = 0; r < n; r++) {
            i32 len = d.seqlen[r];
            // the encoder compares with this+0x1060, which it never writes (ctor value 0): SURVEY 8a quirk
            if (len != 0) {
                sm.same.encode(rc, 0); sm.len1.encode(rc, len & 0xff); sm.len2.encode(rc, (len >> 8) & 0xff);
                if (lng) { sm.len3.encode(rc, (len >> 16) & 0xff); sm.len4.encode(rc, ((u32)len >> 24) & 0xff); }
            } else sm.same.encode(rc, 1);
        }
        rc.finish();
        bs[b].side_len[0] = (u32)(rc.out - (base + sm.off[0])); bs[b].side_cnt[0] = n; ovf = rc.overflow;
    } else if (wid == 1) {                                      // DegeTip: per-read has-non-ACGT flag
        rc.start(base + sm.off[1], base + sm.off[2]);
        sm.tip.init();
        for (u32 r = 0; r < n; r++) sm.tip.encode(rc, tp[r]);
        rc.finish();
        bs[b].side_len[1] = (u32)(rc.out - (base + sm.off[1])); bs[b].side_cnt[1] = n; ovf = rc.overflow;
    } else if (wid == 2) {                                      // DegeCh: the non-ACGT characters
        rc.start(base + sm.off[2], base + sm.off[3]);
        sm.ch.init();
        u32 cnt = 0;
        for (u32 r = 0; r < n; r++) {
            if (!tp[r]) continue;
            const u8 *s = d.seq + so[r];
            u32 len = (u32)d.seqlen[r];
            for (u32 i = 0; i < len; i++) { u32 v = seq_val(s[i]); if (v > 3) { sm.ch.encode(rc, v - 4); cnt++; } }
        }
        rc.finish();
        bs[b].side_len[2] = (u32)(rc.out - (base + sm.off[2])); bs[b].side_cnt[2] = cnt; ovf = rc.overflow;
    } else if (wid == 3) {                                      // DegeMaxQual
        rc.start(base + sm.off[3], base + sm.off[4]);
        sm.maxq.init();
        for (u32 r = 0; r < n; r++) {
            if (!tp[r]) continue;
            const u8 *s = d.seq + so[r], *q = d.qual + so[r];
            u32 len = (u32)d.seqlen[r];
            u8 mq = 0;
            for (u32 i = 0; i < len; i++) if (seq_val(s[i]) > 3 && (int)(int8_t)q[i] > (int)mq) mq = q[i];
            sm.maxq.encode(rc, (u32)(u16)(mq - 33));
        }
        rc.finish();
        bs[b].side_len[3] = (u32)(rc.out - (base + sm.off[3])); ovf = rc.overflow;   // side_cnt[3] = nf already
    } else if (wid == 4) {                                      // NDegeCnt then NDegePos (shared kModel)
        sm.km.init();
        rc.start(base + sm.off[4], base + sm.off[5]);
        u32 npos = 0;
        for (u32 r = 0; r < n; r++) {
            if (!tp[r]) continue;
            const u8 *s = d.seq + so[r], *q = d.qual + so[r];
            u32 len = (u32)d.seqlen[r];
            u8 mq = 0;
            for (u32 i = 0; i < len; i++) if (seq_val(s[i]) > 3 && (int)(int8_t)q[i] > (int)mq) mq = q[i];
            u32 c = 0;
            for (u32 i = 0; i < len; i++) if (!((int)mq < (int)(int8_t)q[i]) && seq_val(s[i]) <= 3) c++;
            sm.km.encode(rc, c);
            npos += c;
        }
        rc.finish();
        bs[b].side_len[4] = (u32)(rc.out - (base + sm.off[4])); bs[b].side_cnt[4] = nf; ovf = rc.overflow;
        rc.start(base + sm.off[5], base + sm.off[6]);
        for (u32 r = 0; r < n; r++) {
            if (!tp[r]) continue;
            const u8 *s = d.seq + so[r], *q = d.qual + so[r];
            u32 len = (u32)d.seqlen[r];
            u8 mq = 0;
            for (u32 i = 0; i < len; i++) if (seq_val(s[i]) > 3 && (int)(int8_t)q[i] > (int)mq) mq = q[i];
            u32 gap = 0;
            for (u32 i = 0; i < len; i++) {
                if ((int)mq < (int)(int8_t)q[i]) continue;
                if (seq_val(s[i]) > 3) gap++;
                else { sm.km.encode(rc, gap); gap = 0; }
            }
        }
        rc.finish();
        bs[b].side_len[5] = (u32)(rc.out - (base + sm.off[5])); bs[b].side_cnt[5] = npos; ovf |= rc.overflow;
    }
    if (ovf) bs[b].status = SQ_ERR_CAPACITY_K;
}

// =============================================================================================
// 5. pack: block record "tag 1 | size(4) | count | Len | ID | Qual | Dege* | Seq" (SURVEY App. A;
// doFqzEncode @0x42d2d0).  k_layout sizes the records (and, for the compact wave buffer, assigns
// offsets by a scan over blocks); k_pack copies the pieces.
struct PackCfg { int md5, lossy, id_rule, ref_mode; };
__constant__ int c_side_tag[6] = {4, 23, 14, 24, 25, 26};

SQ_D u32 id_payload_len(const BlockDesc &d, const BlockState &s, PackCfg cfg) {
    u32 x = cfg.md5 ? 16 : 0;
    if (cfg.id_rule) x += 2 + (d.nreads ? (u32)d.idlen[0] : 0);      // IDProcess::encodeIDS @0x430040
    else x += s.olen[ST_NAME];
    return x;
}
// Framing doAlignEncode @0x42d4c0 adds when a reference index is loaded but no read mapped (ref_mode
// 1 = SE, 2 = PE): count records (tag | 0x84 | u32le) and alignment streams with no symbols, i.e. the
// 8 flush bytes of an untouched range coder.  tested_reads = flags pushed by doAlign @0x411910 /
// @0x413580 for the head (int)(nreads * 0.05) reads before it gives up (tag 27).
SQ_D u32 tested_reads(u32 nreads, int pe) {
    u32 t = (u32)(int)((double)nreads * 0.05);
    u32 c = pe ? 2 * (t / 2 + 1) : t + 1;
    return c < nreads ? c : nreads;
}
SQ_D u32 put_count(u8 *p, int tag, u32 v) { put_tag(p, tag); p[1] = 0x84; put_u32le(p + 2, v); return 6; }
SQ_D u32 put_empty_rc(u8 *p, int tag) {
    put_tag(p, tag); put_size4(p + 1, 8);
    for (int i = 0; i < 8; i++) p[5 + i] = 0;
    return 13;
}
SQ_D u32 record_len(const BlockDesc &d, const BlockState &s, PackCfg cfg) {
    u32 x = 5 + 6;                                                    // tag 1 + size4, count record
    if (cfg.ref_mode) x += 6 + 13 + 4 * 6 + (cfg.ref_mode == 2 ? 6 + 6 + 13 : 0) + 5 * 13;
    x += 5 + s.side_len[0];                                           // Len
    x += 5 + id_payload_len(d, s, cfg);
    x += 5 + ((cfg.md5 && !cfg.lossy) ? 16 : 0) + s.olen[ST_QUAL];
    for (int k = 1; k < 6; k++) if (s.side_cnt[k]) x += 5 + 4 + s.side_len[k];
    x += 5 + (cfg.md5 ? 16 : 0) + s.olen[ST_SEQ];
    return x;
}

// single CTA: lengths and (when compact != 0) consecutive offsets into recbuf
__global__ void __launch_bounds__(1024) k_layout(BlockDesc *bd, BlockState *bs, int B, PackCfg cfg, u8 *recbuf,
                                                 u64 recbuf_cap, u64 *recoff, int compact) {
    __shared__ u64 carry;
    __shared__ u64 wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int b0 = 0; b0 < B; b0 += 1024) {
        int b = b0 + threadIdx.x;
        u64 len = 0;
        if (b < B) { len = record_len(bd[b], bs[b], cfg); bs[b].reclen = (u32)len; }
        u64 inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u64 t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            u64 t = wsum[lane], ti = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u64 u = __shfl_up_sync(FULL, ti, o); if (lane >= o) ti += u; }
            wsum[lane] = ti - t;
        }
        __syncthreads();
        u64 off = carry + wsum[wid] + inc - len;
        if (b < B) {
            recoff[b] = off;
            if (compact) {
                bd[b].rec = recbuf + off;
                bd[b].reccap = off + len <= recbuf_cap ? len : 0;
            }
            if (len > bd[b].reccap && bs[b].status == 0) bs[b].status = SQ_ERR_CAPACITY_K;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = off + len;
        __syncthreads();
    }
    if (threadIdx.x == 0) recoff[B] = carry;
}

SQ_D void copy_bytes(u8 *dst, const u8 *src, u32 n, u32 tid, u32 nthr) {
    for (u32 i = tid; i < n; i += nthr) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) k_pack(const BlockDesc *bd, const BlockState *bs, PackCfg cfg,
                                              const u8 *obuf, const u8 *sidebuf) {
    int b = blockIdx.y;
    const BlockDesc &d = bd[b];
    const BlockState &s = bs[b];
    if (s.status != 0) return;
    u8 *rec = d.rec;
    u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    u32 p = 0;
    bool w0 = (tid == 0);
    // tag 1 + size of everything after the size slot
    if (w0) { put_tag(rec, 1); put_size4(rec + 1, s.reclen - 5); }
    p = 5;
    if (w0) { rec[p] = 0x81; rec[p + 1] = 0x84; put_u32le(rec + p + 2, d.nreads); }   // compressCount @0x422a00
    p += 6;
    // Len (tag 4): pure range-coder bytes
    if (w0) { put_tag(rec + p, 4); put_size4(rec + p + 1, s.side_len[0]); }
    p += 5;
    copy_bytes(rec + p, sidebuf + d.sidebase, s.side_len[0], tid, nthr);   // Len region is at offset 0
    p += s.side_len[0];
    if (cfg.ref_mode) {                                           // @0x42d547-0x42d56f: tag 27, compressOrder (tag 8)
        if (w0) { put_count(rec + p, 27, tested_reads(d.nreads, cfg.ref_mode == 2)); put_empty_rc(rec + p + 6, 8); }
        p += 6 + 13;
    }
    // ID (tag 5)
    u32 idl = id_payload_len(d, s, cfg);
    if (w0) { put_tag(rec + p, 5); put_size4(rec + p + 1, idl); }
    p += 5;
    if (cfg.md5) { copy_bytes(rec + p, s.md5[0], 16, tid, nthr); p += 16; }
    if (cfg.id_rule) {
        u32 l = d.nreads ? (u32)d.idlen[0] : 0;
        if (w0) { rec[p] = (u8)l; rec[p + 1] = (u8)(l >> 8); }
        copy_bytes(rec + p + 2, d.ids, l, tid, nthr);
        p += 2 + l;
    } else {
        copy_bytes(rec + p, obuf + d.obase[ST_NAME], s.olen[ST_NAME], tid, nthr);
        p += s.olen[ST_NAME];
    }
    // Qual (tag 7)
    u32 qmd5 = (cfg.md5 && !cfg.lossy) ? 16 : 0;
    if (w0) { put_tag(rec + p, 7); put_size4(rec + p + 1, qmd5 + s.olen[ST_QUAL]); }
    p += 5;
    if (qmd5) { copy_bytes(rec + p, s.md5[2], 16, tid, nthr); p += 16; }
    copy_bytes(rec + p, obuf + d.obase[ST_QUAL], s.olen[ST_QUAL], tid, nthr);
    p += s.olen[ST_QUAL];
    if (cfg.ref_mode) {                                           // @0x42d5bd-0x42d6c7 (PE part @0x42d830)
        if (w0) {
            u8 *q = rec + p;
            for (int t = 17; t <= 20; t++) q += put_count(q, t, 0);
            if (cfg.ref_mode == 2) { q += put_count(q, 21, 9); q += put_count(q, 22, 0); q += put_empty_rc(q, 9); }
            const int tags[5] = {11, 15, 10, 12, 13};             // Pos, Mis, Rev, CigaL, CigaV
            for (int k = 0; k < 5; k++) q += put_empty_rc(q, tags[k]);
        }
        p += 4 * 6 + (cfg.ref_mode == 2 ? 6 + 6 + 13 : 0) + 5 * 13;
    }
    // Dege streams (tags 23, 14, 24, 25, 26): u32le count + range-coder bytes, omitted when empty
    {
        u64 o = len_region(d);                                     // region offsets as laid out by k_side
        u32 nch = (u32)d.seqbytes - s.acgt, nf = s.side_cnt[3];
        u64 offs[6] = {0, o, o + 2ull * d.nreads + 16, 0, 0, 0};
        offs[3] = offs[2] + (u64)nch + 64;
        offs[4] = offs[3] + 2ull * nf + 16;
        offs[5] = offs[4] + 8ull * nf + 16;
        for (int k = 1; k < 6; k++) {
            if (!s.side_cnt[k]) continue;
            if (w0) { put_tag(rec + p, c_side_tag[k]); put_size4(rec + p + 1, 4 + s.side_len[k]); put_u32le(rec + p + 5, s.side_cnt[k]); }
            p += 9;
            copy_bytes(rec + p, sidebuf + d.sidebase + offs[k], s.side_len[k], tid, nthr);
            p += s.side_len[k];
        }
    }
    // Seq (tag 6)
    if (w0) { put_tag(rec + p, 6); put_size4(rec + p + 1, (cfg.md5 ? 16 : 0) + s.olen[ST_SEQ]); }
    p += 5;
    if (cfg.md5) { copy_bytes(rec + p, s.md5[1], 16, tid, nthr); p += 16; }
    copy_bytes(rec + p, obuf + d.obase[ST_SEQ], s.olen[ST_SEQ], tid, nthr);
}

// =============================================================================================
// host-side launchers (called from sq_api.cu)
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)

static inline u32 cdiv(u64 a, u64 b) { return (u32)((a + b - 1) / b); }

cudaError_t sqk_prep(const EncWave &w, cudaStream_t st, u64 *launches) {
    if (w.R == 0) return cudaSuccess;
    k_read_prep<<<cdiv(w.R, 256), 256, 0, st>>>(w.bd, w.bs, w.rbase, w.B, w.R, w.seqoff, w.idoff);
    ScanArgs a{};
    a.arr[0] = w.seqoff; a.arr[1] = w.idoff;
    k_excl_scan<<<dim3(w.B, 2), 1024, 0, st>>>(a, w.rbase, w.totals + 3 * w.B, w.B);
    k_check_totals<<<cdiv(w.B, 128), 128, 0, st>>>(w.bd, w.bs, w.totals + 3 * w.B, w.totals + 4 * w.B, w.B);
    *launches += 3;
    return cudaGetLastError();
}

cudaError_t sqk_rblock(const EncWave &w, double r, cudaStream_t st, u64 *launches) {
    k_rblock<<<w.B, 32, 0, st>>>(w.bd, w.B, r);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t sqk_count(const EncWave &w, cudaStream_t st, u64 *launches) {
    if (w.R == 0) {
        k_set_counts<<<cdiv(w.B, 128), 128, 0, st>>>(w.bs, w.totals, w.B, !w.id_rule);
        *launches += 1;
        return cudaGetLastError();
    }
    u32 gw = cdiv((u64)w.R * 32, 256);
    k_seq_pass<false><<<gw, 256, 0, st>>>(w.bd, w.bs, w.rbase, w.B, w.R, w.seqoff, w.acgtoff, w.tip, nullptr, w.seqmask);
    k_qual_pass<false><<<gw, 256, 0, st>>>(w.bd, w.bs, w.rbase, w.B, w.R, w.seqoff, w.qlen2, w.qoff, nullptr, w.qlevel);
    if (!w.id_rule)
        k_name_pass<false><<<cdiv(w.R, 128), 128, 0, st>>>(w.bd, w.bs, w.rbase, w.B, w.R, w.idoff, w.np, w.ns, w.noff, nullptr);
    ScanArgs a{};
    a.arr[0] = w.acgtoff; a.arr[1] = w.qoff; a.arr[2] = w.noff;
    k_excl_scan<<<dim3(w.B, w.id_rule ? 2 : 3), 1024, 0, st>>>(a, w.rbase, w.totals, w.B);
    k_set_counts<<<cdiv(w.B, 128), 128, 0, st>>>(w.bs, w.totals, w.B, !w.id_rule);
    *launches += w.id_rule ? 4 : 5;
    return cudaGetLastError();
}

cudaError_t sqk_emit(const EncWave &w, int stream, u64 *elem, cudaStream_t st, u64 *launches) {
    if (w.R == 0) return cudaSuccess;
    u32 gw = cdiv((u64)w.R * 32, 256);
    if (stream == ST_SEQ)
        k_seq_pass<true><<<gw, 256, 0, st>>>(w.bd, w.bs, w.rbase, w.B, w.R, w.seqoff, w.acgtoff, w.tip, elem, w.seqmask);
    else if (stream == ST_QUAL)
        k_qual_pass<true><<<gw, 256, 0, st>>>(w.bd, w.bs, w.rbase, w.B, w.R, w.seqoff, w.qlen2, w.qoff, elem, w.qlevel);
    else
        k_name_pass<true><<<cdiv(w.R, 128), 128, 0, st>>>(w.bd, w.bs, w.rbase, w.B, w.R, w.idoff, w.np, w.ns, w.noff, elem);
    *launches += 1;
    return cudaGetLastError();
}

template <int BITS>
static size_t sort_scatter_smem() { return (size_t)SQ_TILE * 8 + (size_t)(1 << BITS) * 8 + (size_t)SQ_SORT_WARPS * (1 << BITS) * 2; }

template <int BITS>
static cudaError_t sort_pass(const EncWave &w, int stream, const u64 *src, u64 *dst, int shift, u32 maxtiles,
                             cudaStream_t st, u64 *launches) {
    constexpr int D = 1 << BITS;
    dim3 grid(maxtiles, w.B);
    const u64 entries = (u64)D * maxtiles;
    const u32 nchunks = cdiv(entries, SQ_SCAN_CHUNK);
    k_sort_hist<BITS><<<grid, SQ_SORT_THREADS, 0, st>>>(src, w.hist, w.bd, w.bs, stream, shift, maxtiles);
    k_scan_partial<<<dim3(nchunks, w.B), 256, 0, st>>>(w.hist, entries, nchunks, w.scanpart);
    k_scan_mid<<<w.B, 1024, 0, st>>>(w.scanpart, nchunks);
    k_scan_final<<<dim3(nchunks, w.B), 256, 0, st>>>(w.hist, entries, nchunks, w.scanpart);
    k_sort_scatter<BITS><<<grid, SQ_SORT_THREADS, sort_scatter_smem<BITS>(), st>>>(src, dst, w.hist, w.bd, w.bs, stream, shift, maxtiles);
    *launches += 5;
    return cudaGetLastError();
}

// sorts elemA (in place result pointer returned): keybits context bits, digits of SQ_SORT_BITS
cudaError_t sqk_sort(const EncWave &w, int stream, int keybits, u64 *a, u64 *b, u64 **sorted, cudaStream_t st, u64 *launches) {
    u32 maxtiles = cdiv(w.maxbound[stream], SQ_TILE);
    u64 *src = a, *dst = b;
    if (maxtiles == 0) { *sorted = a; return cudaSuccess; }
    // 17..20 key bits (the default order-10 base context): two passes of 10 bits instead of three of 8
    const bool wide = keybits > 2 * SQ_SORT_BITS && keybits <= 2 * SQ_SORT_BITS_WIDE;
    const int step = wide ? SQ_SORT_BITS_WIDE : SQ_SORT_BITS;
    for (int shift = 0; shift < keybits; shift += step) {
        if (wide) CK(sort_pass<SQ_SORT_BITS_WIDE>(w, stream, src, dst, shift, maxtiles, st, launches));
        else CK(sort_pass<SQ_SORT_BITS>(w, stream, src, dst, shift, maxtiles, st, launches));
        u64 *t = src; src = dst; dst = t;
    }
    *sorted = src;
    return cudaSuccess;
}

cudaError_t sqk_replay(const EncWave &w, int stream, const u64 *sorted, void *triples_, cudaStream_t st, u64 *launches) {
    u64 *triples = (u64 *)triples_;
    if (w.maxbound[stream] == 0) return cudaSuccess;
    CK(cudaMemsetAsync(w.nseg, 0, 4 * sizeof(u32), st));
    dim3 g(cdiv(w.maxbound[stream], 256), w.B);
    k_seglist<<<g, 256, 0, st>>>(sorted, w.bd, w.bs, stream, w.seglist, w.nseg, w.segcap);
    *launches += 1;
    int nsm = w.nsm;
    if (stream == ST_SEQ) {
        k_replay_seq<<<nsm * 16, 128, 0, st>>>(sorted, w.bd, w.bs, w.seglist, w.nseg, w.nseg + 1, (u32 *)triples_, w.snap_seq, w.train_seq);
        *launches += 1;
    } else if (stream == ST_QUAL) {
        k_replay_batch<95><<<nsm * 8, 128, 0, st>>>(sorted, w.bd, w.bs, stream, w.seglist, w.nseg, w.nseg + 1, triples,
                                                    0, 1u << 24, w.snap_qual, 293, w.train_qual);
        *launches += 1;
    } else {
        k_replay_batch<256><<<nsm * 8, 128, 0, st>>>(sorted, w.bd, w.bs, stream, w.seglist, w.nseg, w.nseg + 1,
                                                     triples, 0, NK_MIDDLE, w.snap_name, 776, w.train_name);
        k_replay_batch<128><<<nsm * 8, 128, 0, st>>>(sorted, w.bd, w.bs, stream, w.seglist, w.nseg, w.nseg + 2,
                                                     triples, NK_MIDDLE, NK_TOTAL,
                                                     w.snap_name ? w.snap_name + 3ull * 256 * 776 : nullptr, 392,
                                                     w.train_name ? w.train_name + 3ull * 256 * 776 : nullptr);
        *launches += 2;
    }
    return cudaGetLastError();
}

cudaError_t sqk_snap_init(u8 *blob, u64 name_bytes, u32 nseq, u32 nqual, cudaStream_t st, u64 *launches) {
    u64 total = (name_bytes ? 768 + 8192 : 0) + (u64)nseq + nqual;
    k_snap_init<<<cdiv(total, 256), 256, 0, st>>>(blob, name_bytes, nseq, nqual);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t sqk_rc(const EncWave &w, int stream, cudaStream_t st, u64 *launches) {
    k_rc<<<w.B, 64, SQ_CHAIN_SMEM_RC, st>>>(w.bd, w.bs, w.triples[ST_SEQ], w.triples[ST_QUAL], w.triples[ST_NAME], w.obuf, w.id_rule, stream);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t sqk_md5(const EncWave &w, cudaStream_t st, u64 *launches) {
    k_md5<<<3 * w.B, 32, SQ_CHAIN_SMEM_SMALL, st>>>(w.bd, w.bs, w.B);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t sqk_side(const EncWave &w, cudaStream_t st, u64 *launches) {
    k_side<<<w.B, 160, SQ_CHAIN_SMEM_SMALL, st>>>(w.bd, w.bs, w.rbase, w.seqoff, w.tip, w.sidebuf);
    *launches += 1;
    return cudaGetLastError();
}

cudaError_t sqk_pack(const EncWave &w, int compact, cudaStream_t st, u64 *launches) {
    PackCfg cfg{w.md5, w.lossy, w.id_rule, w.ref_mode};
    k_layout<<<1, 1024, 0, st>>>(w.bd, w.bs, w.B, cfg, w.recbuf, w.recbuf_cap, w.recoff, compact);
    k_pack<<<dim3(32, w.B), 256, 0, st>>>(w.bd, w.bs, cfg, w.obuf, w.sidebuf);
    *launches += 2;
    return cudaGetLastError();
}

// Chain kernels are launched on side streams while the main stream keeps every SM busy, and the
// block scheduler then drops their CTAs onto whichever SM frees a slot first: without a limit two
// dozen serial chains end up sharing one SM's four issue ports (measured: 6.6x slower).  Asking for
// shared memory they do not use caps the chains per SM (227 KB / 72 KB = 3, / 36 KB = 6).
cudaError_t sqk_enc_attrs() {
    CK(cudaFuncSetAttribute(k_rc, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_CHAIN_SMEM_RC));
    CK(cudaFuncSetAttribute(k_md5, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_CHAIN_SMEM_SMALL));
    CK(cudaFuncSetAttribute(k_side, cudaFuncAttributeMaxDynamicSharedMemorySize, SQ_CHAIN_SMEM_SMALL));
    CK(cudaFuncSetAttribute(k_sort_scatter<SQ_SORT_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_scatter_smem<SQ_SORT_BITS>()));
    CK(cudaFuncSetAttribute(k_sort_scatter<SQ_SORT_BITS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_scatter_smem<SQ_SORT_BITS_WIDE>()));
    return cudaSuccess;
}
