// sq_api.cu -- the C ABI declared in include/seqarc_b200.h: context, device workspace, wave
// planning and kernel orchestration.  Stands where one EncapFqzComp + SeqArcMemBuf pair stands in
// the reference (ctor @0x42f2a0, doFqzEncode @0x42d2d0, doDecode @0x42eb60), for a batch of blocks.
// There is no CPU path in this file: every coder runs in the kernels of sq_encode.cu / sq_decode.cu.
#include "../../include/seqarc_b200.h"
#include "sq_kernels.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_error;

struct Arena {
    u8 *base = nullptr;
    u64 cap = 0, used = 0;
    void reset() { used = 0; }
    void *take(u64 bytes) {
        u64 a = (used + 255) & ~(u64)255;
        if (a + bytes > cap) return nullptr;
        used = a + bytes;
        return base + a;
    }
};

struct StageTimer {
    struct Span { const char *name; cudaEvent_t e0, e1; };
    std::vector<Span> spans;           // pool; first `n` are live for the current wave
    std::vector<std::pair<const char *, float>> acc;
    size_t n = 0;
    void begin_wave() { n = 0; }
    // open a span on `st`; returns its index
    size_t open(const char *name, cudaStream_t st) {
        if (n == spans.size()) { Span sp{}; cudaEventCreate(&sp.e0); cudaEventCreate(&sp.e1); spans.push_back(sp); }
        spans[n].name = name;
        cudaEventRecord(spans[n].e0, st);
        return n++;
    }
    void close(size_t i, cudaStream_t st) { cudaEventRecord(spans[i].e1, st); }
    void collect() {       // after the streams were synchronised
        for (size_t i = 0; i < n; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, spans[i].e0, spans[i].e1);
            bool found = false;
            for (auto &a : acc) if (a.first == spans[i].name) { a.second += ms; found = true; break; }
            if (!found) acc.push_back({spans[i].name, ms});
        }
    }
    void destroy() { for (auto &sp : spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); } spans.clear(); }
};

}  // namespace

struct sq_ctx {
    sq_params p;
    int nsm = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t aux[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // md5, side, rc x 3 run beside the main stream
    cudaEvent_t ev[8] = {};
    Arena arena;
    std::string err;
    // model snapshot
    u8 *d_model = nullptr;
    u8 *d_ref = nullptr;              // aligned mode: the concatenated reference sequence (sq_set_reference)
    u64 ref_len = 0;
    u64 model_len = 0;
    // measurement
    StageTimer timer;
    u64 launches = 0;
    int last_wave_blocks = 0;
    // last encoded wave (debug fetch)
    EncWave last{};
    std::vector<BlockDesc> last_bd;
    bool last_valid = false;
    // pinned scratch for small D2H reads
    u8 *h_pin = nullptr;
    u64 h_pin_cap = 0;

    u32 NS() const { return 7 + (u32)p.slevel; }
    u32 nseq() const { return 1u << (2 * NS()); }
    u32 nqual() const { return p.qlevel > 2 ? 0x100000u : 0x10000u; }
    u64 snap_name_bytes() const { return p.id_rule ? 0 : 3ull * 256 * 776 + 8192ull * 392; }
};

#define CUDA_OK(c, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    (c)->err = std::string(#x) + ": " + cudaGetErrorString(e_); return SQ_ERR_CUDA; } } while (0)

static u64 al16(u64 x) { return (x + 15) & ~(u64)15; }

static int ensure_pin(sq_ctx *c, u64 bytes) {
    if (bytes <= c->h_pin_cap) return SQ_OK;
    if (c->h_pin) cudaFreeHost(c->h_pin);
    c->h_pin = nullptr; c->h_pin_cap = 0;
    CUDA_OK(c, cudaMallocHost((void **)&c->h_pin, bytes));
    c->h_pin_cap = bytes;
    return SQ_OK;
}

extern "C" int sq_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

extern "C" const char *sq_version(void) { return "seqarc-b200 0.1 (SeqArc v1.6 stream format)"; }

extern "C" const char *sq_last_error(const sq_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

extern "C" int sq_create(const sq_params *p, sq_ctx **out) {
    if (!p || !out) { g_create_error = "null argument"; return SQ_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        return SQ_ERR_NO_DEVICE;
    }
    if (p->device < 0 || p->device >= ndev) { g_create_error = "device ordinal out of range"; return SQ_ERR_ARG; }
    if (p->slevel < 0 || p->slevel > 5) { g_create_error = "slevel must be 0..5 (order 7..12)"; return SQ_ERR_ARG; }
    if (p->qlevel < 1 || p->qlevel > 3) { g_create_error = "qlevel must be 1..3"; return SQ_ERR_ARG; }
    if (p->lossy && !(p->lossy_r > 1.0)) { g_create_error = "lossy_r must be > 1"; return SQ_ERR_ARG; }
    if (p->ref_mode < 0 || p->ref_mode > 2) { g_create_error = "ref_mode must be 0, 1 (SE) or 2 (PE)"; return SQ_ERR_ARG; }
    sq_ctx *c = new sq_ctx();
    c->p = *p;
    auto fail = [&](const std::string &m, int code) { g_create_error = m; sq_destroy(c); return code; };
    if ((e = cudaSetDevice(p->device)) != cudaSuccess) return fail(cudaGetErrorString(e), SQ_ERR_CUDA);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, p->device)) != cudaSuccess) return fail(cudaGetErrorString(e), SQ_ERR_CUDA);
    c->nsm = prop.multiProcessorCount;
    if (p->stream) c->stream = (cudaStream_t)p->stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(cudaGetErrorString(e), SQ_ERR_CUDA);
        c->own_stream = true;
    }
    for (int i = 0; i < 5; i++)
        if ((e = cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking)) != cudaSuccess) return fail(cudaGetErrorString(e), SQ_ERR_CUDA);
    for (int i = 0; i < 8; i++)
        if ((e = cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming)) != cudaSuccess) return fail(cudaGetErrorString(e), SQ_ERR_CUDA);
    u64 want = p->workspace_bytes;
    if (want == 0) {
        size_t fr = 0, tot = 0;
        if ((e = cudaMemGetInfo(&fr, &tot)) != cudaSuccess) return fail(cudaGetErrorString(e), SQ_ERR_CUDA);
        want = (u64)(fr * 0.8);
    }
    if ((e = cudaMalloc((void **)&c->arena.base, want)) != cudaSuccess)
        return fail(std::string("workspace cudaMalloc: ") + cudaGetErrorString(e), SQ_ERR_NOMEM);
    c->arena.cap = want;
    if ((e = sqk_enc_attrs()) != cudaSuccess) return fail(std::string("kernel attributes: ") + cudaGetErrorString(e), SQ_ERR_CUDA);
    *out = c;
    return SQ_OK;
}

extern "C" void sq_destroy(sq_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->p.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int i = 0; i < 5; i++) if (c->aux[i]) { cudaStreamSynchronize(c->aux[i]); cudaStreamDestroy(c->aux[i]); }
    for (int i = 0; i < 8; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    c->timer.destroy();
    if (c->arena.base) cudaFree(c->arena.base);
    if (c->d_model) cudaFree(c->d_model);
    if (c->d_ref) cudaFree(c->d_ref);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" uint64_t sq_model_size(const sq_ctx *c) {
    if (!c) return 0;
    return c->snap_name_bytes() + (u64)c->nseq() * 4 + (u64)c->nqual() * 293;
}

extern "C" int sq_load_model(sq_ctx *c, const uint8_t *blob, uint64_t len) {
    if (!c || !blob) return SQ_ERR_ARG;
    if (len != sq_model_size(c)) { c->err = "model snapshot has the wrong size for this slevel/qlevel/id_rule"; return SQ_ERR_ARG; }
    CUDA_OK(c, cudaSetDevice(c->p.device));
    if (c->d_model) { cudaFree(c->d_model); c->d_model = nullptr; }
    CUDA_OK(c, cudaMalloc((void **)&c->d_model, len));
    CUDA_OK(c, cudaMemcpy(c->d_model, blob, len, cudaMemcpyHostToDevice));
    c->model_len = len;
    // base contexts: the training rule (@0x421fb0: halve above 253, then one increment) bounds every context's
    // total by 254; anything else is not a SaveModelToMem blob
    int bad = 0, *d_flag = nullptr;
    CUDA_OK(c, cudaMalloc((void **)&d_flag, sizeof(int)));
    cudaError_t e = sqk_check_seqsnap(c->d_model + c->snap_name_bytes(), c->nseq(), d_flag, &bad, c->stream);
    cudaFree(d_flag);
    CUDA_OK(c, e);
    if (bad) {
        cudaFree(c->d_model); c->d_model = nullptr; c->model_len = 0;
        c->err = "model snapshot: a base context sums above 254 (corrupt snapshot)";
        return SQ_ERR_FORMAT;
    }
    return SQ_OK;
}

extern "C" int sq_last_timings(const sq_ctx *c, const char **names, float *ms, int cap) {
    if (!c) return 0;
    int n = 0;
    for (auto &a : c->timer.acc) { if (n >= cap) break; names[n] = a.first; ms[n] = a.second; n++; }
    return n;
}
extern "C" uint64_t sq_last_launches(const sq_ctx *c) { return c ? c->launches : 0; }
extern "C" int sq_last_wave_blocks(const sq_ctx *c) { return c ? c->last_wave_blocks : 0; }

// ---------------------------------------------------------------------------------------------
// encode
namespace {

struct BlkCost {
    u64 inputs, perread, eb[ST_COUNT], obuf, side, rec;
};

const u32 NCTX_NAME = NK_TOTAL;

BlkCost block_cost(const sq_ctx *c, const sq_block_in &b, bool host) {
    BlkCost k{};
    u64 n = b.nreads;
    if (host) k.inputs = al16(b.idbytes) + 2 * al16(b.seqbytes) + al16(2 * n) + al16(4 * n) + 5 * 256;
    if (c->p.lossy) k.inputs += al16(b.seqbytes) + 256;
    k.perread = n * 27 + 9 * 256;
    k.eb[ST_SEQ] = b.seqbytes;
    k.eb[ST_QUAL] = b.seqbytes + n;
    k.eb[ST_NAME] = c->p.id_rule ? 0 : b.idbytes + 3 * n;
    k.obuf = 0;
    for (int s = 0; s < ST_COUNT; s++) k.obuf += al16(k.eb[s] + k.eb[s] / 2 + 256);
    k.side = al16(28 * n + b.seqbytes + 4096);
    k.rec = al16(256 + b.idbytes + 2 * b.seqbytes + 8 * n);
    return k;
}

struct WaveAcc {
    u64 lin = 0;                  // linear part
    u64 sumeb[ST_COUNT] = {0, 0, 0};
    u64 maxeb[ST_COUNT] = {0, 0, 0};
    u64 segsum[ST_COUNT] = {0, 0, 0};
    int B = 0;
    u64 total(const sq_ctx *c) const {
        u64 t = lin + 64 * 1024;
        u64 me = 0, ms = 0, mt = 0;
        for (int s = 0; s < ST_COUNT; s++) {
            t += (s == ST_SEQ ? 4 : 8) * sumeb[s] + 256;
            me = std::max(me, sumeb[s]);
            ms = std::max(ms, segsum[s]);
            mt = std::max(mt, (maxeb[s] + SQ_TILE - 1) / SQ_TILE);
        }
        t += 16 * me + 512;                                   // elemA + elemB
        t += 8 * ms + 256;                                    // seglist
        t += (u64)B * (1u << SQ_SORT_BITS_WIDE) * mt * 4 + 256;    // hist
        t += (u64)B * (((u64)(1u << SQ_SORT_BITS_WIDE) * mt + 4095) / 4096 + 1) * 4 + 512;    // its scan's chunk sums
        t += (u64)B * (sizeof(BlockDesc) + sizeof(BlockState) + 6 * 4 + 8 + 4) + 2048;
        (void)c;
        return t;
    }
    void add(const sq_ctx *c, const BlkCost &k) {
        lin += k.inputs + k.perread + k.obuf + k.side + k.rec;
        u64 nctx[ST_COUNT] = {c->nseq(), c->nqual(), NCTX_NAME};
        for (int s = 0; s < ST_COUNT; s++) {
            sumeb[s] += k.eb[s];
            maxeb[s] = std::max(maxeb[s], k.eb[s]);
            segsum[s] += std::min<u64>(k.eb[s], nctx[s]);
        }
        B++;
    }
};

template <class T> T *take(Arena &a, u64 count) { return (T *)a.take(count * sizeof(T)); }

int encode_wave(sq_ctx *c, const sq_block_in *blocks, int n, sq_block_out *out, bool host, u8 *train_blob = nullptr) {
    cudaStream_t st = c->stream;
    Arena &A = c->arena;
    A.reset();
    EncWave w{};
    w.B = n; w.nsm = c->nsm;
    w.md5 = c->p.md5; w.lossy = c->p.lossy; w.id_rule = c->p.id_rule; w.qlevel = c->p.qlevel;
    w.ref_mode = c->p.ref_mode;
    w.seqmask = c->nseq() - 1;
    if (train_blob) {            // encodeForModel @0x4228e0: models start from the constructor state
        w.train_name = c->p.id_rule ? nullptr : train_blob;
        w.train_seq = train_blob + c->snap_name_bytes();
        w.train_qual = w.train_seq + (u64)c->nseq() * 4;
    } else if (c->p.usemodel) {
        if (!c->d_model) { c->err = "usemodel = 1 but no snapshot loaded (sq_load_model)"; return SQ_ERR_ARG; }
        w.snap_name = c->p.id_rule ? nullptr : c->d_model;
        w.snap_seq = c->d_model + c->snap_name_bytes();
        w.snap_qual = w.snap_seq + (u64)c->nseq() * 4;
    }
    std::vector<BlockDesc> bd(n);
    std::vector<u32> rbase(n + 1);
    u64 R = 0;
    u64 sum[ST_COUNT] = {0, 0, 0}, osum = 0, sidesum = 0, recsum = 0, segsum[ST_COUNT] = {0, 0, 0};
    u64 nctx[ST_COUNT] = {c->nseq(), c->nqual(), NCTX_NAME};
    for (int i = 0; i < n; i++) {
        const sq_block_in &b = blocks[i];
        BlkCost k = block_cost(c, b, host);
        BlockDesc &d = bd[i];
        memset(&d, 0, sizeof d);
        d.nreads = b.nreads; d.idbytes = b.idbytes; d.seqbytes = b.seqbytes;
        d.longflag = b.longflag ? 1 : 0;
        d.rbase = (u32)R;
        rbase[i] = (u32)R;
        R += b.nreads;
        for (int s = 0; s < ST_COUNT; s++) {
            d.ebase[s] = sum[s]; d.ebound[s] = k.eb[s];
            sum[s] += k.eb[s];
            w.maxbound[s] = std::max(w.maxbound[s], k.eb[s]);
            segsum[s] += std::min<u64>(k.eb[s], nctx[s]);
            d.obase[s] = osum; d.ocap[s] = al16(k.eb[s] + k.eb[s] / 2 + 256); osum += d.ocap[s];
        }
        d.sidebase = sidesum; d.sidecap = k.side; sidesum += k.side;
        d.reccap = k.rec; recsum += k.rec;
    }
    rbase[n] = (u32)R;
    if (R >= 0xffffffffull) { c->err = "too many reads in one wave"; return SQ_ERR_ARG; }
    w.R = (u32)R;

    // ---- carve the arena
    bool ok = true;
    auto need = [&](void *p) { if (!p) ok = false; return p; };
    w.bd = (BlockDesc *)need(take<BlockDesc>(A, n));
    w.bs = (BlockState *)need(take<BlockState>(A, n));
    w.rbase = (u32 *)need(take<u32>(A, n + 1));
    w.totals = (u32 *)need(take<u32>(A, 6 * (u64)n));
    w.recoff = (u64 *)need(take<u64>(A, n + 1));
    w.nseg = (u32 *)need(take<u32>(A, 4));
    u64 Rr = R ? R : 1;
    w.seqoff = (u32 *)need(take<u32>(A, Rr)); w.idoff = (u32 *)need(take<u32>(A, Rr));
    w.acgtoff = (u32 *)need(take<u32>(A, Rr)); w.qlen2 = (u32 *)need(take<u32>(A, Rr));
    w.qoff = (u32 *)need(take<u32>(A, Rr)); w.noff = (u32 *)need(take<u32>(A, Rr));
    w.tip = (u8 *)need(take<u8>(A, Rr)); w.np = (u8 *)need(take<u8>(A, Rr)); w.ns = (u8 *)need(take<u8>(A, Rr));
    u64 maxsum = std::max(sum[0], std::max(sum[1], sum[2]));
    w.elemA = (u64 *)need(take<u64>(A, maxsum + 1));
    w.elemB = (u64 *)need(take<u64>(A, maxsum + 1));
    for (int s = 0; s < ST_COUNT; s++) w.triples[s] = (u64 *)need(A.take((s == ST_SEQ ? 4 : 8) * (sum[s] + 2)));
    u64 maxtiles = 0;
    for (int s = 0; s < ST_COUNT; s++) maxtiles = std::max(maxtiles, (w.maxbound[s] + SQ_TILE - 1) / SQ_TILE);
    w.hist = (u32 *)need(take<u32>(A, (u64)n * (1u << SQ_SORT_BITS_WIDE) * maxtiles + 1));
    w.scanpart = (u32 *)need(take<u32>(A, (u64)n * (((u64)(1u << SQ_SORT_BITS_WIDE) * maxtiles + 4095) / 4096 + 1) + 1));
    u64 segcap = std::max(segsum[0], std::max(segsum[1], segsum[2])) + 1;
    if (segcap > 0xffffffffull) { c->err = "segment list too large"; return SQ_ERR_ARG; }
    w.segcap = (u32)segcap;
    w.seglist = (u64 *)need(take<u64>(A, segcap));
    w.obuf = (u8 *)need(take<u8>(A, osum + 16));
    w.sidebuf = (u8 *)need(take<u8>(A, sidesum + 16));
    if (host) { w.recbuf = (u8 *)need(take<u8>(A, recsum + 16)); w.recbuf_cap = recsum; }
    // inputs
    for (int i = 0; i < n && ok; i++) {
        const sq_block_in &b = blocks[i];
        BlockDesc &d = bd[i];
        if (host) {
            u8 *ids = (u8 *)need(A.take(al16(b.idbytes) + 16));
            u8 *seq = (u8 *)need(A.take(al16(b.seqbytes) + 16));
            u8 *qual = (u8 *)need(A.take(al16(b.seqbytes) + 16));
            u16 *idlen = (u16 *)need(A.take(al16(2ull * b.nreads) + 16));
            i32 *seqlen = (i32 *)need(A.take(al16(4ull * b.nreads) + 16));
            d.ids = ids; d.seq = seq; d.qual = qual; d.idlen = idlen; d.seqlen = seqlen;
        } else {
            d.ids = b.ids; d.seq = b.seq; d.qual = b.qual; d.idlen = b.idlen; d.seqlen = b.seqlen;
            d.rec = out[i].data; d.reccap = out[i].cap;
        }
        if (c->p.lossy) d.qual_work = (u8 *)need(A.take(al16(b.seqbytes) + 16));
        else d.qual_work = const_cast<u8 *>(d.qual);
    }
    if (!ok) { c->err = "internal: wave does not fit the workspace"; return SQ_ERR_NOMEM; }

    StageTimer &T = c->timer;
    T.begin_wave();
    size_t sp = T.open("h2d", st);
    if (host) {
        for (int i = 0; i < n; i++) {
            const sq_block_in &b = blocks[i];
            BlockDesc &d = bd[i];
            if (b.idbytes) CUDA_OK(c, cudaMemcpyAsync((void *)d.ids, b.ids, b.idbytes, cudaMemcpyHostToDevice, st));
            if (b.seqbytes) {
                CUDA_OK(c, cudaMemcpyAsync((void *)d.seq, b.seq, b.seqbytes, cudaMemcpyHostToDevice, st));
                CUDA_OK(c, cudaMemcpyAsync((void *)d.qual, b.qual, b.seqbytes, cudaMemcpyHostToDevice, st));
            }
            if (b.nreads) {
                CUDA_OK(c, cudaMemcpyAsync((void *)d.idlen, b.idlen, 2ull * b.nreads, cudaMemcpyHostToDevice, st));
                CUDA_OK(c, cudaMemcpyAsync((void *)d.seqlen, b.seqlen, 4ull * b.nreads, cudaMemcpyHostToDevice, st));
            }
        }
    }
    if (c->p.lossy)
        for (int i = 0; i < n; i++)
            if (bd[i].seqbytes) CUDA_OK(c, cudaMemcpyAsync(bd[i].qual_work, bd[i].qual, bd[i].seqbytes, cudaMemcpyDeviceToDevice, st));
    CUDA_OK(c, cudaMemcpyAsync(w.bd, bd.data(), sizeof(BlockDesc) * n, cudaMemcpyHostToDevice, st));
    CUDA_OK(c, cudaMemcpyAsync(w.rbase, rbase.data(), 4 * (n + 1), cudaMemcpyHostToDevice, st));
    CUDA_OK(c, cudaMemsetAsync(w.bs, 0, sizeof(BlockState) * n, st));
    T.close(sp, st);

    // The serial chain kernels (MD5, side streams, one range-coder pass per stream) each use a
    // handful of warps for a long time; they run on their own streams beside the parallel
    // context / sort / replay kernels of the main stream and are joined before the pack.
    u64 *L = &c->launches;
    cudaStream_t s_md5 = c->aux[0], s_side = c->aux[1];
    CUDA_OK(c, cudaEventRecord(c->ev[0], st));                       // inputs resident
    CUDA_OK(c, cudaStreamWaitEvent(s_md5, c->ev[0], 0));
    sp = T.open("md5", s_md5);
    CUDA_OK(c, sqk_md5(w, s_md5, L));
    T.close(sp, s_md5);
    CUDA_OK(c, cudaEventRecord(c->ev[1], s_md5));

    sp = T.open("count", st);
    CUDA_OK(c, sqk_prep(w, st, L));
    if (c->p.lossy) CUDA_OK(c, sqk_rblock(w, c->p.lossy_r, st, L));
    CUDA_OK(c, sqk_count(w, st, L));
    T.close(sp, st);
    CUDA_OK(c, cudaEventRecord(c->ev[2], st));                       // per-read offsets, tip flags, counts
    CUDA_OK(c, cudaStreamWaitEvent(s_side, c->ev[2], 0));
    sp = T.open("side", s_side);
    CUDA_OK(c, sqk_side(w, s_side, L));
    T.close(sp, s_side);
    CUDA_OK(c, cudaEventRecord(c->ev[3], s_side));

    if (train_blob) {
        // the md5 / side kernels launched above are harmless here; only the model replays matter
        int kb[3] = {(int)(2 * c->NS()), c->p.qlevel > 2 ? 20 : 16, 14};
        for (int s = 0; s < ST_COUNT; s++) {
            if (s == ST_NAME && c->p.id_rule) continue;
            u64 *sorted = nullptr;
            CUDA_OK(c, sqk_emit(w, s, w.elemA, st, L));
            CUDA_OK(c, sqk_sort(w, s, kb[s], w.elemA, w.elemB, &sorted, st, L));
            CUDA_OK(c, sqk_replay(w, s, sorted, w.triples[s], st, L));
        }
        CUDA_OK(c, cudaStreamWaitEvent(st, c->ev[1], 0));
        CUDA_OK(c, cudaStreamWaitEvent(st, c->ev[3], 0));
        BlockState hb;
        CUDA_OK(c, cudaMemcpyAsync(&hb, w.bs, sizeof hb, cudaMemcpyDeviceToHost, st));
        CUDA_OK(c, cudaStreamSynchronize(st));
        T.collect();
        if (hb.status) { c->err = "training input rejected"; return hb.status; }
        return SQ_OK;
    }
    static const char *nm_emit[3] = {"seq.ctx", "qual.ctx", "name.ctx"};
    static const char *nm_sort[3] = {"seq.sort", "qual.sort", "name.sort"};
    static const char *nm_rep[3] = {"seq.replay", "qual.replay", "name.replay"};
    static const char *nm_rc[3] = {"seq.rc", "qual.rc", "name.rc"};
    int keybits[3] = {(int)(2 * c->NS()), c->p.qlevel > 2 ? 20 : 16, 14};
    // quality first: its replay + range-coder chain is the longest
    const int order[3] = {ST_QUAL, ST_SEQ, ST_NAME};
    for (int oi = 0; oi < 3; oi++) {
        int s = order[oi];
        cudaStream_t s_rc = c->aux[2 + s];
        if (!(s == ST_NAME && c->p.id_rule)) {
            sp = T.open(nm_emit[s], st);
            CUDA_OK(c, sqk_emit(w, s, w.elemA, st, L));
            T.close(sp, st);
            u64 *sorted = nullptr;
            sp = T.open(nm_sort[s], st);
            CUDA_OK(c, sqk_sort(w, s, keybits[s], w.elemA, w.elemB, &sorted, st, L));
            T.close(sp, st);
            sp = T.open(nm_rep[s], st);
            CUDA_OK(c, sqk_replay(w, s, sorted, w.triples[s], st, L));
            T.close(sp, st);
        }
        CUDA_OK(c, cudaEventRecord(c->ev[4 + s], st));
        CUDA_OK(c, cudaStreamWaitEvent(s_rc, c->ev[4 + s], 0));
        sp = T.open(nm_rc[s], s_rc);
        CUDA_OK(c, sqk_rc(w, s, s_rc, L));
        T.close(sp, s_rc);
        CUDA_OK(c, cudaEventRecord(c->ev[4 + s], s_rc));
    }
    CUDA_OK(c, cudaStreamWaitEvent(st, c->ev[1], 0));
    CUDA_OK(c, cudaStreamWaitEvent(st, c->ev[3], 0));
    for (int s = 0; s < 3; s++) CUDA_OK(c, cudaStreamWaitEvent(st, c->ev[4 + s], 0));
    sp = T.open("pack", st);
    CUDA_OK(c, sqk_pack(w, host ? 1 : 0, st, L));
    T.close(sp, st);

    // ---- results
    u64 small = sizeof(BlockState) * (u64)n + 8 * ((u64)n + 1);
    int rcod = ensure_pin(c, small);
    if (rcod) return rcod;
    BlockState *hbs = (BlockState *)c->h_pin;
    u64 *hoff = (u64 *)(c->h_pin + sizeof(BlockState) * (u64)n);
    CUDA_OK(c, cudaMemcpyAsync(hbs, w.bs, sizeof(BlockState) * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(c, cudaMemcpyAsync(hoff, w.recoff, 8 * ((u64)n + 1), cudaMemcpyDeviceToHost, st));
    CUDA_OK(c, cudaStreamSynchronize(st));
    size_t sp_d2h = host ? T.open("d2h", st) : 0;
    int worst = SQ_OK;
    for (int i = 0; i < n; i++) {
        out[i].status = hbs[i].status;
        out[i].len = hbs[i].reclen;
        if (hbs[i].status == 0 && host) {
            if (out[i].len > out[i].cap) out[i].status = SQ_ERR_CAPACITY;
            else CUDA_OK(c, cudaMemcpyAsync(out[i].data, w.recbuf + hoff[i], out[i].len, cudaMemcpyDeviceToHost, st));
        }
        if (out[i].status != 0 && worst == 0) worst = out[i].status;
    }
    if (host) { T.close(sp_d2h, st); CUDA_OK(c, cudaStreamSynchronize(st)); }
    T.collect();
    c->last = w; c->last_bd = bd; c->last_valid = true;
    if (worst) c->err = "one or more blocks failed; see per-block status";
    return worst;
}

int encode_blocks(sq_ctx *c, const sq_block_in *blocks, int n, sq_block_out *out, bool host) {
    if (!c || !blocks || !out || n <= 0) { if (c) c->err = "bad argument"; return SQ_ERR_ARG; }
    CUDA_OK(c, cudaSetDevice(c->p.device));
    for (int i = 0; i < n; i++) {
        if (blocks[i].seqbytes >= (1ull << 31) || blocks[i].idbytes >= (1ull << 31)) { c->err = "block too large"; return SQ_ERR_ARG; }
        if (blocks[i].nreads && (!blocks[i].idlen || !blocks[i].seqlen)) { c->err = "null length array"; return SQ_ERR_ARG; }
    }
    c->timer.acc.clear();
    c->launches = 0;
    c->last_valid = false;
    int worst = SQ_OK, i = 0;
    c->last_wave_blocks = 0;
    // Waves are filled greedily.  Spreading the blocks evenly over the same number of waves was measured and is not
    // better: a wave's time is set by its per-block chains, and a short last wave finishes sooner (466-block job:
    // 3,331 MB/s greedy against 3,254 MB/s even).
    while (i < n) {
        WaveAcc acc;
        int j = i;
        while (j < n) {
            WaveAcc t = acc;
            t.add(c, block_cost(c, blocks[j], host));
            if (t.total(c) > c->arena.cap || t.B > 60000) break;
            acc = t;
            j++;
        }
        if (j == i) { c->err = "a single block does not fit the device workspace"; return SQ_ERR_NOMEM; }
        int r = encode_wave(c, blocks + i, j - i, out + i, host);
        c->last_wave_blocks = std::max(c->last_wave_blocks, j - i);
        if (r == SQ_ERR_CUDA || r == SQ_ERR_NOMEM || r == SQ_ERR_ARG) return r;
        if (r && !worst) worst = r;
        i = j;
    }
    return worst;
}

}  // namespace

extern "C" int sq_encode_blocks(sq_ctx *c, const sq_block_in *blocks, int n, sq_block_out *out) {
    return encode_blocks(c, blocks, n, out, true);
}
extern "C" int sq_encode_blocks_dev(sq_ctx *c, const sq_block_in *blocks, int n, sq_block_out *out) {
    return encode_blocks(c, blocks, n, out, false);
}

// SeqArcPreProcess::doPreProcess @0x4153b0 -> EncapFqzComp::encodeForModel @0x4228e0 (the *_formodel twins
// update the models without coding and without a reset between the sample blocks) -> SaveModelToMem
// @0x420c00.  The sample blocks are laid end to end as ONE block and pushed through the same context /
// sort / replay kernels as an encode; the replay kernels write each model's final state instead of triples.
extern "C" int sq_train_model(sq_ctx *c, const sq_block_in *blocks, int n, uint8_t *blob, uint64_t cap) {
    if (!c || !blocks || n <= 0 || !blob) { if (c) c->err = "bad argument"; return SQ_ERR_ARG; }
    u64 msize = sq_model_size(c);
    if (cap < msize) { c->err = "model buffer too small"; return SQ_ERR_CAPACITY; }
    CUDA_OK(c, cudaSetDevice(c->p.device));
    u64 nr = 0, ib = 0, sb = 0;
    for (int i = 0; i < n; i++) { nr += blocks[i].nreads; ib += blocks[i].idbytes; sb += blocks[i].seqbytes; }
    if (nr >= (1ull << 31) || sb >= (1ull << 31) || ib >= (1ull << 31)) { c->err = "training sample too large"; return SQ_ERR_ARG; }
    // merged device copy (outside the arena: the arena is carved by encode_wave)
    u8 *d = nullptr;
    u64 o_ids = 0, o_seq = al16(ib) + 16, o_qual = o_seq + al16(sb) + 16, o_il = o_qual + al16(sb) + 16,
        o_sl = o_il + al16(2 * nr) + 16, o_blob = o_sl + al16(4 * nr) + 16, tot = o_blob + msize + 16;
    CUDA_OK(c, cudaMalloc((void **)&d, tot));
    cudaStream_t st = c->stream;
    u64 a = 0, b = 0, r = 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < n && e == cudaSuccess; i++) {
        const sq_block_in &k = blocks[i];
        if (k.idbytes) e = cudaMemcpyAsync(d + o_ids + a, k.ids, k.idbytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && k.seqbytes) e = cudaMemcpyAsync(d + o_seq + b, k.seq, k.seqbytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && k.seqbytes) e = cudaMemcpyAsync(d + o_qual + b, k.qual, k.seqbytes, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && k.nreads) e = cudaMemcpyAsync(d + o_il + 2 * r, k.idlen, 2ull * k.nreads, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && k.nreads) e = cudaMemcpyAsync(d + o_sl + 4 * r, k.seqlen, 4ull * k.nreads, cudaMemcpyHostToDevice, st);
        a += k.idbytes; b += k.seqbytes; r += k.nreads;
    }
    int rc = SQ_OK;
    if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = SQ_ERR_CUDA; }
    if (rc == SQ_OK) {
        u64 l = 0;
        e = sqk_snap_init(d + o_blob, c->snap_name_bytes(), c->nseq(), c->nqual(), st, &l);
        if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = SQ_ERR_CUDA; }
    }
    if (rc == SQ_OK) {
        sq_block_in m{};
        m.nreads = (u32)nr; m.ids = d + o_ids; m.seq = d + o_seq; m.qual = d + o_qual;
        m.idlen = (const u16 *)(d + o_il); m.seqlen = (const i32 *)(d + o_sl);
        m.idbytes = ib; m.seqbytes = sb;
        sq_block_out dummy{};
        sq_params keep = c->p;
        c->p.lossy = 0;                                     // the snapshot is trained on the original qualities
        c->timer.acc.clear();
        c->launches = 0;
        WaveAcc acc;
        acc.add(c, block_cost(c, m, false));
        if (acc.total(c) > c->arena.cap) { c->err = "training sample does not fit the device workspace"; rc = SQ_ERR_NOMEM; }
        else rc = encode_wave(c, &m, 1, &dummy, false, d + o_blob);
        c->p = keep;
        c->last_valid = false;
    }
    if (rc == SQ_OK) {
        e = cudaMemcpy(blob, d + o_blob, msize, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { c->err = cudaGetErrorString(e); rc = SQ_ERR_CUDA; }
    }
    cudaFree(d);
    return rc;
}

extern "C" int sq_debug_fetch(sq_ctx *c, int what, int block, void *dst, uint64_t cap, uint64_t *got) {
    if (!c || !c->last_valid || block < 0 || block >= c->last.B || !dst || !got) return SQ_ERR_ARG;
    CUDA_OK(c, cudaSetDevice(c->p.device));
    BlockState bs;
    CUDA_OK(c, cudaMemcpy(&bs, c->last.bs + block, sizeof bs, cudaMemcpyDeviceToHost));
    if (what == 3) {
        if (cap < 24) return SQ_ERR_CAPACITY;
        u64 v[3] = {bs.nsym[0], bs.nsym[1], bs.nsym[2]};
        memcpy(dst, v, 24);
        *got = 24;
        return SQ_OK;
    }
    if (what < 0 || what > 2) return SQ_ERR_ARG;
    u64 bytes = 8ull * bs.nsym[what];
    if (bytes > cap) return SQ_ERR_CAPACITY;
    if (what == ST_SEQ) {                 // packed 32-bit triples: expand to the common 64-bit form
        std::vector<u32> tmp(bs.nsym[what]);
        CUDA_OK(c, cudaMemcpy(tmp.data(), (const u32 *)c->last.triples[what] + c->last_bd[block].ebase[what], 4ull * bs.nsym[what], cudaMemcpyDeviceToHost));
        u64 *o = (u64 *)dst;
        for (size_t i = 0; i < tmp.size(); i++) o[i] = tr_expand32(tmp[i]);
    } else {
        CUDA_OK(c, cudaMemcpy(dst, c->last.triples[what] + c->last_bd[block].ebase[what], bytes, cudaMemcpyDeviceToHost));
    }
    *got = bytes;
    return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// R-Block on its own
extern "C" int sq_rblock(sq_ctx *c, uint8_t *qual, uint64_t n, double r) {
    if (!c || (!qual && n)) return SQ_ERR_ARG;
    if (n == 0) return SQ_OK;
    CUDA_OK(c, cudaSetDevice(c->p.device));
    Arena &A = c->arena;
    A.reset();
    c->last_valid = false;
    BlockDesc d{};
    BlockDesc *dd = (BlockDesc *)A.take(sizeof d);
    u8 *q = (u8 *)A.take(n + 16);
    if (!dd || !q) return SQ_ERR_NOMEM;
    d.qual_work = q; d.seqbytes = n;
    EncWave w{};
    w.B = 1; w.bd = dd;
    CUDA_OK(c, cudaMemcpyAsync(q, qual, n, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(dd, &d, sizeof d, cudaMemcpyHostToDevice, c->stream));
    u64 l = 0;
    CUDA_OK(c, sqk_rblock(w, r, c->stream, &l));
    CUDA_OK(c, cudaMemcpyAsync(qual, q, n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return SQ_OK;
}

// ---------------------------------------------------------------------------------------------
// decode
namespace {

// EBML-style vint (Encap::getValue @0x4207a0): length from the leading 1 bit
__host__ __device__ int get_vint(const u8 *p, const u8 *end, u64 *v) {
    if (p >= end) return 0;
    int n = 1;
    u8 m = 0x80;
    while (n <= 8 && !(p[0] & m)) { n++; m >>= 1; }
    if (n > 8 || p + n > end) return 0;
    u64 x = p[0] & (m - 1);
    for (int i = 1; i < n; i++) x = (x << 8) | p[i];
    *v = x;
    return n;
}

// walk the tags of one block record (doDecode @0x42eb60 loop over doRealDecode cases): on the host for host buffers,
// in k_parse_records for records that live on the device
__host__ __device__ int parse_record(int md5, int lossy, const u8 *rec, u64 reclen, DecBlock &d) {
    const u8 *p = rec, *end = rec + reclen;
    u64 v;
    int k = get_vint(p, end, &v);
    if (!k || v != 1) return SQ_ERR_FORMAT;          // only Fqz blocks (tag 1); Align blocks are out of scope
    p += k;
    k = get_vint(p, end, &v);
    if (!k || p + k + v > end) return SQ_ERR_FORMAT;
    p += k;
    end = p + v;
    bool have_count = false;
    while (p < end) {
        u64 tag, sz;
        k = get_vint(p, end, &tag); if (!k) return SQ_ERR_FORMAT; p += k;
        k = get_vint(p, end, &sz); if (!k) return SQ_ERR_FORMAT; p += k;
        if (p + sz > end) return SQ_ERR_FORMAT;
        const u8 *pay = p;
        p += sz;
        if (tag >= 32) continue;
        if (tag == 1) { if (sz != 4) return SQ_ERR_FORMAT; d.nreads = pay[0] | (pay[1] << 8) | (pay[2] << 16) | ((u32)pay[3] << 24); have_count = true; continue; }
        int mi = tag == 5 ? 0 : tag == 6 ? 1 : tag == 7 ? 2 : -1;
        if (mi >= 0 && md5 && !(mi == 2 && lossy)) {
            if (sz < 16) return SQ_ERR_FORMAT;
            memcpy(d.md5[mi], pay, 16); d.has_md5[mi] = 1;
            pay += 16; sz -= 16;
        }
        if ((tag == 14 || (tag >= 23 && tag <= 26)) && sz < 4) return SQ_ERR_FORMAT;
        // streams doAlignEncode @0x42d4c0 adds: accepted only when they say "no read of this block mapped"
        if ((tag >= 17 && tag <= 20) || tag == 22) {
            if (sz != 4 || pay[0] || pay[1] || pay[2] || pay[3]) return SQ_ERR_FORMAT;
        } else if ((tag >= 8 && tag <= 13) || tag == 15) {
            if (sz != 8) return SQ_ERR_FORMAT;
        }
        d.off[tag] = (u64)(pay - rec);
        d.len[tag] = (u32)sz;
    }
    if (!have_count) return SQ_ERR_FORMAT;
    if (d.nreads && (!d.len[4] || !d.len[7] || !d.len[6] || !d.len[5])) return SQ_ERR_FORMAT;
    return SQ_OK;
}

__host__ __device__ u32 rec_count(const u8 *rec, const DecBlock &d, int tag) {
    if (!d.len[tag]) return 0;
    const u8 *p = rec + d.off[tag];
    return p[0] | (p[1] << 8) | (p[2] << 16) | ((u32)p[3] << 24);
}

// one thread per device-resident record: its tag walk, and the two counts the host checks against the output capacity
__global__ void k_parse_records(const u8 *const *recs, const u64 *lens, int n, int md5, int lossy, DecBlock *out, u32 *cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DecBlock d;
    memset(&d, 0, sizeof d);
    int r = parse_record(md5, lossy, recs[i], lens[i], d);
    d.status = r;
    cnt[2 * i] = r ? 0u : rec_count(recs[i], d, 14);
    cnt[2 * i + 1] = r ? 0u : rec_count(recs[i], d, 26);
    out[i] = d;
}

int decode_wave(sq_ctx *c, const u8 *const *recs, const u64 *reclens, int n, sq_block_dec *out, bool host,
                const std::vector<DecBlock> &parsed) {
    cudaStream_t st = c->stream;
    Arena &A = c->arena;
    A.reset();
    c->last_valid = false;
    DecWave w{};
    w.B = n; w.nsm = c->nsm;
    w.md5 = c->p.md5; w.lossy = c->p.lossy; w.id_rule = c->p.id_rule; w.qlevel = c->p.qlevel;
    w.nseq = c->nseq(); w.seqmask = w.nseq - 1; w.nqual = c->nqual();
    if (c->p.usemodel) {
        if (!c->d_model) { c->err = "usemodel = 1 but no snapshot loaded (sq_load_model)"; return SQ_ERR_ARG; }
        w.snap_name = c->p.id_rule ? nullptr : c->d_model;
        w.snap_seq = c->d_model + c->snap_name_bytes();
        w.snap_qual = w.snap_seq + (u64)c->nseq() * 4;
    }
    std::vector<DecBlock> db(parsed.begin(), parsed.begin() + n);
    bool ok = true;
    auto need = [&](void *p) { if (!p) ok = false; return p; };
    w.db = (DecBlock *)need(take<DecBlock>(A, n));
    for (int i = 0; i < n && ok; i++) {
        DecBlock &d = db[i];
        sq_block_dec &o = out[i];
        d.idcap = o.idcap; d.seqcap = o.seqcap; d.readcap = o.readcap;
        d.longflag = o.longflag ? 1 : 0;
        if (d.status) continue;
        if (host) {
            d.rec = (u8 *)need(A.take(reclens[i] + 16));
            d.ids = (u8 *)need(A.take(o.idcap + 16));
            d.seq = (u8 *)need(A.take(o.seqcap + 16));
            d.qual = (u8 *)need(A.take(o.seqcap + 16));
            d.idlen = (u16 *)need(A.take(2ull * o.readcap + 16));
            d.seqlen = (i32 *)need(A.take(4ull * o.readcap + 16));
        } else {
            d.rec = recs[i];
            d.ids = o.ids; d.seq = o.seq; d.qual = o.qual; d.idlen = o.idlen; d.seqlen = o.seqlen;
        }
        d.reclen = reclens[i];
        u64 nr = d.nreads;
        d.seqoff = (u32 *)need(A.take(4 * (nr + 1) + 16));
        d.seqtab = (u8 *)need(A.take((u64)w.nseq * 4));
        d.qualtab = (u32 *)need(A.take(((u64)w.nqual + 1) * SQ_WQ_STRIDE * 4));     // + one spare slot (sq_decode.cu: QPend)
        if (!c->p.id_rule) d.nametab = (u32 *)need(A.take(SQ_WNAME_WORDS * 4));
        d.d_tip = (u8 *)need(A.take(nr + 16));
        d.d_maxq = (u8 *)need(A.take(nr + 16));
        d.d_cnt = (u32 *)need(A.take(4 * nr + 16));
        d.d_ch = (u8 *)need(A.take(o.seqcap + 16));
        d.d_pos = (u32 *)need(A.take(4 * o.seqcap + 16));
    }
    if (!ok) { c->err = "internal: decode wave does not fit the workspace"; return SQ_ERR_NOMEM; }
    StageTimer &T = c->timer;
    T.begin_wave();
    size_t sp = T.open("h2d", st);
    if (host)
        for (int i = 0; i < n; i++)
            if (!db[i].status) CUDA_OK(c, cudaMemcpyAsync((void *)db[i].rec, recs[i], reclens[i], cudaMemcpyHostToDevice, st));
    CUDA_OK(c, cudaMemcpyAsync(w.db, db.data(), sizeof(DecBlock) * n, cudaMemcpyHostToDevice, st));
    T.close(sp, st);
    sp = T.open("decode", st);
    CUDA_OK(c, sqk_decode(w, st, &c->launches));
    T.close(sp, st);
    int rc = ensure_pin(c, sizeof(DecBlock) * (u64)n);
    if (rc) return rc;
    DecBlock *hdb = (DecBlock *)c->h_pin;
    CUDA_OK(c, cudaMemcpyAsync(hdb, w.db, sizeof(DecBlock) * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(c, cudaStreamSynchronize(st));
    size_t sp_d2h = host ? T.open("d2h", st) : 0;
    int worst = SQ_OK;
    for (int i = 0; i < n; i++) {
        sq_block_dec &o = out[i];
        o.status = hdb[i].status; o.nreads = hdb[i].nreads; o.idbytes = hdb[i].idbytes; o.seqbytes = hdb[i].seqbytes;
        if (o.status == 0 && host) {
            if (o.idbytes) CUDA_OK(c, cudaMemcpyAsync(o.ids, db[i].ids, o.idbytes, cudaMemcpyDeviceToHost, st));
            if (o.seqbytes) {
                CUDA_OK(c, cudaMemcpyAsync(o.seq, db[i].seq, o.seqbytes, cudaMemcpyDeviceToHost, st));
                CUDA_OK(c, cudaMemcpyAsync(o.qual, db[i].qual, o.seqbytes, cudaMemcpyDeviceToHost, st));
            }
            if (o.nreads) {
                CUDA_OK(c, cudaMemcpyAsync(o.idlen, db[i].idlen, 2ull * o.nreads, cudaMemcpyDeviceToHost, st));
                CUDA_OK(c, cudaMemcpyAsync(o.seqlen, db[i].seqlen, 4ull * o.nreads, cudaMemcpyDeviceToHost, st));
            }
        }
        if (o.status && !worst) worst = o.status;
    }
    if (host) { T.close(sp_d2h, st); CUDA_OK(c, cudaStreamSynchronize(st)); }
    T.collect();
    if (worst) c->err = "one or more blocks failed to decode; see per-block status";
    return worst;
}

u64 dec_block_cost(const sq_ctx *c, u64 reclen, const sq_block_dec &o, u32 nreads, bool host) {
    u64 t = sizeof(DecBlock) + 20 * 256;
    if (host) t += reclen + o.idcap + 2 * o.seqcap + 6ull * o.readcap;
    t += 4ull * (nreads + 1) + (u64)c->nseq() * 4 + ((u64)c->nqual() + 1) * SQ_WQ_STRIDE * 4;
    if (!c->p.id_rule) t += SQ_WNAME_WORDS * 4;
    t += 6ull * nreads + 5 * o.seqcap;
    return t;
}

int decode_blocks(sq_ctx *c, const u8 *const *recs, const u64 *reclens, int n, sq_block_dec *out, bool host) {
    if (!c || !recs || !reclens || !out || n <= 0) { if (c) c->err = "bad argument"; return SQ_ERR_ARG; }
    CUDA_OK(c, cudaSetDevice(c->p.device));
    c->timer.acc.clear();
    c->launches = 0;
    c->last_wave_blocks = 0;
    // the tag walk of every record: on the host for host buffers, by one small kernel for device-resident records
    // (only the parsed headers come back, not the records)
    std::vector<DecBlock> parsed(n);
    std::vector<u32> cnt(2 * (size_t)n);
    if (host) {
        for (int i = 0; i < n; i++) {
            memset(&parsed[i], 0, sizeof(DecBlock));
            int r = parse_record(c->p.md5, c->p.lossy, recs[i], reclens[i], parsed[i]);
            parsed[i].status = r;
            cnt[2 * i] = r ? 0u : rec_count(recs[i], parsed[i], 14);
            cnt[2 * i + 1] = r ? 0u : rec_count(recs[i], parsed[i], 26);
        }
    } else {
        Arena &A = c->arena;
        A.reset();
        const u8 **d_recs = (const u8 **)A.take(8 * (u64)n);
        u64 *d_lens = (u64 *)A.take(8 * (u64)n);
        DecBlock *d_parsed = (DecBlock *)A.take(sizeof(DecBlock) * (u64)n);
        u32 *d_cnt = (u32 *)A.take(8 * (u64)n);
        if (!d_recs || !d_lens || !d_parsed || !d_cnt) { c->err = "the device workspace is too small for the record headers"; return SQ_ERR_NOMEM; }
        cudaStream_t st = c->stream;
        CUDA_OK(c, cudaMemcpyAsync(d_recs, recs, 8 * (u64)n, cudaMemcpyHostToDevice, st));
        CUDA_OK(c, cudaMemcpyAsync(d_lens, reclens, 8 * (u64)n, cudaMemcpyHostToDevice, st));
        k_parse_records<<<(n + 63) / 64, 64, 0, st>>>(d_recs, d_lens, n, c->p.md5, c->p.lossy, d_parsed, d_cnt);
        CUDA_OK(c, cudaGetLastError());
        c->launches++;
        CUDA_OK(c, cudaMemcpyAsync(parsed.data(), d_parsed, sizeof(DecBlock) * (u64)n, cudaMemcpyDeviceToHost, st));
        CUDA_OK(c, cudaMemcpyAsync(cnt.data(), d_cnt, 8 * (u64)n, cudaMemcpyDeviceToHost, st));
        CUDA_OK(c, cudaStreamSynchronize(st));
    }
    for (int i = 0; i < n; i++) {
        if (parsed[i].status) continue;
        if (parsed[i].nreads > out[i].readcap) parsed[i].status = SQ_ERR_CAPACITY;
        else if (cnt[2 * i] > out[i].seqcap || cnt[2 * i + 1] > out[i].seqcap) parsed[i].status = SQ_ERR_FORMAT;
    }
    int worst = SQ_OK, i = 0;
    while (i < n) {
        u64 tot = 64 * 1024;
        int j = i;
        while (j < n) {
            u64 k = dec_block_cost(c, reclens[j], out[j], parsed[j].nreads, host);
            if (tot + k > c->arena.cap) break;
            tot += k;
            j++;
        }
        if (j == i) { c->err = "a single block does not fit the device workspace"; return SQ_ERR_NOMEM; }
        std::vector<DecBlock> part(parsed.begin() + i, parsed.begin() + j);
        int r = decode_wave(c, recs + i, reclens + i, j - i, out + i, host, part);
        c->last_wave_blocks = std::max(c->last_wave_blocks, j - i);
        if (r == SQ_ERR_CUDA || r == SQ_ERR_NOMEM || r == SQ_ERR_ARG) return r;
        if (r && !worst) worst = r;
        i = j;
    }
    return worst;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Aligned-mode side streams: compressOrder @0x424b70, compressPERelation @0x422be0, compressAlignInfo_{Pos, Mis, Rev,
// CigaL, CigaV} @0x425d70-0x426c0f and their decompress* twins.  Each stream is one serial chain (a fresh SIMPLE_MODEL<N>,
// N <= 16, the model in registers / local memory): one thread per stream, all streams of the call in one launch.
namespace {

struct SideJob { const u8 *sym_in; u8 *sym_out; const u8 *bytes_in; u8 *bytes_out; u64 cap; u32 nsym; i32 model; u64 len; i32 status; };

template <int N>
__device__ void side_encode_n(SideJob &j) {
    SmModel<N> m;
    m.init();
    RcEnc rc;
    rc.start(j.bytes_out, j.bytes_out + j.cap);
    for (u32 i = 0; i < j.nsym; i++) {
        u32 s = j.sym_in[i];
        if (s >= (u32)N) { j.status = SQ_ERR_INPUT; return; }
        m.encode(rc, s);
    }
    rc.finish();
    j.len = (u64)(rc.out - j.bytes_out);
    j.status = rc.overflow ? SQ_ERR_CAPACITY : SQ_OK;
}
template <int N>
__device__ void side_decode_n(SideJob &j) {
    SmModel<N> m;
    m.init();
    RcDec rc;
    rc.start(j.bytes_in, j.bytes_in + j.cap);
    for (u32 i = 0; i < j.nsym; i++) j.sym_out[i] = (u8)m.decode(rc);
    j.len = j.cap;
    j.status = SQ_OK;
}
template <bool ENC, int N>
__device__ void side_run(SideJob &j) { if (ENC) side_encode_n<N>(j); else side_decode_n<N>(j); }

template <bool ENC>
__global__ void k_side_streams(SideJob *jobs, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    SideJob j = jobs[i];
    switch (j.model) {          // the sizes the reference uses get their own instance; the rest share the largest
        case 2: side_run<ENC, 2>(j); break;
        case 4: side_run<ENC, 4>(j); break;
        case 5: side_run<ENC, 5>(j); break;
        case 8: side_run<ENC, 8>(j); break;
        case 9: side_run<ENC, 9>(j); break;
        default: j.status = SQ_ERR_ARG; break;
    }
    jobs[i].len = j.len;
    jobs[i].status = j.status;
}

int side_streams(sq_ctx *c, sq_side_stream *ss, int n, bool enc) {
    if (!c || !ss || n <= 0) { if (c) c->err = "bad argument"; return SQ_ERR_ARG; }
    CUDA_OK(c, cudaSetDevice(c->p.device));
    for (int i = 0; i < n; i++) {
        const int N = ss[i].model;
        if (!(N == 2 || N == 4 || N == 5 || N == 8 || N == 9) || (ss[i].nsym && !ss[i].sym) || !ss[i].bytes || (!enc && ss[i].cap < 8)) {
            c->err = "side stream: model size must be 2, 4, 5, 8 or 9, buffers non-null, a stream at least 8 bytes"; return SQ_ERR_ARG;
        }
    }
    Arena &A = c->arena;
    A.reset();
    c->last_valid = false;
    cudaStream_t st = c->stream;
    std::vector<SideJob> jobs(n);
    SideJob *d_jobs = (SideJob *)A.take(sizeof(SideJob) * (u64)n);
    bool ok = d_jobs != nullptr;
    for (int i = 0; i < n && ok; i++) {
        SideJob &j = jobs[i];
        memset(&j, 0, sizeof j);
        j.model = ss[i].model; j.nsym = ss[i].nsym; j.cap = ss[i].cap;
        u8 *d_sym = (u8 *)A.take((u64)ss[i].nsym + 16), *d_bytes = (u8 *)A.take(ss[i].cap + 16);
        if (!d_sym || !d_bytes) { ok = false; break; }
        j.sym_in = j.sym_out = d_sym; j.bytes_in = j.bytes_out = d_bytes;
        if (enc) { if (ss[i].nsym) CUDA_OK(c, cudaMemcpyAsync(d_sym, ss[i].sym, ss[i].nsym, cudaMemcpyHostToDevice, st)); }
        else CUDA_OK(c, cudaMemcpyAsync(d_bytes, ss[i].bytes, ss[i].cap, cudaMemcpyHostToDevice, st));
    }
    if (!ok) { c->err = "side streams do not fit the device workspace"; return SQ_ERR_NOMEM; }
    CUDA_OK(c, cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(SideJob) * (u64)n, cudaMemcpyHostToDevice, st));
    if (enc) k_side_streams<true><<<(n + 31) / 32, 32, 0, st>>>(d_jobs, n);
    else k_side_streams<false><<<(n + 31) / 32, 32, 0, st>>>(d_jobs, n);
    CUDA_OK(c, cudaGetLastError());
    c->launches = 1;
    std::vector<SideJob> res(n);
    CUDA_OK(c, cudaMemcpyAsync(res.data(), d_jobs, sizeof(SideJob) * (u64)n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(c, cudaStreamSynchronize(st));
    int worst = SQ_OK;
    for (int i = 0; i < n; i++) {
        ss[i].status = res[i].status;
        ss[i].len = res[i].status == SQ_OK ? res[i].len : 0;
        if (res[i].status != SQ_OK) { if (!worst) worst = res[i].status; continue; }
        if (enc) CUDA_OK(c, cudaMemcpyAsync(ss[i].bytes, jobs[i].bytes_out, res[i].len, cudaMemcpyDeviceToHost, st));
        else if (ss[i].nsym) CUDA_OK(c, cudaMemcpyAsync(ss[i].sym, jobs[i].sym_out, ss[i].nsym, cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(c, cudaStreamSynchronize(st));
    if (worst) c->err = "one or more side streams failed; see per-stream status";
    return worst;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Aligned mode, decode side: the bases of mapped reads from the reference (getRealPos @0x42de30, AlignInfoToSeq @0x42e020).
// k_align_walk: one thread per block walks its reads in order -- positions and mismatch offsets are variable-width bit
// fields, so where read j's fields start depends on every read before it -- and leaves one descriptor per mapped read.
// k_align_fill: one warp per mapped read copies the reference segment (coalesced), applies its substitutions, reverses and
// complements: the part that touches every base runs at memory speed.
namespace {

struct AlignDesc { u64 pos, off; i32 len; u32 cl_off, cv_off; u8 rev, nmis; };
struct AlignJob {
    u32 nreads; i32 pe;
    const i32 *seqlen; const u8 *order, *pos, *mis, *rev, *cigal, *cigav, *perel;
    u32 npos, nalign, ncigal, ncigav, nperel, insbits;
    u8 *seq; u64 seqbytes;
    AlignDesc *desc; i32 status; u32 nmapped;
};

SQ_HD u32 bitnum64(u64 v) { u32 n = 0; while (v) { n++; v >>= 1; } return n; }        // getbitnum @0x40d470
SQ_D u64 read_bits(const u8 *bits, u32 nbits, u32 &cur, u32 w) {                    // LSB first (getOffsePos @0x42ec80, getCigal @0x42ecc0)
    u64 v = 0;
    for (u32 k = 0; k < w; k++) if (cur + k < nbits) v += (u64)bits[cur + k] << k;
    cur += w;
    return v;
}

__global__ void k_align_walk(AlignJob *jobs, int n, u64 reflen) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    AlignJob &J = jobs[b];
    const u32 W = bitnum64(reflen) > 2 ? bitnum64(reflen) - 2 : 0;                   // SeqArcParam+0x1858
    u32 pc = 0, cl = 0, cv = 0, j = 0, pr = 0;
    u64 cur = 0, off = 0;
    i32 status = SQ_OK;
    for (u32 i = 0; i < J.nreads && status == SQ_OK; i++) {
        const i32 len = J.seqlen[i];
        if (len < 0 || off + (u64)len > J.seqbytes) { status = SQ_ERR_FORMAT; break; }
        const u32 o = J.order[i];
        if (o) {
            if (j >= J.nalign || o > 4) { status = SQ_ERR_FORMAT; break; }
            const bool second = J.pe && (i & 1) && J.order[i - 1];
            u64 pos;
            if (!second) pos = read_bits(J.pos, J.npos, pc, W) + ((u64)(o - 1) << W);
            else {
                const u32 rel = pr < J.nperel ? J.perel[pr] : 4u;
                pr++;
                pos = cur;
                if (rel == 1) pos += read_bits(J.pos, J.npos, pc, J.insbits);
                else if (rel == 0) pos -= read_bits(J.pos, J.npos, pc, J.insbits);
                else if (rel == 2) pos = read_bits(J.pos, J.npos, pc, bitnum64(cur));
                else if (rel == 3) pos += read_bits(J.pos, J.npos, pc, bitnum64(reflen - cur));
            }
            cur = pos;
            const u32 nmis = J.mis[j];
            if (pos < 1 || pos - 1 + (u64)len > reflen || nmis > 16 || pc > J.npos) { status = SQ_ERR_FORMAT; break; }
            AlignDesc d;
            d.pos = pos; d.off = off; d.len = len; d.cl_off = cl; d.cv_off = cv; d.rev = J.rev[j]; d.nmis = (u8)nmis;
            u32 prev = 0;
            for (u32 m = 0; m < nmis; m++) prev += (u32)read_bits(J.cigal, J.ncigal, cl, bitnum64((u64)(len - (i32)prev)));
            cv += nmis;
            if (cl > J.ncigal || cv > J.ncigav) { status = SQ_ERR_FORMAT; break; }
            J.desc[j++] = d;
        }
        off += (u64)len;
    }
    J.status = status;
    J.nmapped = status == SQ_OK ? j : 0;
}

__global__ void __launch_bounds__(128) k_align_fill(const AlignJob *jobs, const u8 *ref) {
    const AlignJob &J = jobs[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (u32 j = warp; j < J.nmapped; j += nwarps) {
        const AlignDesc d = J.desc[j];
        u32 where[16], var[16];
        u32 cl = d.cl_off, prev = 0;
        for (u32 m = 0; m < d.nmis; m++) {                // a handful of bit fields: every lane decodes them
            prev += (u32)read_bits(J.cigal, J.ncigal, cl, bitnum64((u64)(d.len - (i32)prev)));
            where[m] = prev;
            var[m] = J.cigav[d.cv_off + m];
        }
        for (i32 k = lane; k < d.len; k += 32) {
            u32 code = seq_val(ref[d.pos - 1 + (u64)k]);
            if (code > 3) code = 4;
            for (u32 m = 0; m < d.nmis; m++)
                if (where[m] == (u32)k) {                 // mapvar2base @0x40db20: which of the three other bases, 3 = N
                    const u32 v = var[m];
                    const u32 tab = code == 0 ? 0x4312u : code == 1 ? 0x4302u : code == 2 ? 0x4301u : code == 3 ? 0x4012u : 0x0000u;
                    code = (tab >> (4 * (v & 3))) & 15u;
                }
            const u32 out = d.rev ? (u32)(d.len - 1 - k) : (u32)k;
            const u32 c2 = code > 3 ? 4u : (d.rev ? 3u - code : code);     // getch @0x40dc50
            J.seq[d.off + out] = (u8)"ACGTN"[c2];
        }
    }
}

int align_rebuild(sq_ctx *c, sq_align_block *bl, int n) {
    if (!c || !bl || n <= 0) { if (c) c->err = "bad argument"; return SQ_ERR_ARG; }
    if (!c->d_ref) { c->err = "no reference loaded (sq_set_reference)"; return SQ_ERR_ARG; }
    CUDA_OK(c, cudaSetDevice(c->p.device));
    Arena &A = c->arena;
    A.reset();
    c->last_valid = false;
    cudaStream_t st = c->stream;
    std::vector<AlignJob> jobs(n);
    AlignJob *d_jobs = (AlignJob *)A.take(sizeof(AlignJob) * (u64)n);
    bool ok = d_jobs != nullptr;
    u32 maxalign = 0;
    auto up = [&](const void *h, u64 bytes) -> u8 * {
        u8 *d = (u8 *)A.take(bytes + 16);
        if (!d) { ok = false; return nullptr; }
        if (bytes && cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) ok = false;
        return d;
    };
    for (int i = 0; i < n && ok; i++) {
        const sq_align_block &b = bl[i];
        if (!b.seqlen || !b.order || !b.seq || (b.npos && !b.pos) || (b.nalign && (!b.mis || !b.rev)) || (b.ncigal && !b.cigal) ||
            (b.ncigav && !b.cigav) || (b.nperel && !b.perel)) { c->err = "sq_align_block: null array"; return SQ_ERR_ARG; }
        AlignJob &j = jobs[i];
        memset(&j, 0, sizeof j);
        j.nreads = b.nreads; j.pe = b.pe;
        j.npos = b.npos; j.nalign = b.nalign; j.ncigal = b.ncigal; j.ncigav = b.ncigav; j.nperel = b.nperel; j.insbits = b.insbits;
        j.seqbytes = b.seqbytes;
        j.seqlen = (const i32 *)up(b.seqlen, 4ull * b.nreads);
        j.order = up(b.order, b.nreads); j.pos = up(b.pos, b.npos); j.mis = up(b.mis, b.nalign); j.rev = up(b.rev, b.nalign);
        j.cigal = up(b.cigal, b.ncigal); j.cigav = up(b.cigav, b.ncigav); j.perel = up(b.perel, b.nperel);
        j.seq = up(b.seq, b.seqbytes);
        j.desc = (AlignDesc *)A.take(sizeof(AlignDesc) * ((u64)b.nalign + 1));
        if (!j.desc) ok = false;
        maxalign = std::max(maxalign, b.nalign);
    }
    if (!ok) { c->err = "aligned blocks do not fit the device workspace"; return SQ_ERR_NOMEM; }
    CUDA_OK(c, cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(AlignJob) * (u64)n, cudaMemcpyHostToDevice, st));
    k_align_walk<<<(n + 31) / 32, 32, 0, st>>>(d_jobs, n, c->ref_len);
    CUDA_OK(c, cudaGetLastError());
    const u32 gx = std::max<u32>(1, std::min<u32>((maxalign + 3) / 4, (u32)(c->nsm > 0 ? c->nsm : 148) * 8));
    k_align_fill<<<dim3(gx, n), 128, 0, st>>>(d_jobs, c->d_ref);
    CUDA_OK(c, cudaGetLastError());
    c->launches = 2;
    std::vector<AlignJob> res(n);
    CUDA_OK(c, cudaMemcpyAsync(res.data(), d_jobs, sizeof(AlignJob) * (u64)n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(c, cudaStreamSynchronize(st));
    int worst = SQ_OK;
    for (int i = 0; i < n; i++) {
        bl[i].status = res[i].status;
        if (res[i].status != SQ_OK) { if (!worst) worst = res[i].status; continue; }
        if (bl[i].seqbytes) CUDA_OK(c, cudaMemcpyAsync(bl[i].seq, jobs[i].seq, bl[i].seqbytes, cudaMemcpyDeviceToHost, st));
    }
    CUDA_OK(c, cudaStreamSynchronize(st));
    if (worst) c->err = "one or more aligned blocks are inconsistent; see per-block status";
    return worst;
}

}  // namespace

extern "C" int sq_set_reference(sq_ctx *c, const uint8_t *bases, uint64_t n) {
    if (!c || !bases || !n) { if (c) c->err = "bad argument"; return SQ_ERR_ARG; }
    CUDA_OK(c, cudaSetDevice(c->p.device));
    if (c->d_ref) { cudaFree(c->d_ref); c->d_ref = nullptr; c->ref_len = 0; }
    CUDA_OK(c, cudaMalloc((void **)&c->d_ref, n + 16));
    CUDA_OK(c, cudaMemcpy(c->d_ref, bases, n, cudaMemcpyHostToDevice));
    c->ref_len = n;
    return SQ_OK;
}
extern "C" int sq_align_rebuild(sq_ctx *c, sq_align_block *blocks, int n) { return align_rebuild(c, blocks, n); }

extern "C" int sq_side_encode(sq_ctx *c, sq_side_stream *streams, int n) { return side_streams(c, streams, n, true); }
extern "C" int sq_side_decode(sq_ctx *c, sq_side_stream *streams, int n) { return side_streams(c, streams, n, false); }

extern "C" int sq_decode_blocks(sq_ctx *c, const uint8_t *const *recs, const uint64_t *reclens, int n, sq_block_dec *out) {
    return decode_blocks(c, recs, reclens, n, out, true);
}
extern "C" int sq_decode_blocks_dev(sq_ctx *c, const uint8_t *const *recs, const uint64_t *reclens, int n, sq_block_dec *out) {
    return decode_blocks(c, recs, reclens, n, out, false);
}
