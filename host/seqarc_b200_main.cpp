// seqarc_b200 -- command-line driver standing where SeqArc's main / SeqArcContext / reader and worker
// threads stand (main @0x41fd40, doReadAndEncode @0x41a4e0, doTask @0x433dd0), for the no-reference
// path:  -c -1 a.fq [-2 b.fq] -o out [-l r] [-n] [-f]      ->  out.arc
//        -d in.arc -o out [-f]                                ->  out.fastq | out_1.fastq + out_2.fastq
// FASTQ parsing, block cutting and .arc packing run here on the host; every coder runs on the GPU
// behind the C ABI of include/seqarc_b200.h.  ./seqarc.config (BlockSize(M), Slevel, Qlevel) is read
// like the reference does (parseOptFromFile @0x4062e0) because none of the three is stored in the .arc.
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <fcntl.h>
#include <future>
#include <map>
#include <memory>
#include <unistd.h>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/seqarc_b200.h"
#include "arc_file.h"
#include "fastq_block.h"
#include "id_process.h"

using namespace sqhost;

struct Opt {
    bool compress = false, decompress = false, force = false, md5 = true, lossy = false;
    double lossy_r = 0;
    std::string in1, in2, out, arc;
    std::string ref;            // reference FASTA given on the command line (SeqArc -c [ref.fa] ... / -d [ref.fa] x.arc)
    int block_mb = 50, slevel = 3, qlevel = 2, batch = 0;
    std::vector<int> devices;   // -g 0,1,...  (default: every visible GPU)
    int force_usemodel = -1;    // --usemodel 0|1: test hook (the reference decides from the input size alone)
    bool parse_only = false;    // --parse-only: print what the command line means and exit (no GPU needed)
    bool verbose = false;
};

static void read_config(Opt &o) {
    FILE *f = fopen("./seqarc.config", "r");
    if (!f) return;
    char line[256];
    while (fgets(line, sizeof line, f)) {
        char *c = strchr(line, ':');
        if (!c) continue;
        *c = 0;
        int v = atoi(c + 1);
        if (!strcmp(line, "BlockSize(M)")) o.block_mb = v;
        else if (!strcmp(line, "Slevel")) o.slevel = v;
        else if (!strcmp(line, "Qlevel")) o.qlevel = v;
    }
    fclose(f);
}

static std::string base_name(const std::string &p) {      // stored in trailer params 13 / 14, without a ".gz" suffix
    size_t k = p.find_last_of('/');
    std::string b = k == std::string::npos ? p : p.substr(k + 1);
    if (InFile::is_gz_name(b)) b.resize(b.size() - 3);
    return b;
}
static uint64_t file_size(const std::string &p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0 ? (uint64_t)st.st_size : 0;
}
static bool exists(const std::string &p) { struct stat st; return stat(p.c_str(), &st) == 0; }

static void die(const std::string &m) { fprintf(stderr, "seqarc_b200: %s\n", m.c_str()); exit(1); }

static sq_ctx *make_ctx(const Opt &o, int device, bool usemodel, bool lossy, double r, bool md5, int id_rule, int ref_mode = 0) {
    sq_params p{};
    p.ref_mode = ref_mode;
    p.device = device; p.usemodel = usemodel; p.lossy = lossy; p.lossy_r = r; p.md5 = md5;
    p.slevel = o.slevel; p.qlevel = o.qlevel; p.id_rule = id_rule;
    sq_ctx *c = nullptr;
    if (sq_create(&p, &c) != SQ_OK) die(std::string("sq_create: ") + sq_last_error(nullptr));
    return c;
}

static sq_block_in as_in(const MemBuf &m) {
    sq_block_in b{};
    b.nreads = m.nreads();
    b.longflag = m.longflag;
    b.ids = m.ids.data(); b.seq = m.seq.data(); b.qual = m.qual.data();
    b.idlen = m.idlen.data(); b.seqlen = m.seqlen.data();
    b.idbytes = m.ids.size(); b.seqbytes = m.seq.size();
    return b;
}

// SeqArcPreProcess::doPreProcess @0x4153b0: the (up to five) head blocks the snapshot is trained on.
// Each sample block takes whole records until it holds at least BlockSize bytes (measured against the
// reference with oracle/sizefake.c: 5 x 4298 reads of 244 B at BlockSize 1 MiB, single-end).
static void sample_blocks(const Opt &o, std::vector<MemBuf> &out) {
    bool pe = !o.in2.empty();
    InFile f1, f2;
    if (!f1.open(o.in1) || (pe && !f2.open(o.in2))) die("cannot open input");
    // paired-end: BlockSize / 2 per mate file, the same number of records from both files
    // (measured: 5 x 1461 pairs of 359 B per mate at BlockSize 1 MiB)
    uint64_t quota = ((uint64_t)o.block_mb << 20) / (pe ? 2 : 1);
    std::vector<char> t1, t2;
    auto record = [&](InFile &f, std::vector<char> &t) {
        for (int k = 0; k < 4; k++)
            if (!f.getline(t)) return false;
        return true;
    };
    for (int b = 0; b < 5; b++) {
        t1.clear(); t2.clear();
        while (t1.size() < quota) {
            if (!record(f1, t1)) break;
            if (pe && !record(f2, t2)) die("paired files disagree in record count");
        }
        if (t1.empty()) break;
        out.emplace_back();
        parse_block(t1, t2, pe, out.back());
    }
}

// ---- the job sharded over the GPUs of the box ------------------------------------------------------------
// Blocks are independent (every model resets per block), so the only thing the GPUs share is the ORDER of the records
// in the archive: the reference appends each block at "previous offset + previous length" (addBlockPEInfo @0x412f30,
// @0x412f6d-0x412f76), and its -t 1 output is the byte-exact target.  One host thread and one sq_ctx per GPU; the
// reader deals out consecutive batches of blocks, the writer takes the encoded batches back in batch order.
static std::vector<int> pick_devices(const Opt &o) {
    int n = sq_device_count();
    if (n <= 0) die("no CUDA device (this program has no CPU path)");
    std::vector<int> d = o.devices;
    if (d.empty()) for (int i = 0; i < n; i++) d.push_back(i);
    for (int x : d) if (x < 0 || x >= n) die("device ordinal out of range in -g");
    return d;
}

// First touch of freshly allocated memory costs a page fault per 4 KiB; done by the one thread that later receives the
// device-to-host copies it would serialise the whole batch (measured: 1.3 GB/s into fresh pages).  Touch in parallel.
static void prefault(const std::vector<std::pair<uint8_t *, size_t>> &bufs) {
    std::atomic<size_t> next{0};
    unsigned nt = std::max(1u, std::min<unsigned>({std::thread::hardware_concurrency(), 32u, (unsigned)bufs.size()}));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++)
        pool.emplace_back([&]() {
            for (size_t i; (i = next.fetch_add(1)) < bufs.size();)
                for (size_t o = 0; o < bufs[i].second; o += 4096) ((volatile uint8_t *)bufs[i].first)[o] = 0;
        });
    for (auto &th : pool) th.join();
}

struct EncBatch {
    uint64_t seq = 0;
    std::vector<MemBuf> mbs;
    std::vector<std::unique_ptr<uint8_t[]>> recs;
    std::vector<uint64_t> lens;
};

static int do_compress(const Opt &o) {
    std::string arcname = o.out + ".arc";
    if (exists(arcname) && !o.force) die(arcname + " exists (use -f)");
    bool pe = !o.in2.empty();
    uint64_t total = file_size(o.in1) + (pe ? file_size(o.in2) : 0);
    // doCheckSetEncodeOpt @0x408104-0x40831f: usemodel when total * 0.0023947358 > 33,693,269; off for Qlevel > 2
    // (single-precision arithmetic as in the binary; a .gz input's size is scaled by 2.73 first, constants @0x449948 / @0x449950)
    bool gz = InFile::is_gz_name(o.in1);
    long long est = (long long)((float)(long long)total * 0.0023947358f);
    if (gz) est = (long long)((float)est * 2.73f);
    bool usemodel = est > 33693269 && o.qlevel <= 2;
    if (o.force_usemodel >= 0) usemodel = o.force_usemodel != 0;
    ArcParams ap;
    ap.pe = pe; ap.usemodel = usemodel; ap.lossy = o.lossy; ap.md5 = o.md5;
    ap.name1 = base_name(o.in1); ap.name2 = base_name(o.in2);
    ap.gz = gz;
    if (!o.ref.empty()) {
        // A reference index turns on doAlignEncode @0x42d4c0.  Alignment itself is outside this path: every read is
        // coded as unmapped, which is what the reference writes when nothing maps ("Align block:0") -- same entropy
        // streams, empty alignment streams, param 1 = 0 and the reference's MD5 (the <ref>.md5 file `SeqArc -i`
        // leaves beside the index, i.e. the MD5 of the FASTA file) in trailer TAG 8.
        ap.has_ref = true;
        std::vector<uint8_t> d;
        FILE *f = fopen((o.ref + ".md5").c_str(), "rb");
        if (f) { d.resize(16); if (fread(d.data(), 1, 16, f) != 16) d.clear(); fclose(f); }
        if (d.size() == 16) memcpy(ap.ref_md5, d.data(), 16);
        else {
            f = fopen(o.ref.c_str(), "rb");
            if (!f) die("cannot open reference " + o.ref);
            std::vector<uint8_t> all;
            uint8_t buf[1 << 16];
            size_t n;
            while ((n = fread(buf, 1, sizeof buf, f)) > 0) all.insert(all.end(), buf, buf + n);
            fclose(f);
            md5_digest(all.data(), all.size(), ap.ref_md5);
        }
        if (o.verbose) fprintf(stderr, "reference given: reads are not aligned, every block is written as an all-unmapped block\n");
    }
    // SeqArcPreProcess::doPreProcess @0x4153b0: the ID-bin analysis runs over the head sample blocks and stops
    // at the first block whose IDs are not rule-predictable; the same blocks train the snapshot when usemodel
    std::vector<MemBuf> sample;
    sample_blocks(o, sample);
    bool okbins = true;
    for (auto &m : sample)
        if (okbins) okbins = analyse_id_bins(m, pe, ap.idbin);
    ap.comment = !sample.empty() && sample[0].comment;
    if (o.verbose) fprintf(stderr, "bOrderBin %d petype %d\n", ap.idbin[0], ap.idbin[1]);
    const std::vector<int> devs = pick_devices(o);
    const int ndev = (int)devs.size();
    const int ref_mode = ap.has_ref ? (pe ? 2 : 1) : 0;
    std::vector<sq_ctx *> ctx(ndev, nullptr);
    {
        std::vector<std::thread> th;                   // workspace allocation of the devices in parallel
        for (int d = 0; d < ndev; d++) th.emplace_back([&, d]() { ctx[d] = make_ctx(o, devs[d], usemodel, o.lossy, o.lossy_r, o.md5, ap.idbin[0], ref_mode); });
        for (auto &t : th) t.join();
    }
    if (usemodel) {
        std::vector<sq_block_in> in;
        for (auto &m : sample) in.push_back(as_in(m));
        ap.model.resize(sq_model_size(ctx[0]));
        if (sq_train_model(ctx[0], in.data(), (int)in.size(), ap.model.data(), ap.model.size()) != SQ_OK)
            die(std::string("training: ") + sq_last_error(ctx[0]));
        for (int d = 0; d < ndev; d++)
            if (sq_load_model(ctx[d], ap.model.data(), ap.model.size()) != SQ_OK) die(sq_last_error(ctx[d]));
    }
    std::vector<MemBuf>().swap(sample);
    BlockReader rd(o.in1, o.in2, (uint64_t)o.block_mb << 20);
    if (!rd.ok()) die("cannot open input");
    ArcWriter wr;
    if (!wr.open(arcname)) die("cannot create " + arcname);
    // Blocks per GPU call.  The chain kernels take the same time for one block as for a wave of them, so a batch
    // should be this GPU's whole share of the job when the workspace holds it (~0.75 GB per default-size block).
    uint64_t est_blocks = std::max<uint64_t>(1, ((gz ? total * 4 : total) + ((uint64_t)o.block_mb << 20) - 1) / ((uint64_t)o.block_mb << 20));
    uint64_t cap_blocks = std::max<uint64_t>(1, (uint64_t)(140.0 * 50 / std::max(1, o.block_mb)));
    uint64_t share = (est_blocks + ndev - 1) / ndev, rounds = (share + cap_blocks - 1) / cap_blocks;
    int batch = o.batch > 0 ? o.batch : (int)((share + rounds - 1) / rounds);       // equal batches: 274 blocks on one GPU = 2 x 137
    auto t0 = std::chrono::steady_clock::now();

    // reader: cuts block texts sequentially, a pool parses them into MemBufs, batches go to whichever GPU is free
    std::mutex mu;
    std::condition_variable cv;
    std::deque<EncBatch> ready;                        // parsed, waiting for a GPU
    std::map<uint64_t, EncBatch> done;                 // encoded, waiting for the writer
    bool eof = false;
    int inflight = 0;                                  // batches a GPU thread holds
    std::string failure;
    const size_t max_ready = (size_t)ndev + 1;
    std::thread reader([&]() {
        // One thread cuts (the cut is sequential: every block starts with the previous one's leftover), a pool parses.
        // The text buffers cycle through a small free list: a fresh 25-MiB buffer costs its page faults on this one
        // thread (measured: 1.3 GB/s into fresh memory against 4 GB/s into a buffer that has been touched before).
        const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), 32u));
        std::mutex pmu;
        std::condition_variable pcv;
        struct Task { BlockReader::Raw *raw; MemBuf *dst; };
        std::deque<Task> tasks;
        std::vector<std::unique_ptr<BlockReader::Raw>> store;
        std::vector<BlockReader::Raw *> free_raws;
        for (unsigned k = 0; k < nt + 2; k++) { store.emplace_back(new BlockReader::Raw()); free_raws.push_back(store.back().get()); }
        size_t pending = 0;
        bool stop = false;
        std::string perr;
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; t++)
            pool.emplace_back([&]() {
                for (;;) {
                    Task tk;
                    {
                        std::unique_lock<std::mutex> lk(pmu);
                        pcv.wait(lk, [&]() { return !tasks.empty() || stop; });
                        if (tasks.empty()) return;
                        tk = tasks.front();
                        tasks.pop_front();
                    }
                    std::string e;
                    try { BlockReader::parse_raw(*tk.raw, pe, *tk.dst); }
                    catch (const std::exception &ex) { e = ex.what(); }
                    {
                        std::lock_guard<std::mutex> g(pmu);
                        if (!e.empty() && perr.empty()) perr = e;
                        free_raws.push_back(tk.raw);
                        pending--;
                    }
                    pcv.notify_all();
                }
            });
        uint64_t seq = 0;
        try {
            bool more = true;
            while (more) {
                EncBatch eb;
                eb.seq = seq++;
                eb.mbs.resize(batch);
                int got = 0;
                for (; got < batch; got++) {
                    BlockReader::Raw *r;
                    {
                        std::unique_lock<std::mutex> lk(pmu);
                        pcv.wait(lk, [&]() { return !free_raws.empty(); });
                        r = free_raws.back();
                        free_raws.pop_back();
                    }
                    if (!rd.next_raw(*r)) {
                        std::lock_guard<std::mutex> g(pmu);
                        free_raws.push_back(r);
                        more = false;
                        break;
                    }
                    {
                        std::lock_guard<std::mutex> g(pmu);
                        tasks.push_back({r, &eb.mbs[got]});
                        pending++;
                    }
                    pcv.notify_all();
                }
                {
                    std::unique_lock<std::mutex> lk(pmu);
                    pcv.wait(lk, [&]() { return pending == 0; });
                    if (!perr.empty()) throw std::runtime_error(perr);
                }
                if (got == 0) break;
                eb.mbs.resize(got);
                for (auto &m : eb.mbs)
                    if (m.comment != ap.comment) throw std::runtime_error("'+' lines must be all bare or all a copy of the ID");
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&]() { return ready.size() < max_ready || !failure.empty(); });
                    if (!failure.empty()) break;
                    ready.push_back(std::move(eb));
                }
                cv.notify_all();
            }
        } catch (const std::exception &e) {
            std::lock_guard<std::mutex> g(mu);
            if (failure.empty()) failure = e.what();
        }
        { std::lock_guard<std::mutex> g(pmu); stop = true; }
        pcv.notify_all();
        for (auto &th : pool) th.join();
        { std::lock_guard<std::mutex> g(mu); eof = true; }
        cv.notify_all();
    });
    std::vector<std::thread> gpus;
    for (int d = 0; d < ndev; d++)
        gpus.emplace_back([&, d]() {
            for (;;) {
                EncBatch eb;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&]() { return !ready.empty() || eof || !failure.empty(); });
                    if (!failure.empty() || ready.empty()) return;
                    eb = std::move(ready.front());
                    ready.pop_front();
                    inflight++;
                }
                cv.notify_all();
                int n = (int)eb.mbs.size();
                std::vector<sq_block_in> in(n);
                std::vector<sq_block_out> out(n);
                eb.recs.resize(n);
                eb.lens.resize(n);
                // record buffers: 0.6 x the block's payload covers every FASTQ this coder compresses at all (the bench data
                // needs 0.30); a block that does not fit comes back SQ_ERR_CAPACITY and is coded again with the worst case
                std::vector<std::pair<uint8_t *, size_t>> touch;
                for (int i = 0; i < n; i++) {
                    in[i] = as_in(eb.mbs[i]);
                    size_t cap = (size_t)(0.6 * (double)(eb.mbs[i].ids.size() + 2 * eb.mbs[i].seq.size())) + 8 * (size_t)eb.mbs[i].nreads() + 65536;
                    eb.recs[i].reset(new uint8_t[cap]);
                    out[i] = {eb.recs[i].get(), cap, 0, 0};
                    touch.push_back({eb.recs[i].get(), cap});
                }
                prefault(touch);
                int rc = sq_encode_blocks(ctx[d], in.data(), n, out.data());
                if (rc == SQ_ERR_CAPACITY) {
                    for (int i = 0; i < n && rc == SQ_ERR_CAPACITY; i++) {
                        if (out[i].status != SQ_ERR_CAPACITY) continue;
                        size_t cap = eb.mbs[i].ids.size() + 2 * eb.mbs[i].seq.size() + 8 * (size_t)eb.mbs[i].nreads() + 65536;
                        eb.recs[i].reset(new uint8_t[cap]);
                        out[i] = {eb.recs[i].get(), cap, 0, 0};
                        int r1 = sq_encode_blocks(ctx[d], &in[i], 1, &out[i]);
                        if (r1 != SQ_OK) rc = r1;
                    }
                    if (rc == SQ_ERR_CAPACITY) rc = SQ_OK;
                    for (int i = 0; i < n && rc == SQ_OK; i++) if (out[i].status != SQ_OK) rc = out[i].status;
                }
                if (rc != SQ_OK) {
                    std::lock_guard<std::mutex> g(mu);
                    if (failure.empty()) failure = std::string("encode: ") + sq_last_error(ctx[d]) + " (status " + std::to_string(rc) + ")";
                    inflight--;
                    cv.notify_all();
                    return;
                }
                // rule-based IDs: what the decoder will regenerate from each block's first ID must be what went in
                if (ap.idbin[0]) {
                    for (int i = 0; i < n && rc == SQ_OK; i++) {
                        const MemBuf &m = eb.mbs[i];
                        std::vector<uint8_t> pay(2 + (m.nreads() ? m.idlen[0] : 0)), ids;
                        std::vector<uint16_t> idl;
                        if (m.nreads()) { pay[0] = (uint8_t)m.idlen[0]; pay[1] = (uint8_t)(m.idlen[0] >> 8); memcpy(pay.data() + 2, m.ids.data(), m.idlen[0]); }
                        if (!decode_ids(pay.data(), pay.size(), ap.idbin, m.seqlen.data(), m.nreads(), ids, idl) || ids != m.ids || idl != m.idlen) {
                            std::lock_guard<std::mutex> g(mu);
                            if (failure.empty()) failure = "read IDs are not reproduced by the ID rules found in the head of the input (block " + std::to_string(eb.seq * batch + i) + ")";
                            rc = -1;
                        }
                    }
                    if (rc != SQ_OK) { { std::lock_guard<std::mutex> g(mu); inflight--; } cv.notify_all(); return; }
                }
                for (int i = 0; i < n; i++) {
                    eb.lens[i] = out[i].len;
                    MemBuf &m = eb.mbs[i];          // only the index data is still needed
                    std::vector<uint8_t>().swap(m.ids); std::vector<uint8_t>().swap(m.seq); std::vector<uint8_t>().swap(m.qual);
                }
                {
                    std::lock_guard<std::mutex> g(mu);
                    uint64_t k = eb.seq;
                    done.emplace(k, std::move(eb));
                    inflight--;
                }
                cv.notify_all();
            }
        });
    // writer (this thread): the ordered gather
    uint64_t nblocks = 0, in_bytes = 0, want = 0;
    for (;;) {
        EncBatch eb;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&]() { return done.count(want) || !failure.empty() || (eof && ready.empty() && inflight == 0); });
            if (!failure.empty()) break;
            auto it = done.find(want);
            if (it == done.end()) break;               // input exhausted and every batch written
            eb = std::move(it->second);
            done.erase(it);
        }
        for (size_t i = 0; i < eb.mbs.size(); i++) {
            wr.add_block(eb.recs[i].get(), eb.lens[i], eb.mbs[i].src_bytes, eb.mbs[i].src_off, eb.mbs[i].longflag);
            in_bytes += eb.mbs[i].src_bytes[0] + eb.mbs[i].src_bytes[1];
        }
        nblocks += eb.mbs.size();
        want++;
    }
    cv.notify_all();
    reader.join();
    for (auto &t : gpus) if (t.joinable()) t.join();
    if (!failure.empty()) die(failure);
    if (!wr.finish(ap)) die("write failed");
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (o.verbose)
        fprintf(stderr, "blocks %llu  in %llu B  out %llu B  ratio %.3f  usemodel %d  gpus %d  batch %d  %.2f s\n", (unsigned long long)nblocks,
                (unsigned long long)in_bytes, (unsigned long long)wr.bytes(), wr.bytes() ? (double)in_bytes / wr.bytes() : 0.0, (int)usemodel, ndev, batch, s);
    for (auto c : ctx) sq_destroy(c);
    return 0;
}

// tag-5 payload of a block record: [16-byte MD5] + u16le length + first ID (compressID @0x4247c0)
static bool find_id_payload(const uint8_t *rec, uint64_t len, bool md5, const uint8_t **pay, size_t *plen, const uint8_t **dig) {
    auto vint = [](const uint8_t *p, const uint8_t *end, uint64_t *v) {
        if (p >= end) return 0;
        int n = 1;
        uint8_t m = 0x80;
        while (n <= 8 && !(p[0] & m)) { n++; m >>= 1; }
        if (n > 8 || p + n > end) return 0;
        uint64_t x = p[0] & (m - 1);
        for (int i = 1; i < n; i++) x = (x << 8) | p[i];
        *v = x;
        return n;
    };
    const uint8_t *p = rec, *end = rec + len;
    uint64_t tag, sz;
    int k = vint(p, end, &tag); if (!k) return false; p += k;
    k = vint(p, end, &sz); if (!k) return false; p += k;
    while (p < end) {
        k = vint(p, end, &tag); if (!k) return false; p += k;
        k = vint(p, end, &sz); if (!k || p + sz > end) return false; p += k;
        if (tag == 5) {
            *dig = nullptr;
            if (md5) { if (sz < 16) return false; *dig = p; p += 16; sz -= 16; }
            *pay = p; *plen = sz;
            return true;
        }
        p += sz;
    }
    return false;
}

// Decompress: the block index gives every block's place in the output files (source offsets), so the GPUs take
// contiguous block ranges (ranges keep a GPU's records adjacent in the mapped archive) and write where the index says.
static int do_decompress(const Opt &o) {
    ArcReader ar;
    if (!ar.load(o.arc)) die(ar.err);
    const bool rule = ar.p.idbin[0] != 0;
    bool pe = ar.p.pe;
    // without a prefix the outputs take the names of the compressed inputs (README.md:97; measured: SeqArc -d x.arc
    // writes a_1.fq / a_2.fq into the current directory)
    std::string n1, n2;
    if (o.out.empty()) { n1 = ar.p.name1; n2 = ar.p.name2; if (n1.empty() || (pe && n2.empty())) die("the archive holds no input names; give an output prefix"); }
    else { n1 = pe ? o.out + "_1.fastq" : o.out + ".fastq"; n2 = o.out + "_2.fastq"; }
    if (!o.force && (exists(n1) || (pe && exists(n2)))) die("output exists (use -f)");
    const std::vector<int> devs0 = pick_devices(o);
    const size_t nb = ar.idx.size();
    const int ndev = (int)std::max<size_t>(1, std::min<size_t>(devs0.size(), nb));
    std::vector<int> devs(devs0.begin(), devs0.begin() + ndev);
    int fd1 = ::open(n1.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644), fd2 = pe ? ::open(n2.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644) : -1;
    if (fd1 < 0 || (pe && fd2 < 0)) die("cannot create output");
    std::mutex mu;
    std::string failure;
    auto fail = [&](const std::string &m) { std::lock_guard<std::mutex> g(mu); if (failure.empty()) failure = m; };
    auto pwrite_all = [&](int fd, const std::string &t, uint64_t off) {
        size_t done = 0;
        while (done < t.size()) {
            ssize_t k = ::pwrite(fd, t.data() + done, t.size() - done, (off_t)(off + done));
            if (k <= 0) { fail("write failed"); return; }
            done += (size_t)k;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> gpus;
    for (int d = 0; d < ndev; d++)
        gpus.emplace_back([&, d]() {
            size_t lo = nb * d / ndev, hi = nb * (d + 1) / ndev;      // this GPU's contiguous block range
            if (lo >= hi) return;
            sq_ctx *c = make_ctx(o, devs[d], ar.p.usemodel, ar.p.lossy, 1.5, ar.p.md5, rule);
            if (ar.p.usemodel && sq_load_model(c, ar.p.model.data(), ar.p.model.size()) != SQ_OK) { fail(sq_last_error(c)); sq_destroy(c); return; }
            // blocks per GPU call: the chain kernels take the same time for one block as for a whole wave, so as many as
            // the decoder's workspace is sure to hold (~0.17 GB per default-size block with the host buffers)
            size_t batch = o.batch > 0 ? (size_t)o.batch : std::max<size_t>(1, (size_t)(600.0 * 50 / std::max(1, o.block_mb)));
            std::thread writer;                        // renders / writes batch k while the GPU decodes batch k + 1
            for (size_t b0 = lo; b0 < hi && failure.empty(); b0 += batch) {
                int n = (int)std::min<size_t>(batch, hi - b0);
                // output buffers of the batch: uninitialised storage (a zero-filled std::vector would touch every byte twice)
                struct Out {
                    std::unique_ptr<uint8_t[]> ids, seq, qual;
                    std::unique_ptr<uint16_t[]> idlen;
                    std::unique_ptr<int32_t[]> seqlen;
                };
                auto recs = std::make_shared<std::vector<const uint8_t *>>(n);
                auto lens = std::make_shared<std::vector<uint64_t>>(n);
                auto dec = std::make_shared<std::vector<sq_block_dec>>(n);
                auto bufs = std::make_shared<std::vector<Out>>(n);
                for (int i = 0; i < n; i++) {
                    const BlockIndex &bi = ar.idx[b0 + i];
                    (*recs)[i] = ar.base + bi.arc_off;
                    (*lens)[i] = bi.len;
                    uint64_t src = (uint64_t)bi.src_bytes[0] + bi.src_bytes[1];
                    uint64_t idcap = src + 64, scap = src / 2 + 64;           // a record is "@id\n seq\n +\n qual\n": ids < src, seq = qual < src / 2
                    uint32_t rcap = (uint32_t)(src / 6 + 16);
                    Out &u = (*bufs)[i];
                    u.ids.reset(new uint8_t[idcap]); u.seq.reset(new uint8_t[scap]); u.qual.reset(new uint8_t[scap]);
                    u.idlen.reset(new uint16_t[rcap]); u.seqlen.reset(new int32_t[rcap]);
                    (*dec)[i] = {u.ids.get(), idcap, u.seq.get(), u.qual.get(), scap, u.idlen.get(), u.seqlen.get(), rcap, 0, 0, 0, 0,
                                 bi.longflag ? 1 : 0};
                }
                {
                    std::vector<std::pair<uint8_t *, size_t>> touch;
                    for (int i = 0; i < n; i++) {
                        const sq_block_dec &dd = (*dec)[i];
                        // (only the part of the ID buffer a block can fill matters; the rest is never touched)
                        touch.push_back({dd.ids, (size_t)std::min<uint64_t>(dd.idcap, dd.seqcap)});
                        touch.push_back({dd.seq, (size_t)dd.seqcap});
                        touch.push_back({dd.qual, (size_t)dd.seqcap});
                    }
                    prefault(touch);
                }
                int rc = sq_decode_blocks(c, recs->data(), lens->data(), n, dec->data());
                if (rc != SQ_OK) {
                    const char *w = rc == SQ_ERR_MD5_NAME ? "Name md5 unequal" : rc == SQ_ERR_MD5_SEQ ? "Seq md5 unequal" : rc == SQ_ERR_MD5_QUAL ? "Qual md5 unequal" : sq_last_error(c);
                    if (rc == SQ_ERR_FORMAT && ar.p.has_ref)
                        w = "a block holds reads mapped to the reference the archive was written against; this decoder does not rebuild them yet";
                    fail(std::string("decode: ") + w);
                    break;
                }
                if (writer.joinable()) writer.join();
                writer = std::thread([&, recs, lens, dec, bufs, b0, n]() {
                    // regenerate rule-based IDs, render the FASTQ text of the batch's blocks in parallel, write at the recorded offsets
                    std::atomic<int> next{0};
                    auto work = [&]() {
                        std::string t1, t2;
                        std::vector<uint8_t> rids;
                        std::vector<uint16_t> ridl;
                        for (int i; (i = next.fetch_add(1)) < n;) {
                            Out &u = (*bufs)[i];
                            const sq_block_dec &dd = (*dec)[i];
                            const uint8_t *ids = u.ids.get();
                            const uint16_t *idl = u.idlen.get();
                            const char *bad = nullptr;
                            if (rule) {               // IDProcess::decodeIDS @0x430610: regenerate the IDs from the block's first ID
                                const uint8_t *pay = nullptr, *md5 = nullptr;
                                size_t plen = 0;
                                if (!find_id_payload((*recs)[i], (*lens)[i], ar.p.md5, &pay, &plen, &md5)) bad = "decode: block without an ID stream";
                                else {
                                    rids.clear(); ridl.clear();
                                    if (!decode_ids(pay, plen, ar.p.idbin, u.seqlen.get(), dd.nreads, rids, ridl)) bad = "decode: bad rule-based ID record";
                                    else if (md5) {
                                        uint8_t dig[16];
                                        md5_digest(rids.data(), rids.size(), dig);
                                        if (memcmp(dig, md5, 16) != 0) bad = "decode: Name md5 unequal";
                                    }
                                    ids = rids.data(); idl = ridl.data();
                                }
                            }
                            if (bad) { fail(bad); continue; }
                            render_raw(ids, idl, u.seq.get(), u.qual.get(), u.seqlen.get(), dd.nreads, ar.p.comment, pe, t1, t2);
                            const BlockIndex &bi = ar.idx[b0 + i];
                            if (t1.size() != bi.src_bytes[0] || (pe && t2.size() != bi.src_bytes[1])) { fail("decode len is not equal org len"); continue; }
                            pwrite_all(fd1, t1, bi.src_off[0]);
                            if (pe) pwrite_all(fd2, t2, bi.src_off[1]);
                            u = Out();
                        }
                    };
                    unsigned nt = std::max(1u, std::min<unsigned>({std::thread::hardware_concurrency(), 32u, (unsigned)n}));
                    std::vector<std::thread> pool;
                    for (unsigned t = 0; t < nt; t++) pool.emplace_back(work);
                    for (auto &th : pool) th.join();
                });
            }
            if (writer.joinable()) writer.join();
            sq_destroy(c);
        });
    for (auto &t : gpus) t.join();
    ::close(fd1);
    if (fd2 >= 0) ::close(fd2);
    if (!failure.empty()) die(failure);
    if (o.verbose)
        fprintf(stderr, "blocks %zu  gpus %d  %.2f s\n", nb, ndev, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    return 0;
}

int main(int argc, char **argv) {
    Opt o;
    read_config(o);
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) die("missing value after " + a); return argv[++i]; };
        if (a == "-c") o.compress = true;
        else if (a == "-d") o.decompress = true;
        else if (a == "-f") o.force = true;
        else if (a == "-n") o.md5 = false;
        else if (a == "-v") o.verbose = true;
        else if (a == "-t") val();                      // worker threads: the GPU takes whole batches instead
        else if (a == "-1") o.in1 = val();
        else if (a == "-2") o.in2 = val();
        else if (a == "-o") o.out = val();
        else if (a == "-l") { o.lossy = true; o.lossy_r = atof(val().c_str()); }
        else if (a == "-g") {                           // -g 0,2,3: the GPUs to shard the job over (default: all)
            std::string v = val();
            for (size_t p0 = 0; p0 <= v.size();) { size_t p1 = v.find(',', p0); if (p1 == std::string::npos) p1 = v.size(); if (p1 > p0) o.devices.push_back(atoi(v.substr(p0, p1 - p0).c_str())); p0 = p1 + 1; }
        }
        else if (a == "--usemodel") o.force_usemodel = atoi(val().c_str());
        else if (a == "--parse-only") o.parse_only = true;
        else if (a == "-b") o.batch = atoi(val().c_str());
        else if (a == "-h") { printf("usage: seqarc_b200 -c [ref.fa] -1 a.fq [-2 b.fq] -o out [-l r] [-n] [-f]   (or: -c [ref.fa] a.fq [b.fq] out)\n       seqarc_b200 -d [ref.fa] in.arc -o out [-f]                            (or: -d [ref.fa] in.arc out)\n"); return 0; }
        else if (a[0] != '-') pos.push_back(a);
        else die("unknown option " + a);
    }
    // positional forms of the reference CLI:
    //   -c [options] [ref.fa] <input_file> [input_file2] <compress_prefix>      (or -1 / -2 / -o)
    //   -d [options] [ref.fa] <x.arc> [fastq_prefix]
    auto ends_with = [](const std::string &x, const char *suf) { size_t n = strlen(suf); return x.size() >= n && x.compare(x.size() - n, n, suf) == 0; };
    auto is_fasta = [&](const std::string &x) { return ends_with(x, ".fa") || ends_with(x, ".fasta") || ends_with(x, ".fna"); };
    if (o.compress) {
        size_t k = 0;
        if (k < pos.size() && is_fasta(pos[k])) o.ref = pos[k++];
        std::vector<std::string> rest(pos.begin() + k, pos.end());
        if (o.in1.empty()) {                       // inputs (and the prefix, unless -o) are positional
            size_t need_out = o.out.empty() ? 1 : 0;
            if (rest.size() == 1 + need_out) o.in1 = rest[0];
            else if (rest.size() == 2 + need_out) { o.in1 = rest[0]; o.in2 = rest[1]; }
            else if (!rest.empty()) die("cannot tell inputs from the output prefix (see -h)");
            if (need_out && !rest.empty()) o.out = rest.back();
        } else if (rest.size() == 1 && o.out.empty()) o.out = rest[0];
        else if (!rest.empty()) die("unexpected argument " + rest[0]);
    } else if (o.decompress) {
        size_t k = 0;
        if (k + 1 < pos.size() && !ends_with(pos[k], ".arc")) o.ref = pos[k++];
        if (k < pos.size()) o.arc = pos[k++];
        if (k < pos.size() && o.out.empty()) o.out = pos[k++];
        if (k < pos.size()) die("unexpected argument " + pos[k]);
    }
    if (o.parse_only) {                            // argument handling is testable without a GPU
        printf("mode=%s ref=%s in1=%s in2=%s arc=%s out=%s lossy=%d md5=%d\n", o.compress ? "c" : o.decompress ? "d" : "-", o.ref.c_str(),
               o.in1.c_str(), o.in2.c_str(), o.arc.c_str(), o.out.c_str(), (int)o.lossy, (int)o.md5);
        return 0;
    }
    try {
        if (o.compress && !o.in1.empty() && !o.out.empty()) return do_compress(o);
        if (o.decompress && !o.arc.empty()) return do_decompress(o);
    } catch (const std::exception &e) {
        die(e.what());
    }
    die("nothing to do (see -h)");
}
