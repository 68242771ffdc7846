/*
 * seqarc_b200.h -- C ABI of the B200-native SeqArc v1.6 entropy-coding path.
 *
 * The reference (BGI-flexlab/fastqueeze, shipped as the unstripped binary
 * /root/reference/SeqArc-1.6) has no plugin/FFI boundary.  The seam this
 * library replaces is the per-worker coder object:
 *
 *   int EncapFqzComp::doFqzEncode(SeqArcMemBuf*, DebugInfo*)   @0x42d2d0  (called from doTask @0x433e9a)
 *   int EncapFqzComp::doDecode  (SeqArcMemBuf*, DebugInfo*)    @0x42eb60  (called from decodeData)
 *
 * One sq_ctx stands where one EncapFqzComp (+ its SeqArcMemBuf) stands, except that it takes a
 * BATCH of independent blocks (every model resets per block) and runs them on one
 * GPU.  Plain pointers and sizes only; no torch / CUDA types in any signature.
 *
 * All functions return SQ_OK (0) or a negative sq_status; nothing aborts (the reference abort()s
 * on a coder invariant @0x4220b8 and exit(1)s elsewhere).  There is NO CPU fallback: without a
 * usable CUDA device sq_create fails with SQ_ERR_NO_DEVICE.
 */
#ifndef SEQARC_B200_H
#define SEQARC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SQ_OK = 0,
    SQ_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product path never falls back to the CPU */
    SQ_ERR_CUDA = -2,        /* a CUDA call failed; sq_last_error() has the text */
    SQ_ERR_ARG = -3,         /* bad argument (null pointer, n <= 0, unsupported option) */
    SQ_ERR_CAPACITY = -4,    /* an output buffer is too small */
    SQ_ERR_INPUT = -5,       /* input the reference coder cannot represent (ID > 255 B, quality > '~', ...) */
    SQ_ERR_FORMAT = -6,      /* malformed block record on decode */
    SQ_ERR_NOMEM = -7,
    SQ_ERR_MD5_NAME = -8,    /* blockMd5Verify @0x414e00: "Name md5 unequal" */
    SQ_ERR_MD5_SEQ = -9,
    SQ_ERR_MD5_QUAL = -10
} sq_status;

/* The SeqArcParam fields the hot path reads ("Config / flags"). */
typedef struct {
    int32_t device;        /* CUDA ordinal this context is bound to */
    int32_t usemodel;      /* SeqArcParam+0x8: start every block from the snapshot given to sq_load_model */
    int32_t lossy;         /* SeqArcParam+0x1870 (-l) */
    double  lossy_r;       /* SeqArcParam+0x1878 */
    int32_t md5;           /* SeqArcParam+0x1880: embed / skip the 16-byte per-stream MD5s (hidden -n) */
    int32_t slevel;        /* SeqArcParam+0x1b54, default 3 -> order-10 base context */
    int32_t qlevel;        /* SeqArcParam+0x1b58, default 2 */
    int32_t id_rule;       /* SeqArcParam+0x18a4 idbin[0]: 1 = rule-based ID path (first ID only) */
    uint64_t workspace_bytes;  /* device workspace budget; 0 = 80 % of free device memory */
    void   *stream;        /* cudaStream_t to launch on, or NULL for a private stream */
    int32_t ref_mode;      /* 0: no reference index.  1 (SE) / 2 (PE): the archive is written against a reference
                              index and no read of the block mapped (BASELINE configs[4]): doAlignEncode @0x42d4c0
                              frames the same ID / quality / base streams plus empty alignment streams (tags 27, 8,
                              17-22, 9, 11, 15, 10, 12, 13).  Blocks holding mapped reads are out of scope: the
                              decoder answers SQ_ERR_FORMAT for them */
} sq_params;

/* One block as SeqArcMemBuf holds it at the doFqzEncode seam: concatenated buffers
 * without separators, PE interleaved R1,R2,R1,R2 (getBlockReadPE @0x412920). */
typedef struct {
    uint32_t nreads;          /* membuf+0x4  */
    int32_t  longflag;        /* membuf+0x2: set when the block holds a read longer than 65,535 bases; the Len stream is then
                                 written by compressLen_long @0x423710 and the block-index length field carries bit 0 */
    const uint8_t  *ids;      /* membuf+0x80: IDs without '@' */
    const uint8_t  *seq;      /* membuf+0x88 */
    const uint8_t  *qual;     /* membuf+0x90 */
    const uint16_t *idlen;    /* membuf+0xc0 */
    const int32_t  *seqlen;   /* membuf+0xd0 */
    uint64_t idbytes, seqbytes;
} sq_block_in;

/* Encoded block record "tag 1 | size(4) | streams" (SURVEY App. A), as written at membuf+0x70. */
typedef struct {
    uint8_t *data;            /* caller buffer (host for sq_encode_blocks, device for *_dev) */
    uint64_t cap;
    uint64_t len;             /* out: bytes of the record (= return value of doFqzEncode) */
    int32_t  status;          /* out: per-block sq_status */
} sq_block_out;

/* Decoded block: caller-provided buffers with capacities; the decoder fills counts and lengths. */
typedef struct {
    uint8_t  *ids;   uint64_t idcap;
    uint8_t  *seq;   uint8_t *qual;  uint64_t seqcap;
    uint16_t *idlen; int32_t *seqlen; uint32_t readcap;
    uint32_t nreads;          /* out */
    uint64_t idbytes, seqbytes;   /* out */
    int32_t  status;          /* out */
    int32_t  longflag;        /* in: bit 0 of the block's length field in the archive index (decode_len_long @0x427190) */
} sq_block_dec;

typedef struct sq_ctx sq_ctx;

/* EncapFqzComp::EncapFqzComp @0x42f2a0 + initModel @0x42ed00: bind to a device, size the workspace. */
int  sq_create(const sq_params *p, sq_ctx **out);
/* EncapFqzComp::~EncapFqzComp */
void sq_destroy(sq_ctx *c);
/* text of the last error on this context (or of the last failed sq_create when c == NULL) */
const char *sq_last_error(const sq_ctx *c);

/* EncapFqzComp::ReadModelFormMem @0x420fc0: install the model snapshot (trailer tag 6 payload
 * after its "tag 1, u16 = 1" header; SaveModelToMem layout @0x420c00) used when usemodel = 1. */
int  sq_load_model(sq_ctx *c, const uint8_t *blob, uint64_t len);
/* bytes such a snapshot must have for this context's slevel / qlevel / id_rule */
uint64_t sq_model_size(const sq_ctx *c);

/* SeqArcPreProcess::doPreProcess @0x4153b0 -> encodeForModel @0x4228e0 -> SaveModelToMem @0x420c00:
 * train the model snapshot on the sample blocks (HOST buffers; models are not reset between them)
 * and write it to blob[0 .. sq_model_size).  The caller stores it in the .arc trailer and installs it
 * with sq_load_model. */
int  sq_train_model(sq_ctx *c, const sq_block_in *blocks, int n, uint8_t *blob, uint64_t cap);

/* doFqzEncode @0x42d2d0 over a batch: HOST buffers in, HOST records out, order preserved.
 * Copies host->device, runs the kernels, copies the records device->host. */
int  sq_encode_blocks(sq_ctx *c, const sq_block_in *blocks, int n, sq_block_out *out);
/* same with every pointer inside blocks[] / out[] being a DEVICE pointer on this context's GPU
 * (16-byte aligned); no host<->device payload copies are made. */
int  sq_encode_blocks_dev(sq_ctx *c, const sq_block_in *blocks, int n, sq_block_out *out);

/* doDecode @0x42eb60 over a batch of block records (each starting at its tag-1 byte). */
int  sq_decode_blocks(sq_ctx *c, const uint8_t *const *recs, const uint64_t *reclens, int n,
                      sq_block_dec *out);
int  sq_decode_blocks_dev(sq_ctx *c, const uint8_t *const *recs, const uint64_t *reclens, int n,
                          sq_block_dec *out);

/* EncapFqzComp::rblock @0x426c10 on its own (R-Block smoothing of one block's concatenated
 * quality bytes, in place, HOST buffer). */
int  sq_rblock(sq_ctx *c, uint8_t *qual, uint64_t n, double r);

/* Aligned-mode side streams (doAlignEncode @0x42d4c0 / doAlignDecode @0x42c8d0): a block that holds mapped reads carries
 * the aligner's arrays as small streams, each the bare range-coder output of a byte array through a fresh SIMPLE_MODEL<N>
 * (no count in front; the counts travel in their own tags).  These two calls stand where
 *   compressOrder @0x424b70 (tag 8, N = 5), compressPERelation @0x422be0 (tag 9, N = 4),
 *   compressAlignInfo_Pos @0x425d70 (11, N = 2), _Mis @0x425ff0 (15, N = 8 or 9), _Rev @0x426480 (10, N = 2),
 *   _CigaL @0x426700 (12, N = 2), _CigaV @0x426980 (13, N = 4)
 * and their decompress* twins (@0x42b130 ... @0x42c3b0) stand in the reference.  All streams of a call are coded at once,
 * one chain each.  Buffers are HOST memory.  Alignment itself and the block framing of aligned blocks are not part of
 * this library yet: sq_decode_blocks answers SQ_ERR_FORMAT for a block with mapped reads. */
typedef struct sq_side_stream {
    int32_t  model;          /* N of SIMPLE_MODEL<N>: 2, 4, 5, 8 or 9 */
    uint32_t nsym;           /* symbols in the array */
    uint8_t *sym;            /* encode: in (every value < model); decode: out, nsym bytes */
    uint8_t *bytes;          /* encode: out, `cap` bytes available (nsym + 16 always suffices); decode: in */
    uint64_t cap;            /* encode: capacity of `bytes`; decode: length of the stream */
    uint64_t len;            /* encode: out, bytes written */
    int32_t  status;         /* out: SQ_OK, SQ_ERR_CAPACITY, SQ_ERR_INPUT (a symbol >= model) */
} sq_side_stream;
int  sq_side_encode(sq_ctx *c, sq_side_stream *streams, int n);
int  sq_side_decode(sq_ctx *c, sq_side_stream *streams, int n);

/* Aligned mode, decode side: the bases of the MAPPED reads of a block, rebuilt from the reference (getRealPos @0x42de30 +
 * AlignInfoToSeq @0x42e020 without its degenerate-base overlay).  sq_set_reference uploads the concatenated reference
 * sequence (ASCII bases, header lines removed) once; it stays in HBM until sq_destroy.  sq_align_rebuild takes the decoded
 * side arrays of each block (HOST memory; Pos and CigaL are bit arrays, one bit per byte) and fills `seq` at the offsets of
 * the mapped reads; the bytes of unmapped reads are left as they were. */
typedef struct sq_align_block {
    uint32_t nreads;           /* reads of the block (paired-end: interleaved R1, R2, ...) */
    int32_t  pe;               /* 1: paired-end (PERelation / insbits used for second mates) */
    const int32_t *seqlen;     /* nreads */
    const uint8_t *order;      /* nreads: 0 = not mapped, else 1 + the high bits of the position */
    const uint8_t *pos;        /* npos bits */
    const uint8_t *mis;        /* nalign: mismatches of each mapped read (at most 16) */
    const uint8_t *rev;        /* nalign: strand */
    const uint8_t *cigal;      /* ncigal bits: mismatch offsets */
    const uint8_t *cigav;      /* ncigav: which other base each mismatch is */
    const uint8_t *perel;      /* nperel: relation of a second mate to the first (pairs mapped with both mates) */
    uint32_t npos, nalign, ncigal, ncigav, nperel, insbits;
    uint8_t *seq;              /* in / out: the block's concatenated bases */
    uint64_t seqbytes;
    int32_t  status;           /* out: SQ_OK or SQ_ERR_FORMAT (arrays inconsistent, a position outside the reference) */
} sq_align_block;
int  sq_set_reference(sq_ctx *c, const uint8_t *bases, uint64_t n);
int  sq_align_rebuild(sq_ctx *c, sq_align_block *blocks, int n);

/* Measurement hooks (DebugInfo @0x420080 stands here in the reference): device time of the
 * kernels of the last encode / decode call, per stage, in milliseconds (CUDA events on the
 * context's stream).  names[i] are static strings; returns the number of stages filled. */
int  sq_last_timings(const sq_ctx *c, const char **names, float *ms, int cap);
/* number of kernel launches issued by the last encode / decode call */
uint64_t sq_last_launches(const sq_ctx *c);
/* how many blocks one wave of the last call held (workspace-limited) */
int  sq_last_wave_blocks(const sq_ctx *c);

/* Test hook: copy an intermediate device buffer of the LAST encoded wave to the host.
 * what: 0 = seq triples (u64 cum|freq<<16|tot<<32 per coded base, stream order), 1 = qual triples,
 * 2 = name triples, 3 = number of coded symbols per stream (3 x u64). block = index in the wave. */
int  sq_debug_fetch(sq_ctx *c, int what, int block, void *dst, uint64_t cap, uint64_t *got);

/* number of CUDA devices this process can use (0 without a driver / device).  The host driver shards a job's blocks
 * over them: one sq_ctx and one host thread per device, block records gathered in block order (the layout
 * rule of addBlockPEInfo @0x412f30: offset = previous offset + previous length). */
int  sq_device_count(void);

const char *sq_version(void);

#ifdef __cplusplus
}
#endif
#endif
