// sq_kernels.h -- interface between the kernel files and the C-ABI layer (sq_api.cu).
#pragma once
#include "sq_common.cuh"

// status values written by kernels (mirror sq_status in include/seqarc_b200.h)
#define SQ_ERR_CAPACITY_K (-4)
#define SQ_ERR_INPUT_K (-5)
#define SQ_ERR_FORMAT_K (-6)

#define SQ_TILE 8192            // sort tile: elements owned by one CTA
#define SQ_SORT_THREADS 512
#define SQ_SORT_BITS 8          // radix digit width
#define SQ_SORT_BITS_WIDE 10    // ... for 17..20-bit keys (two passes)
#define SQ_CHAIN_SMEM_RC (72 * 1024)      // dynamic shared memory requested only to cap chain CTAs per SM (and, with them, the
#define SQ_CHAIN_SMEM_SMALL (36 * 1024)   // front-end CTAs that can sit beside a chain: measured 48 / 0, 32 / 8, 0 / 0 KB: all slower)

// Everything the encode kernels of one wave need (all pointers are device pointers).
struct EncWave {
    int B;                      // blocks in the wave
    u32 R;                      // reads in the wave
    int nsm;                    // SM count (persistent-kernel grid sizing)
    int md5, lossy, id_rule, qlevel;
    int ref_mode;               // 0, or 1 (SE) / 2 (PE): frame the empty alignment streams of doAlignEncode
    u32 seqmask;                // 4^NS - 1
    BlockDesc *bd;
    BlockState *bs;
    u32 *rbase;                 // [B+1]
    u32 *totals;                // [6*B] scan totals
    // per read (wave-wide, indexed rbase[b] + r)
    u32 *seqoff, *idoff, *acgtoff, *qlen2, *qoff, *noff;
    u8 *tip, *np, *ns;
    // per symbol
    u64 maxbound[ST_COUNT];     // max over blocks of ebound (grid sizing)
    u64 *elemA, *elemB;
    u64 *triples[ST_COUNT];     // [ST_SEQ] really holds u32 (tr_make32)
    u32 *hist;
    u32 *scanpart;              // chunk sums of the histogram scan, [B][ceil(D * maxtiles / 4096)]
    u64 *seglist; u32 segcap; u32 *nseg;   // nseg[0] count, nseg[1..2] work cursors
    u8 *obuf;                   // range-coder scratch
    u8 *sidebuf;
    u8 *recbuf; u64 recbuf_cap; u64 *recoff;   // compact record buffer, [B+1] offsets
    const u8 *snap_name, *snap_seq, *snap_qual;   // model snapshot sections (usemodel) or null
    u8 *train_name, *train_seq, *train_qual;      // training (encodeForModel): where the final model states go, or null
};

cudaError_t sqk_enc_attrs();
cudaError_t sqk_prep(const EncWave &w, cudaStream_t st, u64 *launches);
cudaError_t sqk_rblock(const EncWave &w, double r, cudaStream_t st, u64 *launches);
cudaError_t sqk_count(const EncWave &w, cudaStream_t st, u64 *launches);
cudaError_t sqk_emit(const EncWave &w, int stream, u64 *elem, cudaStream_t st, u64 *launches);
cudaError_t sqk_sort(const EncWave &w, int stream, int keybits, u64 *a, u64 *b, u64 **sorted, cudaStream_t st, u64 *launches);
cudaError_t sqk_replay(const EncWave &w, int stream, const u64 *sorted, void *triples, cudaStream_t st, u64 *launches);
cudaError_t sqk_rc(const EncWave &w, int stream, cudaStream_t st, u64 *launches);
cudaError_t sqk_md5(const EncWave &w, cudaStream_t st, u64 *launches);
cudaError_t sqk_side(const EncWave &w, cudaStream_t st, u64 *launches);
cudaError_t sqk_snap_init(u8 *blob, u64 name_bytes, u32 nseq, u32 nqual, cudaStream_t st, u64 *launches);
cudaError_t sqk_pack(const EncWave &w, int compact, cudaStream_t st, u64 *launches);

// ---------------------------------------------------------------------------------------------
// decode side (sq_decode.cu)
// decoder-private "warp model" tables: (EPL + 1) rows of 32 words per model
#define SQ_WQ_STRIDE (4 * 32)                                  /* SIMPLE_MODEL<95>, 3 entries per lane */
#ifndef SQ_QF
#define SQ_QF 16                                               /* decoder: lanes of a quality model read per symbol (front entries), 8 / 16 / 32 */
#endif
#define SQ_WNAME_WORDS (768ull * 9 * 32 + 8192ull * 5 * 32)     /* 768 x <256> (8 per lane) + 8192 x <128> (4 per lane) */
struct DecBlock {
    const u8 *rec; u64 reclen;          // block record (device)
    // payload pointers / sizes of the streams found in the record (0 = absent); MD5 prefixes skipped
    u64 off[32]; u32 len[32];
    u8 md5[3][16]; int has_md5[3];
    // outputs (device)
    u8 *ids, *seq, *qual; u16 *idlen; i32 *seqlen;
    u64 idcap, seqcap; u32 readcap;
    u32 nreads;
    i32 longflag;                       // from the archive's block index: Len stream written by compressLen_long
    // per-block model tables (device workspace)
    u8 *seqtab; u32 *qualtab; u32 *nametab;
    // dege arrays
    u8 *d_tip, *d_ch, *d_maxq; u32 *d_cnt, *d_pos;
    u32 *seqoff;                        // [nreads+1]
    // results
    u64 idbytes, seqbytes;
    i32 status;
};
struct DecWave {
    int B, nsm;
    int md5, lossy, id_rule, qlevel;
    u32 seqmask, nseq, nqual;
    DecBlock *db;
    const u8 *snap_name, *snap_seq, *snap_qual;
};
cudaError_t sqk_decode(const DecWave &w, cudaStream_t st, u64 *launches);
cudaError_t sqk_check_seqsnap(const u8 *snap_seq, u64 nctx, int *d_flag, int *bad, cudaStream_t st);
