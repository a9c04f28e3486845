/*
 * swb200.h — C ABI of libswb200: B200-native, bit-exact Striped-Smith-Waterman
 * realignment (the hot path of stjude/indelPost).
 *
 * Two nested boundaries:
 *
 *  1. The reference's own C interface, kept symbol-for-symbol so that
 *     indelpost/sswpy.pyx's `cdef extern from "ssw.h"` block
 *     (reference indelpost/sswpy.pyx:57-83) links against this library
 *     unchanged:  ssw_init / init_destroy / ssw_align / align_destroy
 *     (reference indelpost/ssw.h:86, 91, 126-134, 139) and the s_align
 *     result record (reference indelpost/ssw.h:55-66).
 *     Each call runs on the GPU (one pair per launch sequence); there is no
 *     CPU implementation behind these symbols.
 *
 *  2. The batched entry point swb_align_batch(): N read x window pairs as
 *     SoA over (pinned) host buffers, de-duplicated read and window tables,
 *     per-pair gap penalties / mask length, fixed-stride results that carry
 *     every s_align field plus a CIGAR arena.  One swb_ctx per GPU; the
 *     call is thread-safe across contexts and never touches Python state,
 *     so it may be made with the GIL released.
 *
 * Plain pointers and sizes only; no torch / Python types.
 */
#ifndef SWB200_H
#define SWB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ */
/* 1. reference-compatible interface (replaces indelpost/ssw.c)        */
/* ------------------------------------------------------------------ */

struct _profile;
typedef struct _profile s_profile;           /* opaque, as in ssw.h:39-40 */

/* identical layout to reference indelpost/ssw.h:55-66 */
typedef struct {
    uint16_t score1;
    uint16_t score2;
    int32_t  ref_begin1;
    int32_t  ref_end1;
    int32_t  read_begin1;
    int32_t  read_end1;
    int32_t  ref_end2;
    uint32_t* cigar;
    int32_t  cigarLen;
    uint16_t flag;
} s_align;

/* replaces ssw.c:787-808.  Like the reference, the profile aliases `read`
 * and `mat` (ssw.c:803-804): both must outlive it. */
s_profile* ssw_init(const int8_t* read, const int32_t readLen, const int8_t* mat,
                    const int32_t n, const int8_t score_size);
/* replaces ssw.c:810-814 */
void init_destroy(s_profile* p);
/* replaces ssw.c:816-920.  Returns NULL (after a message on stderr) in the
 * cases the reference does (ssw.c:848-860) and when no GPU is usable. */
s_align* ssw_align(const s_profile* prof, const int8_t* ref, int32_t refLen,
                   const uint8_t weight_gapO, const uint8_t weight_gapE,
                   const uint8_t flag, const uint16_t filters, const int32_t filterd,
                   const int32_t maskLen);
/* replaces ssw.c:922-925 */
void align_destroy(s_align* a);

/* header-inline helpers of ssw.h:171-190 as real symbols (BAM CIGAR packing:
 * len<<4 | op, MAPSTR "MIDNSHP=X") */
char     swb_cigar_int_to_op(uint32_t cigar_int);
uint32_t swb_cigar_int_to_len(uint32_t cigar_int);
uint32_t swb_to_cigar_int(uint32_t length, char op_letter);

/* ------------------------------------------------------------------ */
/* 2. batched interface                                                */
/* ------------------------------------------------------------------ */

#define SWB_SEQ_CODES 0   /* sequences are already 0..n-1 codes (what ssw_init / ssw_align take) */
#define SWB_SEQ_ASCII 1   /* raw ASCII, encoded on the GPU with sswpy's DNA_BASE_LUT (sswpy.pyx:16-29) */
#define SWB_SEQ_PACKED4 2 /* codes 0..15, two per byte (low nibble first): half the host->device bytes, N (code 4) representable */
#define SWB_SEQ_PACKED2 3 /* codes 0..3, four per byte (bits 0-1 first): a quarter of the bytes, for sequences without N */

/* per-pair status word (replaces the reference's fprintf(stderr) reporting) */
#define SWB_OK              0
#define SWB_ERR_BYTE_ONLY   1   /* score_size==0 and the 8-bit pass overflowed: ssw_align returns NULL (ssw.c:848-852) */
#define SWB_ERR_BAD_INPUT   2   /* empty read/window, index out of range, code >= n */

/* Input: de-duplicated read table and window table + per-pair indices.
 * All pointers are host pointers (pinned memory gives the best transfer
 * rate but is not required).  Offsets are in bytes from the blob start.
 * SWB_SEQ_CODES tables may place entries anywhere in the blob (entries may
 * even share bytes).  SWB_SEQ_ASCII tables are encoded in place on the
 * device, so their entries must be in ascending offset order and must not
 * overlap (off[i] >= off[i-1] + len[i-1]); a table that breaks the rule
 * fails the call with -1.
 * SWB_SEQ_PACKED4 / SWB_SEQ_PACKED2: every entry starts on a byte boundary
 * (off[i] = byte offset of its first byte, len[i] = length in BASES; the
 * entry occupies ceil(len/2) resp. ceil(len/4) bytes) and is unpacked on the
 * device; swb_pack_table() below produces such tables. */
typedef struct {
    int32_t        n_pairs;
    int32_t        n_reads;
    int32_t        n_windows;
    int32_t        seq_encoding;  /* SWB_SEQ_CODES | SWB_SEQ_ASCII */

    const int8_t*  reads;         /* concatenated read table                     */
    const int64_t* read_off;      /* [n_reads]                                   */
    const int32_t* read_len;      /* [n_reads]                                   */
    const int8_t*  windows;       /* concatenated window table                   */
    const int64_t* win_off;       /* [n_windows]                                 */
    const int32_t* win_len;       /* [n_windows]                                 */

    const int32_t* pair_read;     /* [n_pairs] index into the read table         */
    const int32_t* pair_win;      /* [n_pairs] index into the window table       */
    const int32_t* ref_beg;       /* [n_pairs] or NULL (=0): sswpy start_idx, sswpy.pyx:230,263-275 */
    const int32_t* ref_len;       /* [n_pairs] or NULL (= win_len - ref_beg): sswpy search_length    */
    const uint8_t* gap_open;      /* [n_pairs] already narrowed to uint8 like ssw_align's parameter  */
    const uint8_t* gap_ext;       /* [n_pairs]                                   */
    const int32_t* mask_len;      /* [n_pairs] or NULL (= max(15, read_len/2), sswpy.pyx:209-211)    */

    const int8_t*  mat;           /* n*n substitution matrix (sswpy.pyx:306-336 builds the 5x5 one)  */
    int32_t        n;
    int8_t         score_size;    /* 0: 8-bit only, 1: 16-bit only, 2: 8-bit then 16-bit (ssw.c:793-802) */
    uint8_t        flag;          /* ssw_align flag; sswpy always passes 1       */
    uint16_t       filters;
    int32_t        filterd;
} swb_batch;

/* Output record: every field of s_align, the CIGAR as (offset,len) into the
 * arena, and a status.  40 bytes, fixed stride. */
typedef struct {
    uint16_t score1;
    uint16_t score2;
    int32_t  ref_begin1;
    int32_t  ref_end1;
    int32_t  read_begin1;
    int32_t  read_end1;
    int32_t  ref_end2;
    int32_t  cigar_len;    /* 0 <=> s_align.cigar == NULL                         */
    uint16_t flag;         /* s_align.flag: 0 ok, 1 banded_sw failed, 2 path may miss a part (ssw.c:888-891, 911) */
    uint16_t status;       /* SWB_OK / SWB_ERR_*                                  */
    int64_t  cigar_off;    /* index (in uint32 units) into the arena              */
} swb_result;

typedef struct swb_ctx swb_ctx;

/* per-stage device timings of the last swb_compute()/swb_align_batch() call,
 * measured with CUDA events on the context's own stream */
typedef struct {
    float   ms_total;        /* first kernel start -> last kernel end             */
    float   ms_prepare;      /* encode + classify + pack                          */
    float   ms_forward;      /* forward score/end kernels (fast + exact)          */
    float   ms_reverse;      /* reverse start-position kernels                    */
    float   ms_traceback;    /* banded DP + traceback + CIGAR emit                */
    float   ms_h2d;          /* host->device copies (swb_align_batch only)        */
    float   ms_d2h;          /* device->host copies (swb_align_batch only)        */
    int64_t cells_forward;   /* DP cells actually swept by the forward kernels    */
    int64_t cells_reverse;
    int64_t cells_band;
    int64_t n_fast;          /* pairs finished by the DPX fast path               */
    int64_t n_exact;         /* pairs that needed the exact striped emulation     */
    int64_t n_launches;      /* kernels launched                                  */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    float   ms_band_round0;  /* first banded-DP round (all pairs at their initial band width)   */
    float   ms_band_rest;    /* band-doubling rounds                                             */
    float   ms_certify;      /* overflow certificate + exact 8-bit verification of the remainder */
    int32_t band_rounds;
    int32_t n_sw_certified;  /* 8-bit-final pairs with scores past 128+go+ge that the sandwich sweep certified (forward pass) */
    int32_t n_sw_rejected;   /* ... that it sent to the exact striped emulation                                               */
    int32_t n_sw_verified;   /* overflow verifications settled by the sandwich lower bound instead of the exact 8-bit pass    */
} swb_timing;

int         swb_device_count(void);
swb_ctx*    swb_create(int device);                 /* NULL if the device cannot be used */
void        swb_destroy(swb_ctx* ctx);
const char* swb_last_error(const swb_ctx* ctx);     /* ctx may be NULL: last creation error */

/* One-shot: copy in, align, copy out.  `results` has n_pairs entries; CIGARs
 * are written into `cigar_arena` (capacity `cigar_cap` uint32s); *cigar_used
 * receives the number used.  Returns 0 on success, negative on failure
 * (message via swb_last_error).  If the arena is too small the call fails
 * with -2 and *cigar_used holds the required size. */
int swb_align_batch(swb_ctx* ctx, const swb_batch* batch, swb_result* results,
                    uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used);

/* Split form used to time the device-resident path separately:
 * upload (H2D) -> compute (kernels only, may be repeated) -> download (D2H). */
int swb_upload(swb_ctx* ctx, const swb_batch* batch);
int swb_compute(swb_ctx* ctx);
int swb_download(swb_ctx* ctx, swb_result* results, uint32_t* cigar_arena,
                 int64_t cigar_cap, int64_t* cigar_used);

int swb_get_timing(const swb_ctx* ctx, swb_timing* out);

/* Host helper for callers that stitch the outputs of several contexts (one per GPU) into one result array and one arena:
 * adds `base` to cigar_off of every record that has a CIGAR. */
void swb_rebase_cigar_offsets(swb_result* results, int64_t n, int64_t base);
/* ... and uploads only the part of a sequence table a shard refers to: byte extent [extent[0], extent[1]) of entries
 * off[0..n) / len[0..n) (shift = 0 one byte per base, 1 SWB_SEQ_PACKED4, 2 SWB_SEQ_PACKED2) and their offsets rebased to
 * extent[0] in out_off.  Returns -1 on a negative offset or length. */
int swb_slice_table(const int64_t* off, const int32_t* len, int64_t n, int shift, int64_t* out_off, int64_t* extent);

/* ------------------------------------------------------------------ */
/* 3. CIGAR -> indel records                     */
/* ------------------------------------------------------------------ */
/* One record per I / D token of the alignment's CIGAR after indelPost's make_insertion_first reordering
 * (utilities.pyx:360-401): the integer core of findall_indels (localn.pyx:542-621).  The strings the
 * reference puts into its dicts are slices at these indices: lt_ref = ref[:ref_idx], lt_flank = read[:read_idx],
 * indel_seq = read[read_idx : read_idx+len] (I), del_seq = ref[ref_idx : ref_idx+len] (D), pos = genome_aln_pos + pos_off. */
typedef struct {
    int32_t  pair;       /* index of the alignment                                  */
    uint32_t cigar_op;   /* len << 4 | op (1 = I, 2 = D), BAM packing               */
    int32_t  ref_idx;    /* reference index of the event (relative like ref_begin1) */
    int32_t  read_idx;   /* read index of the event                                 */
    int32_t  pos_off;    /* pos - genome_aln_pos (starts at -1, localn.pyx:544)     */
} swb_indel;

/* Indels of the alignments the context holds after swb_compute() / swb_align_batch().  indel_off / indel_cnt /
 * read_end have n_pairs entries: the records of pair p are indels[indel_off[p] .. +indel_cnt[p]) in CIGAR order,
 * read_end[p] is the read index after the last token (start of rt_clipped, localn.pyx:615).  Returns -2 with
 * *used = required capacity when `cap` records are not enough. */
int swb_indels(swb_ctx* ctx, int64_t* indel_off, int32_t* indel_cnt, int32_t* read_end,
               swb_indel* indels, int64_t cap, int64_t* used);
/* The same for caller-supplied alignments: n CIGARs in BAM packing inside `cigar_arena` (arena_len uint32s). */
int swb_indels_from_cigars(swb_ctx* ctx, int32_t n, const uint32_t* cigar_arena, int64_t arena_len,
                           const int64_t* cigar_off, const int32_t* cigar_len,
                           const int32_t* ref_start, const int32_t* read_start,
                           int64_t* indel_off, int32_t* indel_cnt, int32_t* read_end,
                           swb_indel* indels, int64_t cap, int64_t* used);

/* pinned host memory helpers for the staging buffers */
void* swb_host_alloc(int64_t bytes);
void  swb_host_free(void* p);

/* sswpy.pyx:16-29 DNA_BASE_LUT on the host (for callers that want codes) */
void swb_encode_dna(const char* ascii, int8_t* codes, int64_t len);

/* Pack a sequence table for SWB_SEQ_PACKED4 (bits = 4) or SWB_SEQ_PACKED2 (bits = 2).  `blob` holds codes, or ASCII
 * when src_ascii != 0 (then encoded with DNA_BASE_LUT first).  Entry i is written at dst + dst_off[i] (consecutive,
 * byte aligned); returns the number of bytes written, or -1 if bits = 2 meets a code above 3 or an argument is bad.
 * dst needs sum(ceil(len[i] * bits / 8)) bytes. */
int64_t swb_pack_table(const int8_t* blob, const int64_t* off, const int32_t* len, int32_t n, int src_ascii, int bits,
                       uint8_t* dst, int64_t* dst_off);

const char* swb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SWB200_H */
