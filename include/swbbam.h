/* swbbam.h — C ABI of libswbbam.so: native pileup ingestion for the batched realignment path.
 *
 * Replaces what indelPost reaches through pysam/htslib in pileup.pyx:51-160 (`make_pileup` -> `fetch_reads` ->
 * `bam.fetch / bam.count`, `reference.fetch`, and the per-read field extraction at the top of `dictize_read`):
 *
 *   reference interface                                   | here
 *   ------------------------------------------------------+--------------------------------------------------
 *   pysam.AlignmentFile(path)            pileup.pyx:53     | swb_bam_open / swb_bam_close / swb_bam_n_ref / ...
 *   bam.references                       pileup.pyx:71     | swb_bam_ref_name / swb_bam_ref_len / swb_bam_tid
 *   bam.fetch(chrom, start, stop, until_eof=True)  :134    | swb_bam_fetch      (one columnar batch per region)
 *   bam.count(chrom, pos-1, pos, read_callback=..) :83     | swb_bam_count
 *   AlignedSegment.{query_name, query_sequence, query_qualities, cigarstring, reference_start, reference_end,
 *     mapping_quality, is_reverse, is_duplicate, is_secondary}  pileup.pyx:138-200 | columns of swb_bam_batch
 *   pysam.FastaFile(path).fetch / get_reference_length / references  pileup.pyx:69,290 | swb_fai_*
 *   cigar_ptrn.findall + start/end offsets + locate_indels + get_spliced_subreads + is_end_dirty + is_dirty +
 *     count_lowqual_non_ref_bases + get_ref_seq  (pileup.pyx:173-265, utilities.pyx:187-327) | swb_pileup_columns
 *
 * A batch is COLUMNAR: one array per field plus byte arenas for names / bases / qualities / CIGAR words, so that a whole
 * locus (or many loci) moves into the batched aligner without one Python object per read.  swb_bam_batch_pack4 turns the
 * BAM 4-bit bases into the SWB_SEQ_PACKED4 read table `swb_align_batch` (include/swb200.h) takes, in one pass.
 *
 * The library also WRITES coordinate-sorted BAM + BAI and FASTA + FAI (BASELINE.json's configs are "written to BAM"; pysam
 * is absent from the image), so the measured pipeline reads real files.  Plain pointers and sizes only; no GPU code, links zlib.
 * Every function that can fail returns NULL / a negative value and leaves a message in swb_bam_last_error().
 *
 * Limits (by design, none of indelPost's inputs needs more): BAM with a BAI index (no CRAM, SAM text or CSI); optional fields are
 * stepped over, not decoded, so a CIGAR of more than 65 535 operations kept in a CG tag is reported as the record's in-line
 * placeholder; the writer emits no optional fields and expects its input in coordinate order.  A handle is not thread-safe (block
 * cache); open one per thread -- swb_bam_fetch_pack4 does that internally.
 */
#ifndef SWBBAM_H
#define SWBBAM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct swb_bam swb_bam;
typedef struct swb_bam_writer swb_bam_writer;
typedef struct swb_fai swb_fai;

/* BAM flag bits (SAM spec 1.4) */
#define SWB_BAM_FPAIRED 1
#define SWB_BAM_FUNMAP 4
#define SWB_BAM_FREVERSE 16
#define SWB_BAM_FSECONDARY 256
#define SWB_BAM_FQCFAIL 512
#define SWB_BAM_FDUP 1024
#define SWB_BAM_FSUPPLEMENTARY 2048

/* One region's records, columnar.  Owned by the library; free with swb_bam_batch_free.  Offsets index the arenas. */
typedef struct {
    int64_t   n;            /* records                                                             */
    int32_t*  tid;          /* reference index                                                     */
    int32_t*  pos;          /* 0-based leftmost position        (AlignedSegment.reference_start)   */
    int32_t*  end;          /* 0-based exclusive end, -1 if unmapped / no CIGAR (reference_end)    */
    uint16_t* flag;
    uint8_t*  mapq;
    int32_t*  l_seq;        /* bases                                                               */
    int32_t*  n_cigar;      /* CIGAR operations                                                    */
    int32_t*  next_tid;
    int32_t*  next_pos;
    int32_t*  tlen;
    int64_t*  name_off;     /* into names (NUL terminated)                                         */
    int64_t*  seq_off;      /* into seq and qual (same offset, l_seq bytes each)                   */
    int64_t*  cigar_off;    /* into cigar (uint32 words, BAM packing len<<4|op, "MIDNSHP=X")       */
    char*     names;   int64_t names_len;
    uint8_t*  seq;     int64_t seq_len;    /* ASCII bases ("=ACMGRSVTWYHKDBN" decoding)            */
    uint8_t*  qual;                         /* phred, 0xff-filled when the record has no qualities  */
    uint8_t*  seq4;    int64_t seq4_len;   /* the record's own 4-bit bases, entry i at seq4_off[i]  */
    int64_t*  seq4_off;
    uint32_t* cigar;   int64_t cigar_len;
} swb_bam_batch;

const char* swb_bam_last_error(void);
const char* swb_bam_version(void);

/* ---- reading ---- */
swb_bam* swb_bam_open(const char* path);          /* reads the header; loads <path>.bai (or <stem>.bai) when present */
void     swb_bam_close(swb_bam* b);
int32_t  swb_bam_n_ref(const swb_bam* b);
const char* swb_bam_ref_name(const swb_bam* b, int32_t tid);
int64_t  swb_bam_ref_len(const swb_bam* b, int32_t tid);
int32_t  swb_bam_tid(const swb_bam* b, const char* name);      /* -1 if absent */
const char* swb_bam_header_text(const swb_bam* b, int64_t* len);
int      swb_bam_has_index(const swb_bam* b);

/* Records of reference `tid` overlapping [beg, end) (0-based, half open; htslib's rule: pos < end && pos+reflen > beg,
 * reflen counted as 1 for records without reference-consuming operations), in file order.  tid < 0: every record of the
 * file.  A record is dropped when (flag & exclude) != 0 or (flag & require) != require.  Uses the index when there is
 * one, else scans the file.  NULL on error. */
swb_bam_batch* swb_bam_fetch(swb_bam* b, int32_t tid, int64_t beg, int64_t end, uint32_t require, uint32_t exclude);
void swb_bam_batch_free(swb_bam_batch* batch);
/* Number of such records (pysam's count(); read_callback="all" is exclude = UNMAP|SECONDARY|QCFAIL|DUP). */
int64_t swb_bam_count(swb_bam* b, int32_t tid, int64_t beg, int64_t end, uint32_t require, uint32_t exclude);

/* SWB_SEQ_PACKED4 read table of the batch (codes of sswpy.pyx:16-29's DNA_BASE_LUT: A 0, C 1, G 2, T 3, everything
 * else 4; two per byte, low nibble first): entry i at dst + dst_off[i], l_seq[i] bases.  dst needs sum(ceil(l_seq/2))
 * bytes (= seq4_len).  Returns the bytes written. */
int64_t swb_bam_batch_pack4(const swb_bam_batch* batch, uint8_t* dst, int64_t* dst_off);

/* Many regions in ONE call, on host threads: the records of every region that pass the flag masks (and, optionally, have a
 * CIGAR / do not start at position 0: fetch_reads' conditions, pileup.pyx:138-155) as one SWB_SEQ_PACKED4 read table in region
 * order -- the reads of region r are entries [region_first[r], region_first[r+1]).  Regions are dealt to `threads` workers
 * (<= 0: one per core, at most 16), each with its own file handle and block cache over the shared header and index.  This is
 * the ingest of a whole chunk of loci for swb_align_batch without a Python object per locus.  NULL on error. */
typedef struct {
    int64_t   n_regions, n_reads;
    int64_t*  region_first;             /* n_regions + 1 */
    int64_t*  read_off;   int32_t* read_len;   /* table entries: byte offset, length in bases */
    uint8_t*  table;      int64_t table_len;
    int32_t*  pos;  int32_t* end;  uint16_t* flag;  uint8_t* mapq;     /* per read, like swb_bam_batch */
} swb_bam_pack;
swb_bam_pack* swb_bam_fetch_pack4(swb_bam* b, int64_t n_regions, const int32_t* tid, const int64_t* beg, const int64_t* end,
                                  uint32_t require, uint32_t exclude, int need_cigar, int drop_pos0, int threads);
void swb_bam_pack_free(swb_bam_pack* p);

/* CIGAR strings of the batch ("70M1D80M", what AlignedSegment.cigarstring returns): entry i NUL terminated at dst + off[i],
 * off has n + 1 entries.  Returns the bytes needed (an upper bound); nothing is written when cap is smaller or dst is NULL. */
int64_t swb_bam_batch_cigar_text(const swb_bam_batch* batch, char* dst, int64_t cap, int64_t* off);

/* ---- per-read pileup columns (the integer core of dictize_read, pileup.pyx:160-266) ---- */
typedef struct {
    int32_t aln_start;      /* reference_start + 1                                   pileup.pyx:177 */
    int32_t start_offset;   /* leading soft clip                                      :178           */
    int32_t read_start;     /* aln_start - start_offset                               :179           */
    int32_t aln_end;        /* reference_end                                          :181           */
    int32_t end_offset;     /* trailing soft clip                                     :185           */
    int32_t read_end;       /* aln_end + end_offset                                   :187           */
    int32_t low_qual_base_num;  /* count_lowqual_non_ref_bases, utilities.pyx:187-218; -1 if ref_seq was not supplied */
    uint8_t is_end_dirty;   /* pileup.pyx:345-365                                                    */
    uint8_t is_dirty;       /* > 15 % of the bases at or below the quality threshold  :214           */
    uint8_t is_covering;    /* parse_spliced_read, pileup.pyx:391-403                                */
    uint8_t is_spliced;     /* more than one spliced subread                                         */
    int32_t covering_start, covering_end;  /* the covering subread (valid when is_covering)          */
    int32_t intron_start, intron_end;      /* intron_pattern, (0, 0) when none       :431-433        */
    int32_t n_subreads;     /* spliced subreads (get_spliced_subreads, utilities.pyx:243-278)        */
    int64_t subread_off;    /* into the subreads arena: n_subreads (start, end) int32 pairs          */
    int32_t n_ins, n_del;   /* locate_indels, utilities.pyx:307-328                                  */
    int64_t indel_off;      /* into the indel arena: n_ins insertions then n_del deletions           */
    uint8_t is_reference_seq;   /* read_seq == ref_seq (only when ref_seq was supplied)              */
    uint8_t n_count_gt1;        /* the CIGAR holds more than one N                                   */
    uint8_t pad_[2];
    int32_t splice_pos;     /* the position parse_spliced_read ends up comparing introns with (pos, or rpos once a
                               subread covered rpos only, pileup.pyx:399-403): for the splice_pattern strings */
} swb_pileup_read;

/* One I / D event of a read with the split indices leftalign_indel_read needs (pileup.pyx:303-342 via utilities.split:429-503):
 * lt_flank = read_seq[:read_split], rt_flank = read_seq[read_split:], lt_ref = ref_seq[:ref_split], ... (Python slice
 * semantics; an index may be negative exactly where the reference's would be). */
typedef struct {
    int32_t pos;          /* 1-based position of the base before the event */
    int32_t len;
    int32_t read_split;
    int32_t ref_split;
} swb_pileup_indel;

typedef struct {
    int64_t n;
    swb_pileup_read*  reads;
    int32_t*          subreads;  int64_t n_subreads;  /* pairs */
    swb_pileup_indel* indels;    int64_t n_indels;
    uint8_t*          ref_seq;   int64_t ref_seq_len; /* per-read ref_seq arena (get_ref_seq, pileup.pyx:269-299), when a contig was given */
    int64_t*          ref_seq_off;                    /* n + 1 offsets */
} swb_pileup_cols;

/* pos / rpos: the target's position and the right-most equivalent position (pileup.pyx:66-67); basequalthresh as in
 * VariantAlignment.  contig / contig_start / contig_len: reference bases covering every read ([contig_start,
 * contig_start+contig_len) 0-based; the whole chromosome or a slice), or NULL to skip ref_seq / low_qual_base_num /
 * is_reference_seq.  Unspliced reads take ref_seq from the slice [local_start, local_start+local_len) only, like
 * UnsplicedLocalReference.get_ref_seq (local_reference.pyx:33-36); spliced reads from the contig like reference.fetch.
 * Records without a CIGAR get zeroed columns (fetch_reads drops them, pileup.pyx:138-155). */
swb_pileup_cols* swb_pileup_columns(const swb_bam_batch* batch, int32_t pos, int32_t rpos, int32_t basequalthresh,
                                    const uint8_t* contig, int64_t contig_start, int64_t contig_len,
                                    int64_t local_start, int64_t local_len);
void swb_pileup_cols_free(swb_pileup_cols* c);
int32_t swb_pileup_read_size(void);   /* sizeof(swb_pileup_read), for bindings that mirror the struct */

/* ---- writing (coordinate-sorted input; BAI written by close when index != 0) ---- */
swb_bam_writer* swb_bam_create(const char* path, const char* header_text, int32_t n_ref, const char* const* ref_names,
                               const int64_t* ref_lens, int level);
/* n records, columnar like swb_bam_batch: seq ASCII at seq + seq_off[i] (l_seq[i] bytes), qual at the same offset in
 * `qual` (NULL: no qualities), cigar words at cigar + cigar_off[i], names NUL terminated at names + name_off[i].
 * next_tid / next_pos / tlen may be NULL (-1, -1, 0).  Returns 0, or -1 (unsorted input, bad field). */
int swb_bam_write(swb_bam_writer* w, int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, const uint8_t* mapq,
                  const int32_t* l_seq, const int32_t* n_cigar, const int64_t* name_off, const int64_t* seq_off, const int64_t* cigar_off,
                  const char* names, const uint8_t* seq, const uint8_t* qual, const uint32_t* cigar,
                  const int32_t* next_tid, const int32_t* next_pos, const int32_t* tlen);
int swb_bam_writer_close(swb_bam_writer* w, int index);

/* ---- FASTA + .fai ---- */
int      swb_fasta_write(const char* path, int32_t n, const char* const* names, const char* const* seqs, const int64_t* lens, int line_width); /* writes <path> and <path>.fai */
swb_fai* swb_fai_open(const char* path);     /* builds the index in memory when <path>.fai is missing */
void     swb_fai_close(swb_fai* f);
int32_t  swb_fai_n(const swb_fai* f);
const char* swb_fai_name(const swb_fai* f, int32_t i);
int64_t  swb_fai_len(const swb_fai* f, const char* name);     /* -1 if absent */
/* bases [beg, end) of `name` (clamped to the sequence like pysam's fetch) into dst (end - beg bytes at most); returns the
 * number of bases written, -1 if the sequence is absent. */
int64_t  swb_fai_fetch(const swb_fai* f, const char* name, int64_t beg, int64_t end, char* dst);
/* n slices back to back: slice i at dst + off[i], off has n + 1 entries.  Returns the bytes written, -(bytes needed) when cap is
 * too small, INT64_MIN when a sequence is absent. */
int64_t  swb_fai_fetch_many(const swb_fai* f, int64_t n, const char* const* names, const int64_t* beg, const int64_t* end, char* dst, int64_t cap, int64_t* off);

#ifdef __cplusplus
}
#endif
#endif /* SWBBAM_H */
