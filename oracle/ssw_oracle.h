/* ssw_oracle.h — interface of the CPU parity oracle (TEST INFRASTRUCTURE ONLY, see ssw_oracle.c). */
#ifndef SSW_ORACLE_H
#define SSW_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint16_t score; int32_t ref; int32_t read; } orc_end;   /* alignment_end, ssw.c:104-108 */

/* s_align without the heap pointer (ssw.h:55-66) */
typedef struct {
    uint16_t score1, score2;
    int32_t  ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    int32_t  cigarLen;
    uint16_t flag;
} orc_result;

/* same layout as swb_result (include/swb200.h) */
typedef struct {
    uint16_t score1, score2;
    int32_t  ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    int32_t  cigar_len;
    uint16_t flag;
    uint16_t status;
    int64_t  cigar_off;
} orc_batch_result;

#define ORC_OK              0
#define ORC_NULL_BYTE_ONLY  1   /* ssw_align returns NULL, ssw.c:848-852 */
#define ORC_NULL_NO_PROFILE 2   /* ssw.c:856-859 */

int orc_align(const int8_t* read, int32_t readLen, const int8_t* mat, int32_t n, int8_t score_size,
              const int8_t* ref, int32_t refLen, uint8_t go, uint8_t ge,
              uint8_t flag, uint16_t filters, int32_t filterd, int32_t maskLen,
              orc_result* r, uint32_t** cigar_out);

void orc_encode_dna(const char* s, int8_t* out, int64_t len);

int64_t orc_align_batch(int32_t n_pairs,
                        const int8_t* reads, const int64_t* read_off, const int32_t* read_len,
                        const int8_t* windows, const int64_t* win_off, const int32_t* win_len,
                        const int32_t* pair_read, const int32_t* pair_win,
                        const int32_t* ref_beg, const int32_t* ref_len,
                        const uint8_t* gap_open, const uint8_t* gap_ext, const int32_t* mask_len,
                        const int8_t* mat, int32_t n, int8_t score_size,
                        uint8_t flag, uint16_t filters, int32_t filterd,
                        orc_batch_result* results, uint32_t* cigar_arena, int64_t cigar_cap);
#ifdef __cplusplus
}
#endif
#endif
