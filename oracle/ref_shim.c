/*
 * ref_shim.c — batch driver around the UNMODIFIED reference ssw.c.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with
 * /root/reference/indelpost/ssw.c (from where it lies; no reference source is
 * copied into this repository) into oracle/_ref/libssw_ref.so.  It only calls
 * the reference's public entry points (ssw_init / ssw_align / align_destroy /
 * init_destroy, ssw.h:86-139) in a loop and copies the s_align fields out, so
 * the tests can (a) pin oracle/ssw_oracle.c against the real thing,
 * (b) generate tests/golden/, and (c) time the reference on host cores
 * (bench.py --impl reference, cpu_baseline.kind = "reference").
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "ssw.h"          /* -I/root/reference/indelpost */

typedef struct {
    uint16_t score1, score2;
    int32_t  ref_begin1, ref_end1, read_begin1, read_end1, ref_end2;
    int32_t  cigar_len;
    uint16_t flag;
    uint16_t status;
    int64_t  cigar_off;
} ref_batch_result;

/* pairs [first, first+count) of the batch; CIGARs go to cigar_arena starting at
 * *used (not thread-safe on the arena: callers give each worker its own arena). */
int64_t ref_align_batch(int32_t first, int32_t count,
                        const int8_t* reads, const int64_t* read_off, const int32_t* read_len,
                        const int8_t* windows, const int64_t* win_off, const int32_t* win_len,
                        const int32_t* pair_read, const int32_t* pair_win,
                        const int32_t* ref_beg, const int32_t* ref_len,
                        const uint8_t* gap_open, const uint8_t* gap_ext, const int32_t* mask_len,
                        const int8_t* mat, int32_t n, int8_t score_size,
                        uint8_t flag, uint16_t filters, int32_t filterd,
                        ref_batch_result* results, uint32_t* cigar_arena, int64_t cigar_cap)
{
    int64_t used = 0;
    int32_t p;
    for (p = first; p < first + count; ++p) {
        const int32_t ri = pair_read[p], wi = pair_win[p];
        const int32_t rl = read_len[ri];
        const int32_t rb = ref_beg ? ref_beg[p] : 0;
        const int32_t wl = ref_len ? ref_len[p] : win_len[wi] - rb;
        const int32_t ml = mask_len ? mask_len[p] : (rl / 2 < 15 ? 15 : rl / 2);
        ref_batch_result* o = &results[p];
        s_profile* prof = ssw_init(reads + read_off[ri], rl, mat, n, score_size);
        s_align* a = ssw_align(prof, windows + win_off[wi] + rb, wl, gap_open[p], gap_ext[p],
                               flag, filters, filterd, ml);
        memset(o, 0, sizeof *o);
        if (!a) { o->status = 1; o->ref_begin1 = -1; o->read_begin1 = -1; init_destroy(prof); continue; }
        o->score1 = a->score1; o->score2 = a->score2;
        o->ref_begin1 = a->ref_begin1; o->ref_end1 = a->ref_end1;
        o->read_begin1 = a->read_begin1; o->read_end1 = a->read_end1; o->ref_end2 = a->ref_end2;
        o->flag = a->flag; o->cigar_len = a->cigar ? a->cigarLen : 0; o->cigar_off = used;
        if (a->cigar && a->cigarLen > 0) {
            if (used + a->cigarLen <= cigar_cap) memcpy(cigar_arena + used, a->cigar, (size_t)a->cigarLen * 4);
            used += a->cigarLen;
        }
        align_destroy(a);
        init_destroy(prof);
    }
    return used;
}
