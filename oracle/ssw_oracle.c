/*
 * ssw_oracle.c — CPU restatement of indelPost's Striped-Smith-Waterman path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for libswb200.
 * Nothing in the product (indelpost_b200/, include/) may import, link or call
 * it; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
 *
 * Parity pinning: the reference ships no tests or golden vectors
 * (SURVEY.md §4), so this restatement is pinned against the reference's own
 * ssw.c compiled unmodified into oracle/_ref/libssw_ref.so (oracle/Makefile)
 * — tests/test_oracle.py fuzzes the two against each other when the
 * _ref library is present, and tests/golden/ holds vectors generated from
 * _ref by tests/golden/make_golden.py; at pipeline level tests/test_pipeline_parity.py
 * plugs it into the unmodified reference pipeline (oracle/_ref_pipeline).
 *
 * It is a scalar, intrinsic-free restatement: each SSE2 register of the
 * reference becomes an array of W lanes (W=16 unsigned bytes in "byte mode",
 * W=8 signed 16-bit words in "word mode"); the striped layout, the lazy-F
 * loop with its early exit, the overflow escalation, the end-position
 * tie-breaking, the sub-optimal-score mask rule and banded_sw's rolling
 * buffers are followed literally because the reference's outputs depend on
 * them.
 *
 * Reference citations are to /root/reference/indelpost/ssw.c unless noted.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "ssw_oracle.h"

/* ------------------------------------------------------------------ */
/* small lane helpers                                                  */
/* ------------------------------------------------------------------ */

static inline uint8_t addsat_u8(uint8_t a, uint8_t b) { unsigned s = (unsigned)a + b; return s > 255u ? 255u : (uint8_t)s; }
static inline uint8_t subsat_u8(uint8_t a, uint8_t b) { return a > b ? (uint8_t)(a - b) : 0; }
static inline uint8_t max_u8(uint8_t a, uint8_t b) { return a > b ? a : b; }

static inline int16_t addsat_s16(int16_t a, int16_t b) {
    int32_t s = (int32_t)a + b;
    if (s > 32767) s = 32767;
    if (s < -32768) s = -32768;
    return (int16_t)s;
}
/* _mm_subs_epu16 on the bit patterns (ssw.c:493-499 mix signed max with unsigned saturating subtract) */
static inline int16_t subsat_u16(int16_t a, int16_t b) {
    uint16_t ua = (uint16_t)a, ub = (uint16_t)b;
    return (int16_t)(ua > ub ? (uint16_t)(ua - ub) : 0);
}
static inline int16_t max_s16(int16_t a, int16_t b) { return a > b ? a : b; }

/* ------------------------------------------------------------------ */
/* byte mode: sw_sse2_byte, ssw.c:197-384; profile rule of qP_byte,    */
/* ssw.c:163-188 (pad cells = bias, real cells = mat + bias)           */
/* ------------------------------------------------------------------ */

#define WB 16

static void orc_sw_byte(const int8_t* ref, int ref_dir, int32_t refLen,
                        const int8_t* read, int32_t readLen,
                        const int8_t* mat, int32_t n,
                        uint8_t go, uint8_t ge, uint8_t terminate, uint8_t bias,
                        int32_t maskLen, orc_end out[2])
{
    const int32_t segLen = (readLen + WB - 1) / WB;            /* ssw.c:221 */
    const size_t  colBytes = (size_t)segLen * WB;
    uint8_t* hA   = (uint8_t*)calloc(colBytes ? colBytes : 1, 1);
    uint8_t* hB   = (uint8_t*)calloc(colBytes ? colBytes : 1, 1);
    uint8_t* eCol = (uint8_t*)calloc(colBytes ? colBytes : 1, 1);
    uint8_t* hBest= (uint8_t*)calloc(colBytes ? colBytes : 1, 1);
    uint8_t* colMax = (uint8_t*)calloc(refLen > 0 ? (size_t)refLen : 1, 1);   /* ssw.c:224 */
    uint8_t* hStore = hA; uint8_t* hLoad = hB;
    uint8_t  laneBest[WB], laneMark[WB];
    uint8_t  best = 0;                                          /* ssw.c:218 */
    int32_t  end_read = readLen - 1;                            /* ssw.c:219 */
    int32_t  end_ref = -1;                                      /* ssw.c:220 */
    int32_t  i, j, k, l;
    int32_t  begin = 0, end = refLen, step = 1;
    memset(laneBest, 0, sizeof laneBest);
    memset(laneMark, 0, sizeof laneMark);
    if (ref_dir == 1) { begin = refLen - 1; end = -1; step = -1; }          /* ssw.c:253-257 */

    for (i = begin; i != end; i += step) {
        uint8_t vF[WB], vH[WB], vColMax[WB];
        uint8_t* tmp;
        int stop_lazy = 0;
        memset(vF, 0, sizeof vF);
        memset(vColMax, 0, sizeof vColMax);
        /* diagonal feed: last stripe of the previous column shifted by one lane (ssw.c:264-265) */
        vH[0] = 0;
        for (l = 1; l < WB; ++l) vH[l] = segLen ? hStore[(size_t)(segLen - 1) * WB + (l - 1)] : 0;
        tmp = hLoad; hLoad = hStore; hStore = tmp;              /* ssw.c:269-271 */

        for (j = 0; j < segLen; ++j) {                          /* ssw.c:274-299 */
            for (l = 0; l < WB; ++l) {
                const int32_t r = j + l * segLen;               /* striped row index, ssw.c:180-183 */
                const uint8_t p = (uint8_t)(int8_t)(r >= readLen ? bias : mat[ref[i] * n + read[r]] + bias);
                uint8_t h = addsat_u8(vH[l], p);
                uint8_t e = eCol[(size_t)j * WB + l];
                h = subsat_u8(h, bias);
                h = max_u8(h, e);
                h = max_u8(h, vF[l]);
                vColMax[l] = max_u8(vColMax[l], h);
                hStore[(size_t)j * WB + l] = h;
                h = subsat_u8(h, go);
                e = subsat_u8(e, ge);
                e = max_u8(e, h);
                eCol[(size_t)j * WB + l] = e;
                vF[l] = max_u8(subsat_u8(vF[l], ge), h);
                vH[l] = hLoad[(size_t)j * WB + l];
            }
        }

        /* lazy-F loop with its signed-byte exit test (ssw.c:302-313) */
        for (k = 0; k < WB && !stop_lazy; ++k) {
            for (l = WB - 1; l > 0; --l) vF[l] = vF[l - 1];
            vF[0] = 0;
            for (j = 0; j < segLen; ++j) {
                int any = 0;
                for (l = 0; l < WB; ++l) {
                    uint8_t h = hStore[(size_t)j * WB + l];
                    h = max_u8(h, vF[l]);
                    vColMax[l] = max_u8(vColMax[l], h);
                    hStore[(size_t)j * WB + l] = h;
                    h = subsat_u8(h, go);
                    vF[l] = subsat_u8(vF[l], ge);
                    if ((int8_t)vF[l] > (int8_t)h) any = 1;      /* _mm_cmpgt_epi8 is a SIGNED compare, ssw.c:311 */
                }
                if (!any) { stop_lazy = 1; break; }
            }
        }

        /* running maximum and best column (ssw.c:316-333) */
        {
            int changed = 0;
            for (l = 0; l < WB; ++l) {
                laneBest[l] = max_u8(laneBest[l], vColMax[l]);
                if (laneBest[l] != laneMark[l]) changed = 1;
            }
            if (changed) {
                uint8_t m = 0;
                for (l = 0; l < WB; ++l) { laneMark[l] = laneBest[l]; m = max_u8(m, laneBest[l]); }
                if (m > best) {
                    best = m;
                    if ((int)best + (int)bias >= 255) break;     /* overflow, ssw.c:327 */
                    end_ref = i;
                    memcpy(hBest, hStore, colBytes);
                }
            }
        }
        {
            uint8_t m = 0;
            for (l = 0; l < WB; ++l) m = max_u8(m, vColMax[l]);
            colMax[i] = m;                                       /* ssw.c:336 */
            if (m == terminate) break;                           /* ssw.c:337 */
        }
    }

    /* smallest read index holding the maximum in the best column (ssw.c:341-349) */
    for (i = 0; i < segLen * WB; ++i) {
        if (hBest[i] == best) {
            int32_t r = i / WB + (i % WB) * segLen;
            if (r < end_read) end_read = r;
        }
    }

    out[0].score = ((int)best + (int)bias >= 255) ? 255 : best;  /* ssw.c:358 */
    out[0].ref = end_ref;
    out[0].read = end_read;
    out[1].score = 0; out[1].ref = 0; out[1].read = 0;

    /* sub-optimal: first strict maximum outside the mask (ssw.c:366-379) */
    {
        int32_t edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;
        for (i = 0; i < edge; ++i)
            if (colMax[i] > out[1].score) { out[1].score = colMax[i]; out[1].ref = i; }
        edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
        for (i = edge + 1; i < refLen; ++i)                      /* starts at edge+1 in byte mode */
            if (colMax[i] > out[1].score) { out[1].score = colMax[i]; out[1].ref = i; }
    }
    free(hA); free(hB); free(eCol); free(hBest); free(colMax);
}

/* ------------------------------------------------------------------ */
/* word mode: sw_sse2_word, ssw.c:410-586; profile rule of qP_word,    */
/* ssw.c:386-408 (pad cells = 0, no bias)                              */
/* ------------------------------------------------------------------ */

#define WW 8

static void orc_sw_word(const int8_t* ref, int ref_dir, int32_t refLen,
                        const int8_t* read, int32_t readLen,
                        const int8_t* mat, int32_t n,
                        uint8_t go8, uint8_t ge8, uint16_t terminate,
                        int32_t maskLen, orc_end out[2])
{
    const int32_t segLen = (readLen + WW - 1) / WW;             /* ssw.c:428 */
    const size_t  colElems = (size_t)segLen * WW;
    int16_t* hA   = (int16_t*)calloc(colElems ? colElems : 1, sizeof(int16_t));
    int16_t* hB   = (int16_t*)calloc(colElems ? colElems : 1, sizeof(int16_t));
    int16_t* eCol = (int16_t*)calloc(colElems ? colElems : 1, sizeof(int16_t));
    int16_t* hBest= (int16_t*)calloc(colElems ? colElems : 1, sizeof(int16_t));
    uint16_t* colMax = (uint16_t*)calloc(refLen > 0 ? (size_t)refLen : 1, sizeof(uint16_t));
    int16_t* hStore = hA; int16_t* hLoad = hB;
    int16_t  laneBest[WW], laneMark[WW];
    const int16_t go = (int16_t)go8, ge = (int16_t)ge8;         /* _mm_set1_epi16(uint8), ssw.c:446-449 */
    uint16_t best = 0;                                          /* ssw.c:425 */
    int32_t  end_read = readLen - 1;
    int32_t  end_ref = 0;                                       /* ssw.c:427 (byte mode starts at -1) */
    int32_t  i, j, k, l;
    int32_t  begin = 0, end = refLen, step = 1;
    memset(laneBest, 0, sizeof laneBest);
    memset(laneMark, 0, sizeof laneMark);
    if (ref_dir == 1) { begin = refLen - 1; end = -1; step = -1; }

    for (i = begin; i != end; i += step) {
        int16_t vF[WW], vH[WW], vColMax[WW];
        int16_t* tmp;
        int stop_lazy = 0;
        memset(vF, 0, sizeof vF);
        memset(vColMax, 0, sizeof vColMax);
        vH[0] = 0;
        for (l = 1; l < WW; ++l) vH[l] = segLen ? hStore[(size_t)(segLen - 1) * WW + (l - 1)] : 0;   /* ssw.c:467-468 */
        tmp = hLoad; hLoad = hStore; hStore = tmp;

        for (j = 0; j < segLen; ++j) {                          /* ssw.c:480-504 */
            for (l = 0; l < WW; ++l) {
                const int32_t r = j + l * segLen;
                const int16_t p = (int16_t)(r >= readLen ? 0 : mat[ref[i] * n + read[r]]);
                int16_t h = addsat_s16(vH[l], p);
                int16_t e = eCol[(size_t)j * WW + l];
                h = max_s16(h, e);
                h = max_s16(h, vF[l]);
                vColMax[l] = max_s16(vColMax[l], h);
                hStore[(size_t)j * WW + l] = h;
                h = subsat_u16(h, go);
                e = subsat_u16(e, ge);
                e = max_s16(e, h);
                eCol[(size_t)j * WW + l] = e;
                vF[l] = max_s16(subsat_u16(vF[l], ge), h);
                vH[l] = hLoad[(size_t)j * WW + l];
            }
        }

        for (k = 0; k < WW && !stop_lazy; ++k) {                /* ssw.c:507-518 */
            for (l = WW - 1; l > 0; --l) vF[l] = vF[l - 1];
            vF[0] = 0;
            for (j = 0; j < segLen; ++j) {
                int any = 0;
                for (l = 0; l < WW; ++l) {
                    int16_t h = hStore[(size_t)j * WW + l];
                    h = max_s16(h, vF[l]);
                    vColMax[l] = max_s16(vColMax[l], h);
                    hStore[(size_t)j * WW + l] = h;
                    h = subsat_u16(h, go);
                    vF[l] = subsat_u16(vF[l], ge);
                    if (vF[l] > h) any = 1;                      /* _mm_cmpgt_epi16, ssw.c:516 */
                }
                if (!any) { stop_lazy = 1; break; }
            }
        }

        {
            int changed = 0;
            for (l = 0; l < WW; ++l) {
                laneBest[l] = max_s16(laneBest[l], vColMax[l]);
                if (laneBest[l] != laneMark[l]) changed = 1;
            }
            if (changed) {                                       /* ssw.c:521-535 */
                int16_t m = laneBest[0];
                for (l = 0; l < WW; ++l) { laneMark[l] = laneBest[l]; m = max_s16(m, laneBest[l]); }
                if ((uint16_t)m > best) {
                    best = (uint16_t)m;
                    end_ref = i;
                    memcpy(hBest, hStore, colElems * sizeof(int16_t));
                }
            }
        }
        {
            int16_t m = vColMax[0];
            for (l = 0; l < WW; ++l) m = max_s16(m, vColMax[l]);
            colMax[i] = (uint16_t)m;                             /* ssw.c:538 */
            if (colMax[i] == terminate) break;                   /* ssw.c:539 */
        }
    }

    for (i = 0; i < segLen * WW; ++i) {                          /* ssw.c:543-551 */
        if ((uint16_t)hBest[i] == best) {
            int32_t r = i / WW + (i % WW) * segLen;
            if (r < end_read) end_read = r;
        }
    }

    out[0].score = best; out[0].ref = end_ref; out[0].read = end_read;
    out[1].score = 0; out[1].ref = 0; out[1].read = 0;
    {
        int32_t edge = (end_ref - maskLen) > 0 ? (end_ref - maskLen) : 0;     /* ssw.c:568-581 */
        for (i = 0; i < edge; ++i)
            if (colMax[i] > out[1].score) { out[1].score = colMax[i]; out[1].ref = i; }
        edge = (end_ref + maskLen) > refLen ? refLen : (end_ref + maskLen);
        for (i = edge; i < refLen; ++i)                          /* starts at edge in word mode */
            if (colMax[i] > out[1].score) { out[1].score = colMax[i]; out[1].ref = i; }
    }
    free(hA); free(hB); free(eCol); free(hBest); free(colMax);
}

/* ------------------------------------------------------------------ */
/* banded_sw, ssw.c:588-772.  Followed literally, including the band   */
/* coordinate macros (ssw.c:92-95), the three rolling row buffers and  */
/* which of their slots are cleared at each row start (ssw.c:633):     */
/* when the band is wider than the matrix those details decide what a  */
/* cell reads as its upper neighbour.                                  */
/* ------------------------------------------------------------------ */

static inline int32_t band_u(int32_t w, int32_t i, int32_t j) {           /* set_u, ssw.c:92 */
    int32_t x = i - w; if (x < 0) x = 0; return j - x + 1;
}
static inline int64_t band_d(int32_t w, int32_t i, int32_t j, int32_t p) { /* set_d, ssw.c:95 */
    int32_t x = i - w; if (x < 0) x = 0; return (int64_t)(j - x) * 3 + p;
}

/* returns number of cigar ops written to *cigar_out (malloc'd), or -1 for the
 * "Trace back error" NULL return (ssw.c:711-719) */
static int32_t orc_banded_sw(const int8_t* ref, const int8_t* read, int32_t refLen, int32_t readLen,
                             int32_t score, uint32_t go, uint32_t ge, int32_t band_width,
                             const int8_t* mat, int32_t n, uint32_t** cigar_out)
{
    const int32_t len = refLen > readLen ? refLen : readLen;
    int32_t  best = 0;                                           /* not reset between widenings */
    int32_t* hPrev = NULL; int32_t* ePrev = NULL; int32_t* hCur = NULL;
    int8_t*  dir = NULL;
    int32_t  width = 0, width_d = 0;
    int32_t  i, j;

    do {                                                         /* ssw.c:612-669 */
        width = band_width * 2 + 3; width_d = band_width * 2 + 1;
        free(hPrev); free(ePrev); free(hCur); free(dir);
        /* the reference reallocs (keeping old contents); every slot it later reads is
         * (re)written first except where noted below, so fresh zeroed buffers differ only
         * on uninitialised reads, which the reference leaves undefined */
        hPrev = (int32_t*)calloc((size_t)width + 1, sizeof(int32_t));
        ePrev = (int32_t*)calloc((size_t)width + 1, sizeof(int32_t));
        hCur  = (int32_t*)calloc((size_t)width + 1, sizeof(int32_t));
        dir   = (int8_t*)calloc((size_t)width_d * (size_t)(readLen > 0 ? readLen : 1) * 3 + 3, 1);
        for (j = 1; j < width - 1; ++j) hPrev[j] = 0;            /* ssw.c:627 */
        for (i = 0; i < readLen; ++i) {
            int32_t beg = 0, end = refLen - 1, u = 0, edge, f;
            int8_t* line;
            j = i - band_width; if (j > beg) beg = j;            /* ssw.c:630 */
            j = i + band_width; if (j < end) end = j;            /* ssw.c:631 */
            edge = end + 1 < width - 1 ? end + 1 : width - 1;    /* ssw.c:632 */
            f = hPrev[0] = ePrev[0] = hPrev[edge] = ePrev[edge] = hCur[0] = 0;   /* ssw.c:633 */
            line = dir + (size_t)width_d * (size_t)i * 3;
            for (j = beg; j <= end; ++j) {                       /* ssw.c:636-665 */
                int32_t up, lf, dg, a, b, e1, f1, g, m;
                int64_t de, df, dh;
                u  = band_u(band_width, i, j);
                up = band_u(band_width, i - 1, j);
                lf = band_u(band_width, i, j - 1);
                dg = band_u(band_width, i - 1, j - 1);
                de = band_d(band_width, i, j, 0);
                df = band_d(band_width, i, j, 1);
                dh = band_d(band_width, i, j, 2);

                a = i == 0 ? -(int32_t)go : hPrev[up] - (int32_t)go;       /* ssw.c:644-648 */
                b = i == 0 ? -(int32_t)ge : ePrev[up] - (int32_t)ge;
                ePrev[u] = a > b ? a : b;
                line[de] = a > b ? 3 : 2;

                a = hCur[lf] - (int32_t)go;                                /* ssw.c:650-653 */
                b = f - (int32_t)ge;
                f = a > b ? a : b;
                line[df] = a > b ? 5 : 4;

                e1 = ePrev[u] > 0 ? ePrev[u] : 0;                          /* ssw.c:655-659 */
                f1 = f > 0 ? f : 0;
                g  = e1 > f1 ? e1 : f1;
                m  = hPrev[dg] + mat[ref[j] * n + read[i]];
                hCur[u] = g > m ? g : m;
                if (hCur[u] > best) best = hCur[u];                        /* ssw.c:661 */
                if (g <= m) line[dh] = 1;                                  /* ssw.c:663-664 */
                else line[dh] = e1 > f1 ? line[de] : line[df];
            }
            for (j = 1; j <= u; ++j) hPrev[j] = hCur[j];                   /* ssw.c:666 */
        }
        band_width *= 2;
    } while (best < score && band_width <= len);                          /* ssw.c:669 */
    band_width /= 2;

    /* trace back from the bottom-right corner (ssw.c:672-733) */
    {
        int32_t cap = 16, l = 0, e = 0, state = 2;
        uint32_t* c = (uint32_t*)malloc((size_t)cap * sizeof(uint32_t));
        char op = 'M', prev_op = 'M';
        const int64_t dirBytes = (int64_t)width_d * (int64_t)readLen * 3;
        i = readLen - 1; j = refLen - 1;
        while (i >= 0 && j > 0) {                                /* ssw.c:679 */
            int64_t at = (int64_t)width_d * (int64_t)i * 3 + band_d(band_width, i, j, state);
            int8_t  d = (at >= 0 && at < dirBytes) ? dir[at] : 0;
            switch (d) {
                case 1: --i; --j; state = 2; op = 'M'; break;
                case 2: --i;      state = 0; op = 'I'; break;
                case 3: --i;      state = 2; op = 'I'; break;
                case 4: --j;      state = 1; op = 'D'; break;
                case 5: --j;      state = 2; op = 'D'; break;
                default:
                    free(c); free(hPrev); free(ePrev); free(hCur); free(dir);
                    *cigar_out = NULL;
                    return -1;
            }
            if (op == prev_op) ++e;
            else {
                ++l;
                if (l >= cap) { cap *= 2; c = (uint32_t*)realloc(c, (size_t)cap * sizeof(uint32_t)); }
                c[l - 1] = (uint32_t)e << 4 | (prev_op == 'M' ? 0u : prev_op == 'I' ? 1u : 2u);
                prev_op = op; e = 1;
            }
        }
        if (l + 2 >= cap) { cap += 4; c = (uint32_t*)realloc(c, (size_t)cap * sizeof(uint32_t)); }
        if (op == 'M') {                                         /* ssw.c:734-751 */
            ++l;
            c[l - 1] = (uint32_t)(e + 1) << 4 | 0u;
        } else {
            l += 2;
            c[l - 2] = (uint32_t)e << 4 | (op == 'I' ? 1u : 2u);
            c[l - 1] = (uint32_t)1 << 4 | 0u;
        }
        {   /* reverse (ssw.c:753-762) */
            uint32_t* r = (uint32_t*)malloc((size_t)l * sizeof(uint32_t));
            for (i = 0; i < l; ++i) r[i] = c[l - 1 - i];
            free(c);
            *cigar_out = r;
        }
        free(hPrev); free(ePrev); free(hCur); free(dir);
        return l;
    }
}

/* ------------------------------------------------------------------ */
/* driver: ssw_init + ssw_align, ssw.c:787-920                          */
/* ------------------------------------------------------------------ */

int orc_align(const int8_t* read, int32_t readLen, const int8_t* mat, int32_t n, int8_t score_size,
              const int8_t* ref, int32_t refLen, uint8_t go, uint8_t ge,
              uint8_t flag, uint16_t filters, int32_t filterd, int32_t maskLen,
              orc_result* r, uint32_t** cigar_out)
{
    orc_end bests[2], rev[2];
    int have_byte = (score_size == 0 || score_size == 2);        /* ssw.c:793-802 */
    int have_word = (score_size == 1 || score_size == 2);
    int word = 0;
    int32_t bias = 0, i;
    int8_t* rread;
    memset(r, 0, sizeof *r);
    *cigar_out = NULL;
    r->ref_begin1 = -1; r->read_begin1 = -1;                     /* ssw.c:832-836 */
    if (have_byte) {
        for (i = 0; i < n * n; ++i) if (mat[i] < bias) bias = mat[i];   /* ssw.c:795-797 */
        bias = abs(bias);
    }
    if (have_byte) {                                             /* ssw.c:842-860 */
        orc_sw_byte(ref, 0, refLen, read, readLen, mat, n, go, ge, (uint8_t)-1, (uint8_t)bias, maskLen, bests);
        if (have_word && bests[0].score == 255) {
            orc_sw_word(ref, 0, refLen, read, readLen, mat, n, go, ge, (uint16_t)-1, maskLen, bests);
            word = 1;
        } else if (bests[0].score == 255) {
            return ORC_NULL_BYTE_ONLY;
        }
    } else if (have_word) {
        orc_sw_word(ref, 0, refLen, read, readLen, mat, n, go, ge, (uint16_t)-1, maskLen, bests);
        word = 1;
    } else {
        return ORC_NULL_NO_PROFILE;
    }
    r->score1 = bests[0].score; r->ref_end1 = bests[0].ref; r->read_end1 = bests[0].read;
    if (maskLen >= 15) { r->score2 = bests[1].score; r->ref_end2 = bests[1].ref; }   /* ssw.c:864-870 */
    else { r->score2 = 0; r->ref_end2 = -1; }
    if (flag == 0 || (flag == 2 && r->score1 < filters)) return ORC_OK;               /* ssw.c:872 */

    /* reverse pass on the reversed read prefix (ssw.c:875-886) */
    {
        int32_t rl = r->read_end1 + 1;
        rread = (int8_t*)calloc((size_t)(rl > 0 ? rl : 1), 1);
        for (i = 0; i < rl; ++i) rread[i] = read[rl - 1 - i];    /* seq_reverse, ssw.c:774-785 */
        if (!word) orc_sw_byte(ref, 1, r->ref_end1 + 1, rread, rl, mat, n, go, ge, (uint8_t)r->score1, (uint8_t)bias, maskLen, rev);
        else       orc_sw_word(ref, 1, r->ref_end1 + 1, rread, rl, mat, n, go, ge, r->score1, maskLen, rev);
        free(rread);
    }
    r->ref_begin1 = rev[0].ref;
    r->read_begin1 = r->read_end1 - rev[0].read;
    if (r->score1 > rev[0].score) r->flag = 2;                   /* ssw.c:888-891 */

    if ((7 & flag) == 0 || ((2 & flag) != 0 && r->score1 < filters) ||
        ((4 & flag) != 0 && (r->ref_end1 - r->ref_begin1 > filterd || r->read_end1 - r->read_begin1 > filterd)))
        return ORC_OK;                                           /* ssw.c:894 */

    {
        int32_t bRef = r->ref_end1 - r->ref_begin1 + 1;          /* ssw.c:897-900 */
        int32_t bRead = r->read_end1 - r->read_begin1 + 1;
        int32_t bw = abs(bRef - bRead) + 1;
        const int8_t* refp;
        int8_t one = 0;
        int32_t cl;
        /* score1 == 0 in byte mode leaves ref_begin1 == -1 and the reference reads ref[-1]
         * (undefined).  The 1x1 problem it then solves yields "1M" whatever that byte is
         * (SURVEY.md §10.1), so the oracle substitutes a defined base. */
        if (r->ref_begin1 < 0) { refp = &one; } else refp = ref + r->ref_begin1;
        cl = orc_banded_sw(refp, read + r->read_begin1, bRef, bRead, r->score1, go, ge, bw, mat, n, cigar_out);
        if (cl < 0) { r->flag = 1; r->cigarLen = 0; }            /* ssw.c:911 */
        else r->cigarLen = cl;
    }
    return ORC_OK;
}

/* sswpy.pyx:16-29 */
void orc_encode_dna(const char* s, int8_t* out, int64_t len)
{
    int64_t i;
    for (i = 0; i < len; ++i) {
        int8_t v = 4;
        switch (s[i]) {
            case 'A': case 'a': case 'U': case 'u': v = 0; break;
            case 'C': case 'c': v = 1; break;
            case 'G': case 'g': v = 2; break;
            case 'T': case 't': v = 3; break;
            default: v = 4;
        }
        out[i] = v;
    }
}

/* Batch driver with the same SoA layout as include/swb200.h (swb_batch).  One
 * thread; bench.py parallelises over processes for the cpu_baseline. */
int64_t orc_align_batch(int32_t n_pairs,
                        const int8_t* reads, const int64_t* read_off, const int32_t* read_len,
                        const int8_t* windows, const int64_t* win_off, const int32_t* win_len,
                        const int32_t* pair_read, const int32_t* pair_win,
                        const int32_t* ref_beg, const int32_t* ref_len,
                        const uint8_t* gap_open, const uint8_t* gap_ext, const int32_t* mask_len,
                        const int8_t* mat, int32_t n, int8_t score_size,
                        uint8_t flag, uint16_t filters, int32_t filterd,
                        orc_batch_result* results, uint32_t* cigar_arena, int64_t cigar_cap)
{
    int64_t used = 0;
    int32_t p;
    for (p = 0; p < n_pairs; ++p) {
        const int32_t ri = pair_read[p], wi = pair_win[p];
        const int32_t rl = read_len[ri];
        const int32_t rb = ref_beg ? ref_beg[p] : 0;
        const int32_t wl = ref_len ? ref_len[p] : win_len[wi] - rb;
        const int32_t ml = mask_len ? mask_len[p] : (rl / 2 < 15 ? 15 : rl / 2);
        orc_result r; uint32_t* cg = NULL;
        orc_batch_result* o = &results[p];
        int rc = orc_align(reads + read_off[ri], rl, mat, n, score_size,
                           windows + win_off[wi] + rb, wl, gap_open[p], gap_ext[p],
                           flag, filters, filterd, ml, &r, &cg);
        memset(o, 0, sizeof *o);
        if (rc != ORC_OK) { o->status = 1; o->ref_begin1 = -1; o->read_begin1 = -1; continue; }
        o->score1 = r.score1; o->score2 = r.score2;
        o->ref_begin1 = r.ref_begin1; o->ref_end1 = r.ref_end1;
        o->read_begin1 = r.read_begin1; o->read_end1 = r.read_end1; o->ref_end2 = r.ref_end2;
        o->flag = r.flag; o->cigar_len = r.cigarLen; o->cigar_off = used;
        if (r.cigarLen > 0) {
            if (used + r.cigarLen <= cigar_cap) memcpy(cigar_arena + used, cg, (size_t)r.cigarLen * 4);
            used += r.cigarLen;
        }
        free(cg);
    }
    return used;
}
