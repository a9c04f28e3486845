// swb_cert.cuh — overflow certificate for provisionally accepted 16-bit results.
//
// ssw_align runs the 8-bit pass first and only escalates to 16 bits when that pass overflows
// (ssw.c:842-847).  The 8-bit pass is not plain Smith-Waterman (signed lazy-F exit test, ssw.c:311), so
// "true score >= 255 - bias" does not by itself prove that it overflowed.  What does:
// the 8-bit H of every cell is >= the H of the recurrence WITHOUT any vertical-gap (F) term, because the
// quirk can only drop F contributions and every operation is monotone.  So if some path made only of
// diagonal and horizontal-gap moves scores >= 255 - bias, the 8-bit pass certainly saturated and the
// 16-bit result stands.  We look for such a path inside the alignment the traceback just produced:
// walk the CIGAR, keep a running (Kadane) score that restarts after every insertion.  Pairs without a
// certificate are re-run through the exact 8-bit emulation (k_exact<0,0>), which either confirms the
// overflow or yields the byte-mode result.
#pragma once
#include "swb_common.cuh"

__global__ void k_certify(SwbDev d, int32_t p0, int32_t p1)
{
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    const int st = d.p_state[p];
    if (!(st & PST_NEED_CERT)) return;
    const swb_result& r = d.res[p];
    const int limit = 255 - d.bias;
    bool ok = false;
    if (r.cigar_len > 0 && r.ref_begin1 >= 0 && r.read_begin1 >= 0) {
        const int8_t* read = d.reads + d.p_roff[p];
        const int8_t* ref = d.windows + d.p_woff[p];
        const int L = d.p_rlen[p], nc = d.p_wlen[p];
        const int go = d.gap_open[p], ge = d.gap_ext[p], n = d.n;
        const uint32_t* cg = d.cigar + r.cigar_off;
        int i = r.read_begin1, j = r.ref_begin1, S = 0;
        for (int k = 0; k < r.cigar_len && !ok; ++k) {
            const int len = (int)(cg[k] >> 4), op = (int)(cg[k] & 15);
            if (op == 0) {
                for (int q = 0; q < len && i < L && j < nc; ++q, ++i, ++j) {
                    S += d.mat[ref[j] * n + read[i]];
                    if (S < 0) S = 0;
                    if (S >= limit) { ok = true; break; }
                }
            } else if (op == 2) {           // deletion = horizontal gap (E): stays inside the F-free recurrence
                S -= go + (len - 1) * ge;
                if (S < 0) S = 0;
                j += len;
            } else {                        // insertion = vertical gap (F): restart
                S = 0;
                i += len;
            }
        }
    }
    if (ok) { d.p_state[p] = st & ~PST_NEED_CERT; return; }
    d.p_state[p] = (st & ~PST_NEED_CERT) | PST_HAVE_WORD;
    atomicAdd(d.counters + CNT_CERT_FAIL, 1);
    list_push(d.list[LIST_BYTE_FWD], d.counters + CNT_BYTE_FWD, p);
}
