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
//
// Insertions the running score may continue across (everything else restarts it):
//  (a) the gap opens and ends inside ONE striped segment (rows r with the same r / segLen, segLen =
//      ceil(readLen/16), ssw.c:169): the main loop carries vF through those rows (ssw.c:294-295), the
//      lazy loop is not involved;
//  (c) the gap sits in a window column before the first column whose maximum reaches 128+go+ge (recorded by the
//      forward sweep): every F value of that column is < 128+ge, the signed test is exact there.
//  (b) every value of the chain is >= 128 + go.  The lazy loop goes on while some lane has
//      (int8)(vF - ge) > (int8)(H - go) (ssw.c:309-311) with H >= vF (ssw.c:306); a live chain is mis-seen
//      only if vF - ge >= 128 > H - go, i.e. only for vF in [128 + ge, 127 + go].  The vF a lane carries is
//      the maximum over all paths, hence >= the value on our path; if ours never drops below 128 + go
//      neither does the true one, both operands of the test are >= 128 and the signed compare orders
//      them correctly.
// A deletion right after an insertion also restarts: E is computed from H before the lazy correction
// (ssw.c:287-291, 301).
#pragma once
#include "swb_common.cuh"

// returns true if the pair's provisional 16-bit result is certified (or needs no certificate); otherwise flags it
// PST_HAVE_WORD and queues it for the exact 8-bit pass.  Called by the thread that just emitted the CIGAR.
__device__ __forceinline__ void certify_pair(const SwbDev& d, int p, int verifyList)
{
    const int st = d.p_state[p];
    if (!(st & PST_NEED_CERT) || !(st & PST_BAND_DONE)) return;
    const swb_result& r = d.res[p];
    const int limit = 255 - d.bias;
    bool ok = false;
    if (r.cigar_len > 0 && r.ref_begin1 >= 0 && r.read_begin1 >= 0 && r.cigar_off + r.cigar_len <= d.cigar_cap) {
        const int8_t* read = d.reads + d.p_roff[p];
        const int8_t* ref = d.windows + d.p_woff[p];
        const int L = d.p_rlen[p], nc = d.p_wlen[p];
        const int go = d.gap_open[p], ge = d.gap_ext[p], n = d.n;
        const uint32_t* cg = d.cigar + r.cigar_off;
        int i = r.read_begin1, j = r.ref_begin1, S = 0;
        const int segLen = (L + 15) / 16;
        const int csafe = d.p_csafe[p];
        bool afterIns = false;
        for (int k = 0; k < r.cigar_len && !ok; ++k) {
            const int len = (int)(cg[k] >> 4), op = (int)(cg[k] & 15);
            if (op == 0) {
                for (int q = 0; q < len && i < L && j < nc; ++q, ++i, ++j) {
                    S += d.mat[ref[j] * n + read[i]];
                    if (S < 0) S = 0;
                    if (S >= limit) { ok = true; break; }
                }
            } else if (op == 2) {           // deletion = horizontal gap (E): stays inside the F-free recurrence
                if (afterIns) S = 0;
                S -= go + (len - 1) * ge;
                if (S < 0) S = 0;
                j += len;
            } else {                        // insertion = vertical gap (F) over read rows i .. i+len-1, opened from row i-1
                const int last = S - go - (len - 1) * ge;
                const bool sameSeg = i >= 1 && (i - 1) / segLen == (i + len - 1) / segLen;
                const bool above = last >= 128 + go;
                const bool lowCol = (j - 1) < csafe;          // the vertical gap runs in the column of the last matched cell
                if (sameSeg || above || lowCol) { S = last; if (S < 0) S = 0; }
                else S = 0;
                afterIns = true;
                i += len;
                continue;
            }
            afterIns = false;
        }
    }
    if (ok) { d.p_state[p] = st & ~PST_NEED_CERT; return; }
    d.p_state[p] = (st & ~PST_NEED_CERT) | PST_HAVE_WORD;
    atomicAdd(d.counters + CNT_CERT_FAIL, 1);
    list_push(d.list[verifyList], d.counters + verifyList, p);
}

#ifdef SWB_WITH_CERT_KERNEL   // defined by the one translation unit that launches it (swb200.cu)
// certificate pass over the pairs whose traceback has finished (PST_BAND_DONE) and that still carry PST_NEED_CERT
__global__ void k_certify_rest(SwbDev d, int32_t p0, int32_t p1, int verifyList)
{
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    certify_pair(d, p, verifyList);
}
#endif
