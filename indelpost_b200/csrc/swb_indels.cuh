// swb_indels.cuh — CIGAR -> indel records on the device: the integer core of indelPost's findall_indels
// (localn.pyx:542-621) including make_insertion_first / merge_consecutive_gaps (utilities.pyx:360-401), which every
// caller runs right after an SSW call.  One thread per alignment; the CIGARs are a few ops
// long, the records go to a compact arena (warp-aggregated bump allocation, per-pair offset + count).
//
// Reference semantics kept literally:
//  * a gap token swallows the gap tokens that directly follow it -- except that when the run reaches the END of the
//    CIGAR its last token is left out (utilities.pyx:366-375: the scan index has not stepped past a non-gap token);
//  * a merged run holding both kinds is reversed as a whole when it STARTS with a deletion (utilities.pyx:389-396);
//  * pos starts at genome_aln_pos - 1 and advances over M and D tokens (localn.pyx:544, 587, 610); ref_idx / read_idx
//    start at reference_start / read_start; any op other than I / D advances all three (localn.pyx:592-611).
#pragma once
#include "swb_common.cuh"

__device__ __forceinline__ bool indel_is_gap(uint32_t c) { const uint32_t op = c & 15u; return op == 1u || op == 2u; }

// walks the CIGAR of one alignment; emits to `out` (nullptr: count only).  Returns the number of records and, through
// read_end, the read index after the last token (start of rt_clipped, localn.pyx:615).
__device__ __forceinline__ int indel_walk(const uint32_t* __restrict__ c, int n, int ref_start, int read_start, int pair, swb_indel* out, int& read_end)
{
    int ref_idx = ref_start, read_idx = read_start, pos = -1, cnt = 0;
    auto token = [&](uint32_t t) {
        const int len = (int)(t >> 4);
        const uint32_t op = t & 15u;
        if (op == 1u || op == 2u) {
            if (out) { swb_indel r; r.pair = pair; r.cigar_op = t; r.ref_idx = ref_idx; r.read_idx = read_idx; r.pos_off = pos; out[cnt] = r; }
            ++cnt;
            if (op == 1u) read_idx += len; else { ref_idx += len; pos += len; }
        } else { ref_idx += len; read_idx += len; pos += len; }
    };
    int i = 0;
    while (i < n) {
        const uint32_t t = c[i];
        if (!indel_is_gap(t)) { token(t); ++i; continue; }
        int k = i + 1;
        while (k < n && indel_is_gap(c[k])) ++k;
        int merged = k - (i + 1);
        if (k == n) --merged;                              // run reaches the end of the CIGAR: its last token stays on its own
        if (merged < 0) merged = 0;
        bool hasI = false, hasD = false;
        for (int q = i; q <= i + merged; ++q) { if ((c[q] & 15u) == 1u) hasI = true; else hasD = true; }
        if (hasI && hasD && (t & 15u) == 2u) { for (int q = i + merged; q >= i; --q) token(c[q]); }
        else { for (int q = i; q <= i + merged; ++q) token(c[q]); }
        i += merged + 1;
    }
    read_end = read_idx;
    return cnt;
}

// AOS = true: alignments come from the context's own result records (after swb_compute); false: SoA arrays of the caller
template <bool AOS>
__global__ void k_indels(const swb_result* __restrict__ res, const uint32_t* __restrict__ cigar, const int64_t* __restrict__ c_off, const int32_t* __restrict__ c_len,
                         const int32_t* __restrict__ ref_start, const int32_t* __restrict__ read_start, int32_t n,
                         int64_t* out_off, int32_t* out_cnt, int32_t* out_read_end, swb_indel* recs, int64_t cap, unsigned long long* bump, int32_t* overflow)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int64_t off; int len, rs, qs;
    if (AOS) { const swb_result& r = res[p]; off = r.cigar_off; len = r.cigar_len; rs = r.ref_begin1; qs = r.read_begin1; }
    else { off = c_off[p]; len = c_len[p]; rs = ref_start[p]; qs = read_start[p]; }
    const uint32_t* c = cigar + off;
    int read_end = qs;
    const int cnt = len > 0 ? indel_walk(c, len, rs, qs, p, nullptr, read_end) : 0;
    const unsigned long long o = warp_bump(bump, (unsigned long long)cnt);
    out_off[p] = (int64_t)o; out_cnt[p] = cnt; out_read_end[p] = read_end;
    if (cnt == 0) return;
    if ((long long)o + cnt > cap) { atomicAdd(overflow, 1); return; }
    indel_walk(c, len, rs, qs, p, recs + o, read_end);
}
