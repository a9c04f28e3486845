// swb_bandwarp.cuh — banded_sw (ssw.c:588-772) for WIDE bands, one WARP per alignment.
//
// A free gap extension (indelPost's grid has gap_extension = 0 in 40 % of its calls) lets alignments span deletions of tens
// to hundreds of bases, so |refLen - readLen| + 1 = W reaches 25-150.  One thread per alignment then walks (2W+1) x readLen cells
// alone (k_band); here the 32 lanes share the row: lane l owns the columns [l*C, l*C + C) of the matrix (H of the previous row
// and the vertical-gap state of those columns in registers) and the rows run through the lanes as a pipeline -- lane l works on
// row s - l at step s and receives from lane l-1 the left neighbour's H and F and the old H of its last column (the diagonal
// neighbour of the first column here).  Every dependency of a cell points left or up, so a one-step skew is enough.
//
// Two kinds of jobs come here.  REGULAR ones (refLen >= 2W + 2: the band is narrower than the matrix): exactly as argued in
// swb_bandreg.cuh, the rolling-buffer algebra of the reference is then a plain banded recurrence with zero boundaries, so it can
// be written in column space.  And bands that never slide (W >= readLen - 1, so set_u's x is 0 in every row and slot = column + 1
// literally): there the only thing that differs from the plain recurrence is ssw.c:633 -- `edge` = min(end + 1, 2W + 2) is the
// slot of the LAST column once the band has reached it (end = refLen - 1 < 2W + 2), so from that row on the last column reads
// 0 for its upper neighbour's H and E.  Direction nibbles are stored by COLUMN (word j >> 3 of a 64-word row; a lane owns whole words),
// the traceback is walked by lane 0.  The band is doubled in place (ssw.c:668-669; the layout does not depend on W) while the
// doubled band is still one of the two kinds, and jobs the other band kernels widened into this kernel's range arrive with their
// width and running maximum in t_bw / t_best.  Anything else -- the walk leaving the band, a doubled band that slides and is wider
// than the matrix, scratch exhausted -- hands the job to the literal kernel with the state it needs to carry on exactly.
#pragma once
#include "swb_common.cuh"
#include "swb_band.cuh"
#include "swb_bandreg.cuh"

#define SWB_BANDWARP_C 16                      // columns per lane: windows of up to 512 columns
#define SWB_BANDWARP_ROWWORDS 64               // direction words per row (column layout)
#define SWB_BANDWARP_WARPS 4                   // warps (jobs) per block
#define SWB_BANDWARP_MAXREF (32 * SWB_BANDWARP_C)

__global__ void __launch_bounds__(32 * SWB_BANDWARP_WARPS, 6)      // latency bound: 24 warps per SM (<= 80 registers) beat 16
k_band_warp(SwbDev d, const int32_t* __restrict__ jobs, int njobs, int nextBase)
{
    constexpr int C = SWB_BANDWARP_C;
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ unsigned long long s_rowTab[8];            // per read base: its scores against every window base, 8 x int8
    if (threadIdx.x < 8) {
        unsigned long long tab = 0;
        if ((int)threadIdx.x < d.n) for (int nt = 0; nt < d.n; ++nt) tab |= (unsigned long long)(uint8_t)d.mat[nt * d.n + threadIdx.x] << (8 * nt);
        s_rowTab[threadIdx.x] = tab;
    }
    __syncthreads();
    __shared__ uint8_t s_read[SWB_BANDWARP_WARPS][SWB_BANDREG_MAXROWS];      // the job's read codes: one shared-memory load per row
    const int lane = threadIdx.x & 31;
    const int job = blockIdx.x * SWB_BANDWARP_WARPS + (threadIdx.x >> 5);
    if (job >= njobs) return;                              // whole warp
    const int p = jobs[job];
    swb_result& r = d.res[p];
    const int refLen = r.ref_end1 - r.ref_begin1 + 1;      // ssw.c:897-899
    const int readLen = r.read_end1 - r.read_begin1 + 1;
    const int dl = refLen - readLen;
    int W = (dl < 0 ? -dl : dl) + 1;
    int best = 0;
    if (d.t_bw[p] != 0) { W = d.t_bw[p]; best = d.t_best[p]; }      // re-queued by another kernel: its width and running maximum
    const int score = r.score1;
    const int go = d.gap_open[p], ge = d.gap_ext[p];
    const int len = refLen > readLen ? refLen : readLen;
    const int8_t* ref = d.windows + d.p_woff[p] + r.ref_begin1;
    const int8_t* read = d.reads + d.p_roff[p] + r.read_begin1;

    auto to_literal = [&](int bw, int best) {              // lane 0 only
        d.t_bw[p] = bw; d.t_best[p] = best;
        const int c = band_class(bw, refLen);
        list_push(d.list[nextBase + c], d.counters + nextBase + c, p);
    };

    for (int k = lane; k < readLen; k += 32) s_read[threadIdx.x >> 5][k] = (uint8_t)(read[k] & 7);
    __syncwarp();
    const uint8_t* rd = s_read[threadIdx.x >> 5];

    // scratch for the direction words (column layout)
    unsigned long long off = 0;
    const long long need = (long long)SWB_BANDWARP_ROWWORDS * 4 * readLen;
    if (lane == 0) off = atomicAdd(&d.bump[0], (unsigned long long)need);
    off = __shfl_sync(FULL, off, 0);
    if ((long long)off + need > d.band_cap) {
        if (lane == 0) { atomicAdd(d.counters + CNT_BAND_OVERFLOW, 1); to_literal(W, 0); }
        return;
    }
    uint32_t* dir = reinterpret_cast<uint32_t*>(d.band + off);

    // this lane's window columns as PRMT selectors (byte rc of the score row, sign-extended), 16 x 4 bits packed would need
    // unpacking per cell: keep the codes, 4 per register
    const int j0 = lane * C;
    uint32_t codes[C / 4];
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { const int j = j0 + 4 * q + k; if (j < refLen) w |= (uint32_t)(ref[j] & 7) << (8 * k); }
        codes[q] = w;
    }

    int bestIn;                                            // running maximum before the current width's pass
    for (;;) {                                             // band doubling, ssw.c:612-669 (the direction words do not depend on W)
    bestIn = best;
    // a band that never slides and is wider than the matrix: the last column loses its upper neighbour (see above)
    const bool cutLast = refLen < 2 * W + 2;
    int Hp[C], Ev[C];                                      // H of the previous row / vertical-gap state, per own column
#pragma unroll
    for (int k = 0; k < C; ++k) { Hp[k] = 0; Ev[k] = 0; }
    int outH = 0, outF = 0, outDiag = 0;                   // what lane+1 receives at the next step
    const int nsteps = readLen + 31;
    for (int s = 0; s < nsteps; ++s) {
        int inH = __shfl_up_sync(FULL, outH, 1), inF = __shfl_up_sync(FULL, outF, 1), inDiag = __shfl_up_sync(FULL, outDiag, 1);
        if (lane == 0) { inH = 0; inF = 0; inDiag = 0; }
        const int i = s - lane;                            // the row this lane is on
        if (i < 0 || i >= readLen) continue;
        const int beg = i - W > 0 ? i - W : 0;             // band of row i (ssw.c:630-631)
        const int end = i + W < refLen - 1 ? i + W : refLen - 1;
        outDiag = Hp[C - 1];                               // row i-1's value of this lane's last column: diagonal neighbour of lane+1's first
        if (end < j0 || beg >= j0 + C) { outH = 0; outF = 0; continue; }
        const unsigned long long tab = s_rowTab[rd[i]];
        const uint32_t tabLo = (uint32_t)tab, tabHi = (uint32_t)(tab >> 32);
        uint32_t words[C / 8];
#pragma unroll
        for (int q = 0; q < C / 8; ++q) words[q] = 0;
        int hLeft = inH, f = inF, hDiag = inDiag;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int j = j0 + k;
            int hUp = Hp[k], eUp = Ev[k];
            if (cutLast && j == refLen - 1) { hUp = 0; eUp = 0; }      // ssw.c:633 with edge == slot of the last column (end == refLen - 1 here)
            if (j >= beg && j <= end) {
                if (j == beg) { hLeft = 0; f = 0; if (beg == 0) hDiag = 0; }      // ssw.c:633: h_c[0] = f = 0; column -1 does not exist
                const uint32_t rc = (codes[k >> 2] >> (8 * (k & 3))) & 7u;
                const int sc = (int)prmt(tabLo, tabHi, bandreg_sel(rc, d.one));
                int fn = 0;
                const int h = bandreg_cell<false>(k & 7, hUp, eUp, hDiag, hLeft, f, fn, Ev[k], sc, go, ge, words[k >> 3]);
                Hp[k] = h;
                best = max(best, h);                       // ssw.c:661
                hLeft = h;
            }
            hDiag = hUp;                                   // row i-1's value of column j: diagonal neighbour of column j+1
        }
        const bool lastIn = j0 + C - 1 >= beg && j0 + C - 1 <= end;
        outH = lastIn ? hLeft : 0; outF = lastIn ? f : 0;
#pragma unroll
        for (int q = 0; q < C / 8; ++q) dir[(size_t)i * SWB_BANDWARP_ROWWORDS + lane * (C / 8) + q] = words[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
    __syncwarp();
    {
        long long cells = 0;                               // statistics: cells of the band inside the matrix
        for (int i = lane; i < readLen; i += 32) cells += min(i + W, refLen - 1) - max(i - W, 0) + 1;
        warp_count(d.counters + CNT_CELLS_BAND, (unsigned long long)cells);
    }
    if (!(best < score && W * 2 <= len)) break;            // ssw.c:668-669
    W *= 2;                                                // widen and redo: here while the band is still one of the two kinds
    if (!(refLen >= 2 * W + 2 || W >= readLen - 1)) { if (lane == 0) to_literal(W, best); return; }
    }
    if (lane != 0) return;

    // ---- traceback (ssw.c:672-751), lane 0, column layout; the word of the row above is fetched one step ahead ------------
    __threadfence_block();
    BandOps ops; ops.n = 0;
    int i = readLen - 1, j = refLen - 1, e = 0, state = 2, op = 0, prev_op = 0;
    uint32_t* out = nullptr; int total = 0;
    for (int pass = 0; pass < 2; ++pass) {
        i = readLen - 1; j = refLen - 1; e = 0; state = 2; op = 0; prev_op = 0; ops.n = 0;
        bool leftBand = false;
        while (i >= 0 && j > 0) {                          // ssw.c:679
            const int beg = i - W > 0 ? i - W : 0, end = i + W < refLen - 1 ? i + W : refLen - 1;
            if (j < beg || j > end) { leftBand = true; break; }
            const uint32_t w = __ldcg(dir + (size_t)i * SWB_BANDWARP_ROWWORDS + (j >> 3));
            const int b = (int)((w >> (4 * (j & 7))) & 15u);
            const int de = 2 + (b & 1), df = 4 + ((b >> 1) & 1);
            const int sel = (b >> 2) & 3;
            const int dv = state == 0 ? de : (state == 1 ? df : (sel == 0 ? 1 : (sel == 1 ? de : df)));
            switch (dv) {
                case 1: --i; --j; state = 2; op = 0; break;
                case 2: --i;      state = 0; op = 1; break;
                case 3: --i;      state = 2; op = 1; break;
                case 4: --j;      state = 1; op = 2; break;
                default: --j;     state = 2; op = 2; break;
            }
            if (op == prev_op) ++e;
            else { band_push(ops, ((uint32_t)e << 4) | (uint32_t)prev_op, out, total); prev_op = op; e = 1; }
        }
        if (leftBand) { to_literal(W, bestIn); return; }         // the reference's index arithmetic takes over (literal kernel, from scratch)
        if (op == 0) band_push(ops, ((uint32_t)(e + 1) << 4) | 0u, out, total);     // ssw.c:734-751
        else { band_push(ops, ((uint32_t)e << 4) | (uint32_t)op, out, total); band_push(ops, (1u << 4) | 0u, out, total); }
        if (pass == 1) break;
        const int l = ops.n;
        const unsigned long long coff = atomicAdd(&d.bump[1], (unsigned long long)l);
        r.cigar_len = l; r.cigar_off = (int64_t)coff;
        if ((long long)coff + l > d.cigar_cap) { atomicAdd(d.counters + CNT_CIGAR_OVERFLOW, 1); return; }
        if (l <= BAND_OPBUF) {
#pragma unroll
            for (int q = 0; q < BAND_OPBUF; ++q) if (q < l) d.cigar[coff + (l - 1 - q)] = ops.op[q];   // reverse (ssw.c:753-762)
            break;
        }
        out = d.cigar + coff; total = l;                   // more ops than the register buffer holds: walk again, writing in place
    }
    d.p_state[p] |= PST_BAND_DONE;
}
