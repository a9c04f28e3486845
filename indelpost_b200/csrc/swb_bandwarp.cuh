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
//
// Q = 8 (round 2): bands of half-width 25 .. 51 keep only 4-5 of a warp's 32 column blocks busy, so FOUR alignments share a warp,
// eight lanes each.  The 16-column blocks are dealt to the lanes cyclically (block b -> lane b mod 8) and block b still works on
// row s - b at step s, so the pipeline is the same; a band of 2W + 1 <= 103 columns touches at most seven consecutive blocks in a
// row, so at any step a lane has at most one block in the band (its registers are reset when it moves on to block b + 8: the cells
// above a block's first band row are out of band, i.e. zero).  A lane stays with a block for one step after the block's last column
// has left the band, because that step hands H(i-1, last column) -- the diagonal neighbour of the band's first cell in row i -- to
// the next lane.  Only regular bands (refLen >= 2W + 2) come here; the direction words keep the column layout, so the traceback is
// the same walk.
#pragma once
#include "swb_common.cuh"
#include "swb_band.cuh"
#include "swb_bandreg.cuh"

#define SWB_BANDWARP_C 16                      // columns per lane: windows of up to 512 columns
#define SWB_BANDWARP_ROWWORDS 64               // direction words per row (column layout)
#define SWB_BANDWARP_WARPS 4                   // warps (jobs) per block
#define SWB_BANDWARP_MAXREF (32 * SWB_BANDWARP_C)
#define SWB_BANDQ_MAXW 51                     // widest half-width of the eight-lanes-per-alignment schedule (see below)

// jobs whose band fits the eight-lane schedule
__device__ __forceinline__ bool bandq_ok(int bw, int refLen) { return bw <= SWB_BANDQ_MAXW && refLen >= 2 * bw + 2; }

// splits a warp-kernel job list by schedule: eight lanes per alignment (listQ) / a whole warp (listW)
__global__ void k_bandwarp_split(SwbDev d, const int32_t* __restrict__ jobs, int njobs, int listQ, int listW)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= njobs) return;
    const int p = jobs[t];
    const swb_result& r = d.res[p];
    const int refLen = r.ref_end1 - r.ref_begin1 + 1, readLen = r.read_end1 - r.read_begin1 + 1;
    const int dl = refLen - readLen;
    const int W = d.t_bw[p] != 0 ? d.t_bw[p] : (dl < 0 ? -dl : dl) + 1;
    const int dst = bandq_ok(W, refLen) ? listQ : listW;
    list_push(d.list[dst], d.counters + dst, p);
}

template <int Q>
__global__ void __launch_bounds__(32 * SWB_BANDWARP_WARPS, 6)      // latency bound: 24 warps per SM (<= 80 registers) beat 16
k_band_warp(SwbDev d, const int32_t* __restrict__ jobs, int njobsMax, const int32_t* __restrict__ njobs_ptr, int nextBase)
{
    constexpr int C = SWB_BANDWARP_C;
    constexpr int JPW = 32 / Q;                            // alignments per warp
    __shared__ unsigned long long s_rowTab[8];            // per read base: its scores against every window base, 8 x int8
    if (threadIdx.x < 8) {
        unsigned long long tab = 0;
        if ((int)threadIdx.x < d.n) for (int nt = 0; nt < d.n; ++nt) tab |= (unsigned long long)(uint8_t)d.mat[nt * d.n + threadIdx.x] << (8 * nt);
        s_rowTab[threadIdx.x] = tab;
    }
    __syncthreads();
    __shared__ uint8_t s_read[SWB_BANDWARP_WARPS * JPW][SWB_BANDREG_MAXROWS];      // the job's read codes: one shared-memory load per row
    const int njobs = njobs_ptr ? min(njobsMax, *njobs_ptr) : njobsMax;
    const int lane = threadIdx.x & 31;
    const int ql = lane % Q;                               // lane within the alignment's group
    const int grp = lane / Q;
    const unsigned GM = Q == 32 ? 0xffffffffu : (((1u << Q) - 1u) << (grp * Q));
    const int slot = (threadIdx.x >> 5) * JPW + grp;
    const int job = (blockIdx.x * SWB_BANDWARP_WARPS + (threadIdx.x >> 5)) * JPW + grp;
    if (job >= njobs) return;                              // whole group
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

    auto to_literal = [&](int bw, int best) {              // group lane 0 only
        d.t_bw[p] = bw; d.t_best[p] = best;
        const int c = band_class(bw, refLen);
        list_push(d.list[nextBase + c], d.counters + nextBase + c, p);
    };
    auto to_next = [&](int bw, int best) {                 // group lane 0 only: warp-per-alignment kernel of the next round, else literal
        d.t_bw[p] = bw; d.t_best[p] = best;
        requeue_band(d, nextBase, p, bw, refLen, readLen, r);
    };

    for (int k = ql; k < readLen; k += Q) s_read[slot][k] = (uint8_t)(read[k] & 7);
    __syncwarp(GM);
    const uint8_t* rd = s_read[slot];

    // scratch for the direction words (column layout)
    unsigned long long off = 0;
    const long long need = (long long)SWB_BANDWARP_ROWWORDS * 4 * readLen;
    if (ql == 0) off = atomicAdd(&d.bump[0], (unsigned long long)need);
    off = __shfl_sync(GM, off, grp * Q);
    if ((long long)off + need > d.band_cap) {
        if (ql == 0) { atomicAdd(d.counters + CNT_BAND_OVERFLOW, 1); to_literal(W, 0); }
        return;
    }
    uint32_t* dir = reinterpret_cast<uint32_t*>(d.band + off);

    // a block's window columns as codes, 4 per register (PRMT selectors are rebuilt per cell: bandreg_sel)
    uint32_t codes[C / 4];
    auto load_codes = [&](int b) {
        const int j0 = b * C;
#pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const int j = j0 + 4 * q + k; if (j < refLen) w |= (uint32_t)(ref[j] & 7) << (8 * k); }
            codes[q] = w;
        }
    };
    if constexpr (Q == 32) load_codes(ql);
    const int nblocks = (refLen + C - 1) / C;

    int bestIn;                                            // running maximum before the current width's pass
    for (;;) {                                             // band doubling, ssw.c:612-669 (the direction words do not depend on W)
    bestIn = best;
    // a band that never slides and is wider than the matrix: the last column loses its upper neighbour (see above)
    const bool cutLast = Q == 32 && refLen < 2 * W + 2;
    int Hp[C], Ev[C];                                      // H of the previous row / vertical-gap state, per own column
#pragma unroll
    for (int k = 0; k < C; ++k) { Hp[k] = 0; Ev[k] = 0; }
    int cb = Q == 32 ? ql : -1;                            // block this lane's registers hold
    int outH = 0, outF = 0, outDiag = 0;                   // what the next lane receives at the next step
    const int nsteps = readLen + (Q == 32 ? 31 : nblocks);
    for (int s = 0; s < nsteps; ++s) {
        // left neighbour = the lane of block b - 1 (cyclic for Q = 8)
        const int src = grp * Q + (ql + Q - 1) % Q;
        int inH = __shfl_sync(GM, outH, src), inF = __shfl_sync(GM, outF, src), inDiag = __shfl_sync(GM, outDiag, src);
        int b;
        if constexpr (Q == 32) b = ql;
        else {
            // the (at most eight) consecutive blocks that can hold band cells of their row at this step, plus the one-step hand-over
            const int t17 = s - W - 16;
            const int lo = t17 >= 0 ? (t17 + 16) / 17 : -((-t17) / 17);      // ceil(t17 / 17)
            b = lo + (((ql - lo) % Q) + Q) % Q;
        }
        if (b == 0) { inH = 0; inF = 0; inDiag = 0; }
        const int i = s - b;                               // the row block b is on
        if (b < 0 || b >= nblocks || i < 0 || i >= readLen) { if (Q != 32) { outH = 0; outF = 0; outDiag = 0; } continue; }
        if (Q != 32 && b != cb) {                          // moved on to the next block of this lane: nothing above it is in the band yet
#pragma unroll
            for (int k = 0; k < C; ++k) { Hp[k] = 0; Ev[k] = 0; }
            load_codes(b);
            cb = b;
        }
        const int j0 = b * C;
        const int beg = i - W > 0 ? i - W : 0;             // band of row i (ssw.c:630-631)
        const int end = i + W < refLen - 1 ? i + W : refLen - 1;
        outDiag = Hp[C - 1];                               // row i-1's value of this block's last column: diagonal neighbour of the next block's first
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
        for (int q = 0; q < C / 8; ++q) dir[(size_t)i * SWB_BANDWARP_ROWWORDS + b * (C / 8) + q] = words[q];
    }
#pragma unroll
    for (int o = Q / 2; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(GM, best, o, Q));
    __syncwarp(GM);
    {
        long long cells = 0;                               // statistics: cells of the band inside the matrix
        for (int i = ql; i < readLen; i += Q) cells += min(i + W, refLen - 1) - max(i - W, 0) + 1;
        warp_count(d.counters + CNT_CELLS_BAND, (unsigned long long)cells);
    }
    if (!(best < score && W * 2 <= len)) break;            // ssw.c:668-669
    W *= 2;                                                // widen and redo: here while the band still fits this schedule
    if (Q == 32) { if (!(refLen >= 2 * W + 2 || W >= readLen - 1)) { if (ql == 0) to_literal(W, best); return; } }
    else if (!bandq_ok(W, refLen)) { if (ql == 0) to_next(W, best); return; }
    }
    if (ql != 0) return;

    // ---- traceback (ssw.c:672-751), group lane 0, column layout -----------------------------------------------------------
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
