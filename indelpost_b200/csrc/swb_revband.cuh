// swb_revband.cuh — reverse (begin-position) pass of the DPX fast path restricted to a provably sufficient band.
//
// ssw_align's reverse pass (ssw.c:875-886) re-runs the score kernel on the reversed read[0..read_end1] against
// ref[ref_end1..0] and stops at the first column whose maximum equals score1; ref_begin1 / read_begin1 are that
// column and the smallest row holding score1 in it.  For the pairs the fast path owns (striped result == plain
// Gotoh, see swb_fast.cuh) only cells holding exactly T = score1 matter, and T is the maximum of the whole forward
// matrix with (ref_end1, read_end1) the first column / smallest row attaining it.  Hence every alignment scoring T
// inside the reverse rectangle ends (forward view) exactly in that corner, i.e. in reversed coordinates it STARTS
// with the cell (0, 0).  Let Lr = read_end1 + 1 rows, mx = max(mat), B = mx*Lr - T >= 0.  A path that starts in
// (0, 0), scores T and at some point deviates from the main diagonal by
//     D columns to the right (net deletions)  pays >= go + (D-1)*ge            =>  D <= (B - go)/ge + 1
//     I rows downwards (net insertions)       pays that and loses mx per row   =>  I <= (B - go + ge)/(mx + ge)
// (go >= ge on the fast path, so several gaps cost at least as much as one).  Cells outside the band
// [-I, +D] around the main diagonal are on no such path; computing the recurrence with them treated as 0 gives
// values <= the true ones everywhere (all operations are monotone) and the true value wherever it equals T.
// So "first column holding T, smallest row in it" is unchanged.  Pad rows (ssw.c:169/391) score 0 and can only
// copy a T diagonally into a LATER column, so they are not computed at all.
//
// Work decomposition: one thread per lane pair (two alignments in the two s16x2 halves, same number format and
// score tables as swb_fast.cuh), the band row (H and the vertical-gap state V per diagonal, M = WI + WD + 1 of each)
// in registers, rows visited sequentially.  Diagonal d of row i is column j = i + d - WI; its diagonal neighbour
// is H[d] of the previous row, its upper neighbour H[d+1] / V[d+1], its left neighbour is carried in registers.
// Per-column PRMT selectors (both windows' bases) sit in shared memory as [column][thread] u16: every thread of a
// warp is on the same column, so the loads are conflict free.  6 ALU-pipe instructions per cell pair, no shuffles.
// Pairs are bucketed by band class in the forward sweep (k_fast, DIR == 0); wider bands and gap_extension == 0
// (no deletion bound) stay with the wavefront sweep k_fast<R,1>.
#pragma once
#include "swb_common.cuh"
#include "swb_fast.cuh"

#define SWB_REVB_THREADS 64

struct RevbPar { const int8_t* readA; const int8_t* readB; const int8_t* refA; const int8_t* refB; int LA, LB, nA, nB; };

// per-thread shared-memory region: [front pad + columns: u16 selectors][rows: u16 = codeA | codeB << 8], a whole number of
// 32-bit words and an ODD number of them, so the threads of a warp (all on the same column / row) hit distinct banks
__host__ __device__ __forceinline__ int revb_sel_cols(int rows, int M) { return ((M + 1) & ~1) + ((rows + M + 1) & ~1); }
__host__ __device__ __forceinline__ int revb_stride_words(int rows, int M) { return ((revb_sel_cols(rows, M) + ((rows + 1) & ~1)) / 2) | 1; }

template <int WI, int WD>
__global__ void __launch_bounds__(SWB_REVB_THREADS)
k_rev_band(SwbDev d, const int32_t* __restrict__ jobs, const int32_t* __restrict__ njobs_ptr, int rowsAlloc)
{
    constexpr int M = WI + WD + 1;
    constexpr int T = SWB_REVB_THREADS;
    constexpr int PF = (WI + 1) & ~1;                                    // even front pad: columns "left of 0" of the first rows
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_rowtab[SWB_MAX_N];
    const int npairs = *njobs_ptr;
    const int ngroups = (npairs + 1) >> 1;
    if (blockIdx.x * T >= ngroups) return;
    if (threadIdx.x < d.n) {
        uint32_t t = 0;
        for (int nt = 0; nt < 4; ++nt) t |= (uint32_t)(uint8_t)(int8_t)(FAST_SCALE * d.mat[nt * d.n + threadIdx.x]) << (8 * nt);
        s_rowtab[threadIdx.x] = t;
    }
    __syncthreads();
    const int grp = blockIdx.x * T + threadIdx.x;
    if (grp >= ngroups) return;

    // ---- the two alignments of this thread ---------------------------------------------------------
    const int pA = jobs[2 * grp], pB = (2 * grp + 1 < npairs) ? jobs[2 * grp + 1] : -1;
    const int qB = pB < 0 ? pA : pB;                                     // odd tail: lane B shadows lane A, its result is dropped
    const swb_result rA = d.res[pA], rB = d.res[qB];
    const int LA = rA.read_end1 + 1, LB = rB.read_end1 + 1;              // rows (ssw.c:875-877)
    const int nA = rA.ref_end1 + 1, nB = rB.ref_end1 + 1;                // columns
    uint32_t goP = pack2(FAST_SCALE * d.gap_open[pA], FAST_SCALE * d.gap_open[qB]);
    uint32_t ngeP = pack2(-FAST_SCALE * d.gap_ext[pA], -FAST_SCALE * d.gap_ext[qB]);
    const uint32_t targetV = pack2(FAST_SCALE * rA.score1 + FAST_C, pB < 0 ? 0x7fff : FAST_SCALE * rB.score1 + FAST_C);
    const int Lmax = max(LA, LB);

    // ---- staging: every thread copies what its band can touch of its two (reversed) windows and reads into its own
    //      shared-memory region.  Selector of column c sits at index PF + c.
    const int strideW = revb_stride_words(rowsAlloc, M);
    const int selCols = revb_sel_cols(rowsAlloc, M);
    uint32_t* const region = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)threadIdx.x * strideW;
    {
        uint16_t* selW = reinterpret_cast<uint16_t*>(region);
        uint16_t* rowW = reinterpret_cast<uint16_t*>(region + selCols / 2);
        const int ncols = min(Lmax + M + 1, selCols - PF);               // columns the band can reach
        for (int c = 0; c < PF + ncols; ++c) selW[c] = 0xC480u;          // base A / A: columns left of 0 and right of the windows
        const int kA = min(nA, ncols), kB = min(nB, ncols);              // reversed column c = window position n - 1 - c
        for_each_byte16_pair(d.windows + d.p_woff[pA] + (nA - kA), kA, [&](int i, uint32_t v) { selW[PF + kA - 1 - i] |= (uint16_t)((v & 3u) * 0x11u); },
                             d.windows + d.p_woff[qB] + (nB - kB), kB, [&](int i, uint32_t v) { selW[PF + kB - 1 - i] |= (uint16_t)((v & 3u) * 0x1100u); });
        for (int i = 0; i < Lmax; ++i) rowW[i] = 0;
        for_each_byte16_pair(d.reads + d.p_roff[pA], LA, [&](int i, uint32_t v) { rowW[LA - 1 - i] |= (uint16_t)(v & 0xffu); },
                             d.reads + d.p_roff[qB], LB, [&](int i, uint32_t v) { rowW[LB - 1 - i] |= (uint16_t)((v & 0xffu) << 8); });
    }
    const uint16_t* selT = reinterpret_cast<const uint16_t*>(region) + (PF - WI);
    const uint16_t* rowT = reinterpret_cast<const uint16_t*>(region + selCols / 2);

    uint32_t H[M], V[M];
#pragma unroll
    for (int k = 0; k < M; ++k) { H[k] = FAST_CPACK; V[k] = FAST_CPACK; }
    asm volatile("" : "+r"(goP), "+r"(ngeP));

    int bestColA = 0x7fffffff, bestRowA = 0, bestColB = 0x7fffffff, bestRowB = 0;
    uint32_t rw = rowT[0];                                              // software prefetch of the next row's read bases

    for (int i = 0; i < Lmax; ++i) {
        const uint32_t tA = s_rowtab[rw & 0xffu], tB = s_rowtab[rw >> 8];
        rw = rowT[min(i + 1, Lmax - 1)];
        const uint16_t* sp = selT + i;                                   // selector of diagonal k: sp[k]
        uint32_t G = FAST_CPACK, rm = 0;
        if (i < WI) {
            // first rows: diagonals left of column 0 do not exist; keep them at the zero level
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const bool ok = k >= WI - i;
                const uint32_t s = prmt(tA, tB, sp[k]);
                const uint32_t vin = k + 1 < M ? V[k + 1] : FAST_CPACK;
                uint32_t h = __viaddmax_s16x2(H[k], s, vin);
                h = __vimax3_s16x2(h, G, FAST_CPACK);
                const uint32_t hg = h - goP;
                const uint32_t v = __viaddmax_s16x2(vin, ngeP, hg);
                const uint32_t g2 = __viaddmax_s16x2(G, ngeP, hg);
                H[k] = ok ? h : FAST_CPACK; V[k] = ok ? v : FAST_CPACK; G = ok ? g2 : FAST_CPACK;
                if (ok) rm = vmax2(rm, h);
            }
        } else {
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const uint32_t s = prmt(tA, tB, sp[k]);
                const uint32_t vin = k + 1 < M ? V[k + 1] : FAST_CPACK;
                const uint32_t t = __viaddmax_s16x2(H[k], s, vin);           // max(Hdiag + s, V)
                const uint32_t h = __vimax3_s16x2(t, G, FAST_CPACK);         // max(., G, 0)
                H[k] = h;
                const uint32_t hg = h - goP;                                 // no lane borrow: h >= 0x4000 > 16*go
                V[k] = __viaddmax_s16x2(vin, ngeP, hg);                      // vertical-gap state of the cell below
                // horizontal-gap state of the cell to the right: max(G - ge, h - go) = max(G - ge, max(t, 0) - go) because
                // G - go <= G - ge (go > ge on the fast path): the row's serial chain is ONE instruction per cell, h is off it
                G = __viaddmax_s16x2(G, ngeP, vmax2(t, FAST_CPACK) - goP);
                rm = vmax2(rm, h);
            }
        }
        // Cells of the matrix never exceed score1, but the band also runs over columns / rows beyond a lane's own
        // rectangle (they depend on nothing inside it and nothing inside depends on them), whose values are arbitrary:
        // look closer whenever the row maximum reaches score1.
        const uint32_t x = vmax2(rm, targetV) ^ rm;
        if ((x & 0xffffu) == 0 || (x >> 16) == 0) {
            // smallest column first, rows are visited in ascending order
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const uint32_t y = H[k] ^ targetV;
                const int j = i + k - WI;
                if ((y & 0xffffu) == 0 && j >= 0 && j < nA && i < LA && j < bestColA) { bestColA = j; bestRowA = i; }
                if ((y >> 16) == 0 && j >= 0 && j < nB && i < LB && j < bestColB) { bestColB = j; bestRowB = i; }
            }
        }
    }

    // ---- results (ssw.c:885-891) ----------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int p = s ? pB : pA;
        if (p < 0) continue;
        const int hc = s ? bestColB : bestColA, hr = s ? bestRowB : bestRowA;
        swb_result& r = d.res[p];
        if (hc == 0x7fffffff) {
            // no cell reaches score1: let the exact path reproduce ssw.c literally (flag = 2 case)
            const int md = d.p_mode[p];
            d.p_state[p] &= ~PST_FAST;
            list_push(d.list[md ? LIST_WORD_REV2 : LIST_BYTE_REV2], d.counters + (md ? LIST_WORD_REV2 : LIST_BYTE_REV2), p);
        } else {
            warp_count(d.counters + CNT_CELLS_REV, (unsigned long long)(s ? LB : LA) * (hc + 1));
            r.ref_begin1 = r.ref_end1 - hc;
            r.read_begin1 = r.read_end1 - hr;
            const int f = d.flag;
            const bool noCigar = (7 & f) == 0 || ((2 & f) != 0 && (int)r.score1 < (int)d.filters) ||
                                 ((4 & f) != 0 && (r.ref_end1 - r.ref_begin1 > d.filterd || r.read_end1 - r.read_begin1 > d.filterd));
            if (!noCigar) push_band(d, p, r); else d.p_state[p] |= PST_BAND_DONE;
        }
    }
}
