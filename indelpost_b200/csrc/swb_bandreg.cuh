// swb_bandreg.cuh — banded_sw (ssw.c:588-772) with the rolling band rows in REGISTERS, for the first band width of
// the "regular" jobs: half-width W = |refLen - readLen| + 1 <= SWB_BANDW_MAX known at compile time, refLen >= 2W + 2
// (the band is narrower than the matrix, so none of the reference's buffer-reuse quirks is reachable, see below),
// matrix edge n <= 8.  Band doubling (ssw.c:668-669) continues in place through the 2W, 4W ... instantiations while the band stays
// regular and within SWB_BANDW_MAX (bandreg_solve).  Everything else — wider or irregular bands — goes to the warp-per-alignment
// kernel or the literal kernel k_band (swb_band.cuh), which also provides the traceback walk used here.
//
// Slot algebra (ssw.c:92-95, set_u): cell (i, j) of row i lives in slot u = j - max(0, i - W) + 1 of the rolling
// buffers; x = u - 1 below.
//   rows i <= W      : x = j.            upper neighbour (i-1, j) = slot x, diagonal neighbour = slot x-1 (0 for x = 0)
//   rows i >= W + 1  : x = j - (i - W).  upper neighbour = slot x+1, diagonal neighbour = slot x
// so the band is a fixed array of 2W+1 registers per buffer updated in place, and the position of a cell's direction
// nibble inside its row (set_d) is x in both regimes.  Slots the reference clears at a row start (ssw.c:633: 0 and
// `edge` = the slot right of the row's last cell) are exactly the out-of-band neighbours: constants 0 here.  With
// refLen >= 2W+2 a row is clipped by the matrix edge only for i >= W+1, its `edge` slot (2W+2) is then not read, and
// the slots a later row reads are always ones the copy loop ssw.c:666 refreshed; cells right of the matrix edge are
// computed (nothing valid depends on them) but excluded from the running maximum.
// One thread per alignment; jobs are listed per exact W so a warp runs one instantiation.  Substitution scores:
// one PRMT per cell from the read base's 8-byte score row and a per-column selector kept in shared memory
// ([thread][column], odd word stride = conflict free), staged by the warp with coalesced loads.
#pragma once
#include "swb_common.cuh"
#include "swb_band.cuh"
#include "swb_fast.cuh"

#define SWB_BANDREG_THREADS 64
#define SWB_BANDREG_INPLACE 1          // band doublings a register-band kernel performs itself before it re-queues the job ...
#define SWB_BANDREG_INPLACE_MAXW 14    // ... up to this doubled half-width (a kernel needs the registers of its widest band: wider chains would halve the occupancy of the W = 9 .. 12 kernels)

struct BandRegPar { const int8_t* read; const int8_t* ref; int readLen, refLen; };

// per-thread shared-memory region: [columns: u8 window codes][rows: u8 read codes], odd number of 32-bit words (the PRMT
// selector of a column is rebuilt from its code with one IMAD on the FMA pipe: half the bytes of a u16 selector array, and
// shared memory is what limits the occupancy of this kernel)
__host__ __device__ __forceinline__ int bandreg_sel_cols(int rows) { return (rows + SWB_BANDW_MAX + 2 * SWB_BANDW_MAX + 4 + 3) & ~3; }
__host__ __device__ __forceinline__ int bandreg_stride_words(int rows) { return ((bandreg_sel_cols(rows) + ((rows + 3) & ~3)) / 4) | 1; }

// PRMT selector of a window code: byte rc of the score row, sign-extended to 32 bits (`one` is 1 but opaque to the
// compiler, so this stays a real IMAD on the FMA pipe)
__device__ __forceinline__ uint32_t bandreg_sel(uint32_t rc, int one) { return rc * (uint32_t)(0x1111 * one) + 0x8880u; }

// one DP cell (ssw.c:637-664).  hUp/eUp: upper neighbour, hDiag: diagonal neighbour, hLeft/f: left neighbour state.
// Returns H; writes E back through eOut; ORs the 4-bit direction code (bit0: E opened from H, bit1: F opened from H,
// bits 2-3: 0 diagonal / 1 E / 2 F) into `word` at nibble NIB (a constant after unrolling).
// CHAIN: gap_open >= gap_extension, the usual case: the F value handed to the next cell, max(f - ge, h - go), then equals
// max(f - ge, max(e1, m) - go) -- h - go adds only f1 - go <= max(f - ge, -go) and max(e1, m) >= 0 -- so the row's serial
// dependency is ONE add-max per cell and everything else (H, the direction bits) hangs off it.  fNext carries that value; the
// F of THIS cell (f) is still formed exactly as the reference does, from hLeft and the previous f, for the direction bit.
template <bool CHAIN>
__device__ __forceinline__ int bandreg_cell(const int NIB, int hUp, int eUp, int hDiag, int hLeft, int& f, int& fNext, int& eOut, int sc, int go, int ge, uint32_t& word)
{
    int a = hUp - go, b = eUp - ge;                       // ssw.c:644-648
    const int ev = a > b ? a : b;
    uint32_t bits = a > b ? (1u << (4 * NIB)) : 0u;
    eOut = ev;
    a = hLeft - go; b = f - ge;                           // ssw.c:650-653
    const int fv = CHAIN ? fNext : (a > b ? a : b);
    bits |= a > b ? (2u << (4 * NIB)) : 0u;
    const int e1 = ev > 0 ? ev : 0, f1 = fv > 0 ? fv : 0; // ssw.c:655-659
    const int gmax = e1 > f1 ? e1 : f1;
    const int m = hDiag + sc;
    const int h = gmax > m ? gmax : m;
    const uint32_t selv = gmax <= m ? 0u : (e1 > f1 ? (4u << (4 * NIB)) : (8u << (4 * NIB)));   // ssw.c:663-664
    word |= bits | selv;
    if (CHAIN) { const int w = (e1 > m ? e1 : m) - go; const int x = fv - ge; fNext = x > w ? x : w; }
    f = fv;
    return h;
}

// Traceback walk (ssw.c:672-751) for the regular in-band case, reading the packed direction rows through a window of
// the thread's shared-memory region (the selectors are dead by now): the window is refilled with independent loads
// whenever the walk climbs above it, so the walk pays one L2 round trip per WROWS rows instead of one per row.
// Returns the op count, -1 for the reference's "Trace back error", -2 if the walk leaves the cells this kernel
// computed (the caller then repeats it with the literal index arithmetic of band_traceback).
template <int NW>
__device__ __forceinline__ int bandreg_traceback(const uint32_t* __restrict__ dir, const BandGeom& g, uint32_t* win, const int wrows,
                                                 BandOps& ops, uint32_t* out, int total)
{
    int i = g.readLen - 1, j = g.refLen - 1;
    int e = 0, state = 2;
    int op = 0, prev_op = 0;                           // 0 M, 1 I, 2 D  (BAM op codes)
    ops.n = 0;
    int winBase = 0x7fffffff;                          // first row held in the window
    while (i >= 0 && j > 0) {                          // ssw.c:679
        const int x = j - band_x(g.w, i);
        if (x < 0 || x >= g.width_d || j > g.end(i)) return -2;
        if (i < winBase) {
            winBase = max(0, i - (wrows - 1));
            const int nwords = (i - winBase + 1) * NW;
            const uint32_t* src = dir + (size_t)winBase * NW;
#pragma unroll 8
            for (int k = 0; k < nwords; ++k) win[k] = src[k];
        }
        const uint32_t w = win[(i - winBase) * NW + (NW > 1 ? (x >> 3) : 0)];
        const int b = (int)((w >> (4 * (x & 7))) & 15u);
        const int de = 2 + (b & 1), df = 4 + ((b >> 1) & 1);
        const int sel = (b >> 2) & 3;
        const int dv = state == 0 ? de : (state == 1 ? df : (sel == 0 ? 1 : (sel == 1 ? de : df)));
        switch (dv) {
            case 1: --i; --j; state = 2; op = 0; break;
            case 2: --i;      state = 0; op = 1; break;
            case 3: --i;      state = 2; op = 1; break;
            case 4: --j;      state = 1; op = 2; break;
            default: --j;     state = 2; op = 2; break;     // 5 (every cell of a regular band row was written: no other value occurs)
        }
        if (op == prev_op) ++e;
        else {
            band_push(ops, ((uint32_t)e << 4) | (uint32_t)prev_op, out, total);
            prev_op = op; e = 1;
        }
    }
    if (op == 0) {                                     // ssw.c:734-751
        band_push(ops, ((uint32_t)(e + 1) << 4) | 0u, out, total);
    } else {
        band_push(ops, ((uint32_t)e << 4) | (uint32_t)op, out, total);
        band_push(ops, (1u << 4) | 0u, out, total);
    }
    return ops.n;
}

// The DP of one band width (ssw.c:612-667) over the staged sequences; returns the running maximum, writes the packed
// direction rows to `dir`.  CHAIN: see bandreg_cell.
template <int W, bool CHAIN>
__device__ __forceinline__ int bandreg_dp(const BandGeom& g, const uint8_t* selT, const uint8_t* rowT, const unsigned long long* s_rowTab,
                                          const int go, const int ge, const int one, uint32_t* __restrict__ dir, int best, long long& cells)
{
    constexpr int NX = 2 * W + 1;
    constexpr int NW = (NX + 7) / 8;                      // direction words per row
    int Hs[NX], Es[NX];
#pragma unroll
    for (int x = 0; x < NX; ++x) { Hs[x] = 0; Es[x] = 0; }

    // ---- rows 0 .. W: slot x = column x, cells x <= i + W ---------------------------------------------------------
    const int rowsA = min(W + 1, g.readLen);
    for (int i = 0; i < rowsA; ++i) {
        const unsigned long long tab = s_rowTab[rowT[i]];
        const uint32_t tabLo = (uint32_t)tab, tabHi = (uint32_t)(tab >> 32);
        uint32_t words[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) words[k] = 0;
        int f = 0, fNext = CHAIN ? -ge : 0, hLeft = 0, hDiag = 0;     // ssw.c:633: f = h_c[0] = 0, so the first cell's F is max(-go, -ge)
        const int lim = i + W;                              // regular jobs: lim <= 2W <= refLen - 2, no clipping here
#pragma unroll
        for (int x = 0; x < NX; ++x) {
            if (x <= lim) {
                const int hUp = Hs[x], eUp = Es[x];
                const int sc = (int)prmt(tabLo, tabHi, bandreg_sel(selT[x], one));
                const int h = bandreg_cell<CHAIN>(x & 7, hUp, eUp, hDiag, hLeft, f, fNext, Es[x], sc, go, ge, words[x >> 3]);
                Hs[x] = h;
                best = max(best, h);                        // ssw.c:661
                hDiag = hUp; hLeft = h;
            }
        }
#pragma unroll
        for (int k = 0; k < NW; ++k) dir[(size_t)i * NW + k] = words[k];
        cells += lim + 1;
    }

    // ---- rows W+1 .. readLen-1: slot x = column x + i - W --------------------------------------------------------
    for (int i = W + 1; i < g.readLen; ++i) {
        const unsigned long long tab = s_rowTab[rowT[i]];
        const uint32_t tabLo = (uint32_t)tab, tabHi = (uint32_t)(tab >> 32);
        const uint8_t* sp = selT + (i - W);
        uint32_t words[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) words[k] = 0;
        int f = 0, fNext = CHAIN ? -ge : 0, hLeft = 0;
        const int xmax = g.refLen - 1 - (i - W);            // last slot inside the matrix (>= 2W: row not clipped)
#pragma unroll
        for (int x = 0; x < NX; ++x) {
            const int hUp = x + 1 < NX ? Hs[x + 1] : 0, eUp = x + 1 < NX ? Es[x + 1] : 0;
            const int sc = (int)prmt(tabLo, tabHi, bandreg_sel(sp[x], one));
            const int h = bandreg_cell<CHAIN>(x & 7, hUp, eUp, Hs[x], hLeft, f, fNext, Es[x], sc, go, ge, words[x >> 3]);
            Hs[x] = h;
            if (x <= xmax) best = max(best, h);
            hLeft = h;
        }
#pragma unroll
        for (int k = 0; k < NW; ++k) dir[(size_t)i * NW + k] = words[k];
        cells += min(xmax, 2 * W) + 1;
    }
    return best;
}

// One band width for one staged job: DP, then either the traceback, or -- the running maximum still below score1 (ssw.c:668-669) --
// the doubled width right here while it is still a regular register band (the staging does not depend on W), so that a widened job
// costs its own thread a second DP instead of the whole batch another round of latency-bound launches.  Widths beyond the register
// kernels go to the warp-per-alignment / literal kernels of the next round.
template <int W, int DEPTH = 0>
__device__ __forceinline__ void bandreg_solve(const SwbDev& d, const int p, swb_result& r, BandGeom g, const uint8_t* selT, const uint8_t* rowT,
                                              const unsigned long long* s_rowTab, uint32_t* region, const int strideW, const int go, const int ge,
                                              const int score, const int best0, const int nextBase, const int nextBaseW)
{
    constexpr int NX = 2 * W + 1;
    constexpr int NW = (NX + 7) / 8;                      // direction words per row
    g.w = W; g.width_d = NX; g.strideW = NW;
    const int len = g.refLen > g.readLen ? g.refLen : g.readLen;
    // ---- scratch for the packed direction words -----------------------------------------------------------------
    const long long need = (long long)NW * 4 * g.readLen;
    const unsigned long long off = warp_bump(&d.bump[0], (unsigned long long)need);
    if ((long long)off + need > d.band_cap) {               // out of scratch: the literal kernel retries in a later round
        d.t_bw[p] = W; d.t_best[p] = best0;
        atomicAdd(d.counters + CNT_BAND_OVERFLOW, 1);
        const int c = band_class(W, g.refLen);
        list_push(d.list[nextBase + c], d.counters + nextBase + c, p);
        return;
    }
    uint32_t* dir = reinterpret_cast<uint32_t*>(d.band + off);

    long long cells = 0;
    // a widened job carries its running maximum along (ssw.c:661 is not reset)
    const int best = go >= ge ? bandreg_dp<W, true>(g, selT, rowT, s_rowTab, go, ge, d.one, dir, best0, cells)
                              : bandreg_dp<W, false>(g, selT, rowT, s_rowTab, go, ge, d.one, dir, best0, cells);
    warp_count(d.counters + CNT_CELLS_BAND, (unsigned long long)cells);

    if (best < score && W * 2 <= len) {                     // ssw.c:668-669: widen and redo
        // (one doubling in place: a kernel's register count is that of its widest band, and a second doubling is rare)
        if constexpr (2 * W <= SWB_BANDREG_INPLACE_MAXW && DEPTH < SWB_BANDREG_INPLACE) {
            if (g.refLen >= 4 * W + 2) {                    // still narrower than the matrix: the 2W register band, in place
                bandreg_solve<2 * W, DEPTH + 1>(d, p, r, g, selT, rowT, s_rowTab, region, strideW, go, ge, score, best, nextBase, nextBaseW);
                return;
            }
        }
        d.t_bw[p] = 2 * W; d.t_best[p] = best;
        if (nextBaseW >= 0 && 2 * W <= SWB_BANDW_MAX && g.refLen >= 4 * W + 2) {
            list_push(d.list[nextBaseW + 2 * W - 1], d.counters + nextBaseW + 2 * W - 1, p);      // still regular: this kernel's 2W instantiation, next round
        } else {
            requeue_band(d, nextBase, p, 2 * W, g.refLen, g.readLen, r);                            // warp-per-alignment or literal kernel
        }
        return;
    }

    // ---- traceback: one walk (ops buffered in registers), allocate, emit reversed ---------------------------------
    BandOps ops;
    constexpr bool regRows = NW <= BAND_ROWWORDS;
    const int wrows = min(32, (strideW - 1) / NW);
    int l = bandreg_traceback<NW>(dir, g, region, wrows, ops, nullptr, 0);
    const bool literal = l == -2;
    if (literal) l = band_traceback<regRows>(dir, g, ops, nullptr, 0);
    if (l < 0) { r.flag = 1; r.cigar_len = 0; r.cigar_off = 0; d.p_state[p] |= PST_BAND_DONE; return; }      // ssw.c:911
    const unsigned long long coff = warp_bump(&d.bump[1], (unsigned long long)l);
    r.cigar_len = l; r.cigar_off = (int64_t)coff;
    if ((long long)coff + l > d.cigar_cap) { atomicAdd(d.counters + CNT_CIGAR_OVERFLOW, 1); return; }
    if (l <= BAND_OPBUF) {
#pragma unroll
        for (int q = 0; q < BAND_OPBUF; ++q) if (q < l) d.cigar[coff + (l - 1 - q)] = ops.op[q];   // reverse (ssw.c:753-762)
    } else if (literal) {
        band_traceback<regRows>(dir, g, ops, d.cigar + coff, l);
    } else {
        bandreg_traceback<NW>(dir, g, region, wrows, ops, d.cigar + coff, l);
    }
    d.p_state[p] |= PST_BAND_DONE;
}

template <int W>
__global__ void __launch_bounds__(SWB_BANDREG_THREADS)
k_band_reg(SwbDev d, const int32_t* __restrict__ jobs, int njobs, int nextBase, int nextBaseW, int resume, int rowsAlloc)
{
    constexpr int NX = 2 * W + 1;
    constexpr int T = SWB_BANDREG_THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_rowTab[8];            // per read base: its scores against every window base, 8 x int8
    if (threadIdx.x < 8) {
        unsigned long long tab = 0;
        if ((int)threadIdx.x < d.n) for (int nt = 0; nt < d.n; ++nt) tab |= (unsigned long long)(uint8_t)d.mat[nt * d.n + threadIdx.x] << (8 * nt);
        s_rowTab[threadIdx.x] = tab;
    }
    __syncthreads();
    const int t = blockIdx.x * T + threadIdx.x;
    if (t >= njobs) return;
    const int p = jobs[t];
    BandGeom g; g.w = W; g.width_d = NX; g.strideW = (NX + 7) / 8;
    {
        const swb_result& r0 = d.res[p];
        g.refLen = r0.ref_end1 - r0.ref_begin1 + 1;       // ssw.c:897-899
        g.readLen = r0.read_end1 - r0.read_begin1 + 1;
    }

    // ---- staging: every thread copies its own window / read segment into its shared-memory region ------------------
    // (the region does not depend on W: selectors for every column a band of up to SWB_BANDW_MAX can touch)
    const int strideW = bandreg_stride_words(rowsAlloc);
    const int selCols = bandreg_sel_cols(rowsAlloc);
    uint32_t* const region = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)threadIdx.x * strideW;
    uint8_t* selW = reinterpret_cast<uint8_t*>(region);
    uint8_t* rowW = reinterpret_cast<uint8_t*>(region + selCols / 4);
    {
        const swb_result& r0 = d.res[p];
        for_each_byte16_pair(d.windows + d.p_woff[p] + r0.ref_begin1, g.refLen, [&](int c, uint32_t v) { selW[c] = (uint8_t)(v & 7u); },
                             d.reads + d.p_roff[p] + r0.read_begin1, g.readLen, [&](int i, uint32_t v) { rowW[i] = (uint8_t)(v & 7u); });
        const int ncols = min(g.refLen + 2 * SWB_BANDW_MAX + 2, selCols);
        for (int c = g.refLen; c < ncols; ++c) selW[c] = 0;
    }
    swb_result& r = d.res[p];
    if constexpr (W == 1) {
        // Gapless shortcut.  refLen == readLen and the main diagonal of the sub-matrix scores exactly score1, where score1 is the true
        // Smith-Waterman maximum (fast-path pairs: their result is plain Gotoh).  Every band cell is <= the true local score <= score1,
        // and H(i,i) >= the diagonal prefix sum; were some H(i,i) larger, following the diagonal from there would end above score1.  So
        // H(i,i) equals the prefix sum, the diagonal candidate m = H(i-1,i-1) + s attains it, the tie rule `gmax <= m` (ssw.c:663) picks
        // the diagonal in every cell of the walk, the maximum reaches score1 at the first width, and the traceback (ssw.c:672-751) emits
        // the single op readLen M.  The majority of a real pileup (reads without an indel against their window) ends here.
        if (!resume && g.refLen == g.readLen && (d.p_state[p] & PST_FAST)) {
            int s0 = 0, s1 = 0;
            const int n = g.readLen;
            int i = 0;
            for (; i + 1 < n; i += 2) {
                const unsigned long long t0 = s_rowTab[rowW[i]], t1 = s_rowTab[rowW[i + 1]];
                s0 += (int)prmt((uint32_t)t0, (uint32_t)(t0 >> 32), bandreg_sel(selW[i], d.one));
                s1 += (int)prmt((uint32_t)t1, (uint32_t)(t1 >> 32), bandreg_sel(selW[i + 1], d.one));
            }
            if (i < n) { const unsigned long long t0 = s_rowTab[rowW[i]]; s0 += (int)prmt((uint32_t)t0, (uint32_t)(t0 >> 32), bandreg_sel(selW[i], d.one)); }
            if (s0 + s1 == (int)r.score1) {
                const unsigned long long coff = warp_bump(&d.bump[1], 1ull);
                r.cigar_len = 1; r.cigar_off = (int64_t)coff;
                if ((long long)coff + 1 > d.cigar_cap) { atomicAdd(d.counters + CNT_CIGAR_OVERFLOW, 1); return; }
                d.cigar[coff] = ((uint32_t)n << 4) | 0u;
                d.p_state[p] |= PST_BAND_DONE;
                return;
            }
        }
    }
    // a job this kernel re-runs at a doubled width carries its running maximum along (ssw.c:661 is not reset)
    bandreg_solve<W>(d, p, r, g, selW, rowW, s_rowTab, region, strideW, d.gap_open[p], d.gap_ext[p], r.score1, resume ? d.t_best[p] : 0, nextBase, nextBaseW);
}
