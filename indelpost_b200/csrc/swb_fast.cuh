// swb_fast.cuh — DPX (s16x2) inter-sequence Smith-Waterman sweep for the pairs whose striped result is
// provably plain Gotoh on the zero-padded read:
//     gap_open > gap_extension, window without N, |mat| <= 7, maxScore*readLen <= 1023, readLen <= G*R.
//
// Work decomposition
//   * the two 16-bit lanes of every register hold two DIFFERENT alignments (A in the low half, B in
//     the high half): every DPX instruction updates two cells;
//   * a group of G threads owns one such lane-pair; thread g keeps rows g*R .. g*R+R-1 of BOTH reads in
//     registers (H, E, the per-row score table, the per-row key offset) and the group sweeps the window
//     as a wavefront: at step t thread g is on column t-g and receives (H, F, column-best) of the row
//     above from thread g-1 by __shfl_up.  No shared memory in the inner loop except one 16-bit selector
//     per column.
//   * substitution scores come from ONE PRMT per cell pair: each row holds a 4-byte table
//     {s(A),s(C),s(G),s(T)} per read (two registers), the per-column selector picks the two bytes named by
//     the two windows' bases and sign-extends them into the two 16-bit lanes.
//
// Number representation: every DP value is stored as 16*v + 0x4000 in its lane.
//   * the bias keeps lanes positive, so `x - gap` is a plain 32-bit IADD (FMA pipe) with no borrow
//     between lanes, leaving the ALU pipe to the DPX instructions (measured: 64 lanes/clk/SM each,
//     profiles/dpx_microbench_r01.md);
//   * the factor 16 leaves 4 tag bits: key = value - rowInThread orders cells by (score desc, row asc),
//     which is exactly ssw.c's "smallest read index holding the maximum" rule (ssw.c:341-349,543-551).
//   * rows are END-aligned: the padded read occupies the last Lp of the G*R row slots of its lane; the
//     slots before it get a table of -128 for every base, which pins their H to 0 (= the H[-1][.] = 0
//     boundary of row 0), so there are no dead rows and the key offset of row k is the immediate k.
// Column maxima (needed for the sub-optimal score, ssw.c:366-379 / 568-581) are carried down the
// wavefront as (value, row) and stored per column in shared memory by the last thread; best score, first
// best column (ssw.c:325-333 / 530-534), mask rule and the reverse pass' "first column whose maximum
// equals score1" (ssw.c:337/539) are resolved by a short scan after the sweep.
//
// SW = 1, the "sandwich" sweep: for 8-bit-final pairs whose scores can pass 128+go+ge the result IS the 8-bit pass
// with its signed lazy-F exit test (ssw.c:309-311), which differs from Gotoh only when that test mis-reads a live vertical-gap
// chain -- possible only while the chain's value lies in W = [128, 127+go-ge].  Next to Gotoh (U) the sweep carries a lower
// bound L: the same recurrence, but a vertical-gap CONTINUATION is dropped whenever the interval [L's value, U's value] of the
// chain meets W, and E opens from H without F (ssw.c computes E before the lazy correction).  L <= H(8-bit) <= U cell by cell; if
// L equals U at the cells the outputs are read from -- the best cell (smallest row holding the maximum) of the best column, of the
// sub-optimal column and, in the reverse pass, of the hit column -- every output of the 8-bit pass is Gotoh's (the 8-bit column
// maxima lie between L's and U's, and the columns scanned before those are strictly lower in U already).  One bit per column and lane ("L differs at the column best") rides in bit 15 of the row word down the wavefront.
// Pairs with a flagged column go to the exact kernels.  The same sweep settles overflow verifications (PST_HAVE_WORD): if L's
// maximum reaches 255-bias the 8-bit pass certainly overflowed.
#pragma once
#include <type_traits>
#include "swb_common.cuh"

#define FAST_C      0x4000            // lane bias
#define FAST_CPACK  0x40004000u
#define FAST_SCALE  16
#define FAST_G      16                // threads per lane-pair
#define FAST_MAX_SCORE 1023           // (32767 - 0x4000) / 16

__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { uint32_t d; asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
// raw PRMT (default mode): selector nibble bit 3 = replicate the sign of the selected byte.  The __byte_perm intrinsic
// masks the selector with 0x7777 and cannot express that.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }
__device__ __forceinline__ uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

// per-lane (= per alignment) parameters resolved at kernel start
struct FastLane {
    const int8_t* read; const int8_t* ref;
    int L, Lp, off, ncols, go, ge, mask, target;   // Lp: rows that count (read padded to 8 or 16, ssw.c:169/391); off = G*R - Lp
    int p;                                    // pair index (-1: lane unused)
};

// One group's lane pair: staging, sweep, post-sweep scans.  job list entry j -> pairs (jobs[2j], jobs[2j+1]); an odd tail is paired with itself
template <int R, int DIR, bool GCOLS, int SW, int G>
__device__ __forceinline__ void fast_group(const SwbDev& d, const int32_t* __restrict__ jobs, const int npairs, const int grp, const bool valid,
                                           const int colAlloc, const int verifyX, unsigned char* smem_raw, const uint32_t* s_rowtab)
{
    static_assert(G == 16 || (G == 8 && SW == 0 && DIR == 0), "8-thread groups: plain forward sweep only");
    static_assert(R <= (G == 8 ? 32 : 16), "row tags must fit the scale");
    // representation (see the header): 16 threads x R <= 16 rows with scale 16 / bias 0x4000, or 8 threads x R <= 32 rows with scale 32 /
    // bias 0x2000 (5 tag bits; scores up to 767, matrix entries in [-4, 3] so that 32 * s fits the signed byte PRMT sign-extends)
    constexpr int SC = G == 8 ? 32 : FAST_SCALE;
    constexpr int CB = G == 8 ? 0x2000 : FAST_C;
    constexpr uint32_t CP = (uint32_t)CB * 0x00010001u;
    constexpr uint32_t TAGS = (uint32_t)(SC - 1) * 0x00010001u;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int g = lane % G;
    const int groupInBlock = threadIdx.x / G;

    FastLane ln[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        FastLane& q = ln[s];
        q.p = -1; q.L = 0; q.Lp = 0; q.off = G * R; q.ncols = 0; q.go = 1; q.ge = 0; q.mask = 15; q.target = -1; q.read = nullptr; q.ref = nullptr;
        if (valid) {
            int idx = 2 * grp + s;
            if (idx >= npairs) idx = 2 * grp;                   // odd tail: duplicate lane A (its result is written once)
            const int p = jobs[idx];
            q.p = (s == 1 && 2 * grp + 1 >= npairs) ? -1 : p;
            q.read = d.reads + d.p_roff[p];
            q.ref = d.windows + d.p_woff[p];
            q.go = d.gap_open[p]; q.ge = d.gap_ext[p]; q.mask = d.p_mask[p];
            const int pad = d.p_mode[p] ? 8 : 16;               // p_mode is the semantic chosen by k_classify: 1 word, 0 byte
            if (DIR == 0) { q.L = d.p_rlen[p]; q.ncols = d.p_wlen[p]; }
            else {
                const swb_result& r = d.res[p];
                q.L = r.read_end1 + 1; q.ncols = r.ref_end1 + 1; q.target = r.score1;
            }
            q.Lp = (q.L + pad - 1) / pad * pad;
            q.off = G * R - q.Lp;
        }
    }

    // ---- shared memory: per group, per column: selector (u16), column best value (u32), its row (u32) ----
    // short windows: everything in shared memory (10 bytes per column).  Long windows: only the selectors stay in
    // shared memory (2 bytes per column) and the column bests go to a global scratch (written once per step by the
    // last thread, re-read by the post-sweep scans from L2), which keeps the occupancy register-limited.
    // (a template parameter, not a run-time pointer choice, so the shared-memory case keeps LDS/STS in the hot loop)
    // 8-thread groups keep the two row bytes of a column in 16 bits (8 bytes per column: twice as many groups per warp share the SM's
    // shared memory); 16-thread groups keep the 32-bit row word, whose lane bit 15 carries the sandwich flag
    using RowT = std::conditional_t<G == 8, uint16_t, uint32_t>;
    uint32_t* colv; RowT* colr; uint16_t* selS;
    if constexpr (GCOLS) {
        selS = reinterpret_cast<uint16_t*>(smem_raw + (size_t)groupInBlock * ((size_t)colAlloc * 2));
        colv = d.fast_cols + (size_t)grp * 2 * colAlloc;
        colr = reinterpret_cast<RowT*>(colv + colAlloc);
    } else {
        unsigned char* gbase = smem_raw + (size_t)groupInBlock * ((size_t)colAlloc * (6 + sizeof(RowT)));
        colv = reinterpret_cast<uint32_t*>(gbase);
        colr = reinterpret_cast<RowT*>(colv + colAlloc);
        selS = reinterpret_cast<uint16_t*>(colr + colAlloc);
    }
    // row of lane s (0 / 1) in a column's row word, and its sandwich flag
    auto rowOf = [&](int c, int s) -> int { return G == 8 ? (int)((colr[c] >> (8 * s)) & 0xffu) : (int)((colr[c] >> (16 * s)) & 0x7fffu); };

    const int maxcols = max(ln[0].ncols, ln[1].ncols);
    // PRMT selector per column: byte0 = tabA[bA], byte1 = sign(byte0), byte2 = tabB[bB], byte3 = sign(byte2).  The windows are read as
    // aligned 16-byte chunks spread over the group's threads (up to 15 bytes before / after a window are touched: every sequence blob is
    // a device allocation with 16 bytes of slack), lane A's bases first, then lane B's are OR-ed in.
    for (int c = g; c < maxcols; c += G) selS[c] = (uint16_t)0xC480u;
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int len = ln[s].ncols;
        if (len > 0) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(ln[s].ref);
            const int mis = (int)(a & 15);
            const uint4* base = reinterpret_cast<const uint4*>(a - mis);
            const int nch = (mis + len + 15) >> 4;
            const uint32_t mul = s ? 0x1100u : 0x11u;
            for (int ch = g; ch < nch; ch += G) {
                const uint4 v = __ldg(base + ch);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                const int i0 = ch * 16 - mis;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int idx = i0 + q;
                    if ((unsigned)idx < (unsigned)len) {
                        const int c = DIR ? len - 1 - idx : idx;
                        selS[c] |= (uint16_t)(((w[q >> 2] >> (8 * (q & 3))) & 3u) * mul);
                    }
                }
            }
        }
        __syncwarp();
    }

    // ---- per-row registers (slot g*R+k of the lane holds read row slot - off) -------------------------
    uint32_t H[R], E[R], tA[R], tB[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const int rA = g * R + k - ln[0].off, rB = g * R + k - ln[1].off;
        uint32_t a = 0, b = 0;                                   // pad rows score 0 against everything (ssw.c:402)
        if (rA < 0) a = 0x80808080u; else if (rA < ln[0].L) a = s_rowtab[DIR ? ln[0].read[ln[0].L - 1 - rA] : ln[0].read[rA]];
        if (rB < 0) b = 0x80808080u; else if (rB < ln[1].L) b = s_rowtab[DIR ? ln[1].read[ln[1].L - 1 - rB] : ln[1].read[rB]];
        tA[k] = a; tB[k] = b;
        H[k] = CP; E[k] = CP;
    }
    uint32_t HL[SW ? R : 1], EL[SW ? R : 1];                    // the lower-bound recurrence (SW only)
    if constexpr (SW) {
#pragma unroll
        for (int k = 0; k < R; ++k) { HL[k] = CP; EL[k] = CP; }
    }
    // SW: drop window W = [128, 127 + go - ge] as sign-bit tests: (hiX - contL) has lane bit 15 set iff contL <= hi,
    // (contU + loY) iff contU >= lo (all lane values are below 0x8000, no carry between the lanes)
    uint32_t geP = pack2(SC * ln[0].ge, SC * ln[1].ge);
    uint32_t hiX = pack2(SC * (127 + ln[0].go - ln[0].ge) + CB + 0x8000, SC * (127 + ln[1].go - ln[1].ge) + CB + 0x8000);
    const uint32_t loY = pack2(0x8000 - (SC * 128 + CB), 0x8000 - (SC * 128 + CB));
    if constexpr (SW) asm volatile("" : "+r"(geP), "+r"(hiX));
    uint32_t goP = pack2(SC * ln[0].go, SC * ln[1].go);
    uint32_t ngeP = pack2(-SC * ln[0].ge, -SC * ln[1].ge);
    uint32_t rowBase = pack2(g * R, g * R);
    const uint32_t one = (uint32_t)d.one;
    // keep the loop invariants in registers (ptxas otherwise rematerialises them every step)
    asm volatile("" : "+r"(goP), "+r"(ngeP), "+r"(rowBase));
#pragma unroll
    for (int k = 0; k < R; ++k) asm volatile("" : "+r"(tA[k]), "+r"(tB[k]));
    __syncwarp();

    // warp-uniform step count
    int nsteps = maxcols + G - 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nsteps = max(nsteps, __shfl_xor_sync(FULL, nsteps, o));
    if (!valid) nsteps = max(nsteps, 0);

    uint32_t outH = CP, outF = CP, outV = 0, outRow = 0, prevInH = CP;
    uint32_t outHL = CP, outFL = CP, prevInHL = CP, runL = 0;
    const uint32_t targetV = pack2(ln[0].target >= 0 ? SC * ln[0].target + CB : 0x7fff, ln[1].target >= 0 ? SC * ln[1].target + CB : 0x7fff);
    bool done = false;                                         // reverse pass: both lanes have hit their target

    // boundary of the first thread of a group as masks (loop invariants kept in registers: no per-step predicate set-up)
    // boundary lane (g == 0): inputs become the constants of row -1.  Done as x * m1 + c0 with an opaque m1 in {0, 1}: a true
    // IMAD (FMA pipe) instead of a LOP3 on the ALU pipe that bounds this loop
    uint32_t m1 = g == 0 ? 0u : 1u, c0 = g == 0 ? CP : 0u;
    asm volatile("" : "+r"(m1), "+r"(c0));
    // steps in which every thread of the warp is on a valid column need no range check: [G-1, smallest maxcols of the warp's groups)
    int steadyEnd = valid ? maxcols : 0;
    steadyEnd = min(steadyEnd, __shfl_xor_sync(FULL, steadyEnd, 16));
    if (G == 8) steadyEnd = min(steadyEnd, __shfl_xor_sync(FULL, steadyEnd, 8));
    if (DIR == 1) steadyEnd = 0;                                // the reverse sweep stops early: keep it checked

    auto step = [&](const int t, auto checkedTag) {
        constexpr bool CHECKED = decltype(checkedTag)::value;
        uint32_t inH = __shfl_up_sync(FULL, outH, 1, G);
        uint32_t inF = __shfl_up_sync(FULL, outF, 1, G);
        uint32_t inV = __shfl_up_sync(FULL, outV, 1, G);
        uint32_t inRow = __shfl_up_sync(FULL, outRow, 1, G);
        inH = inH * m1 + c0; inF = inF * m1 + c0; inV *= m1; inRow *= m1;
        uint32_t inHL = 0, inFL = 0;
        if constexpr (SW) {
            inHL = __shfl_up_sync(FULL, outHL, 1, G) * m1 + c0;
            inFL = __shfl_up_sync(FULL, outFL, 1, G) * m1 + c0;
        }
        const int c = t - g;
        if (!CHECKED || (c >= 0 && c < maxcols && !done)) {
            const uint32_t sel = selS[c];
            uint32_t F = inF, hd = prevInH, cm = 0;
            prevInH = inH;
            [[maybe_unused]] uint32_t FL = inFL, hdL = prevInHL, cmL = 0;
            prevInHL = inHL;
#pragma unroll
            for (int k = 0; k < R; k += 2) {
                uint32_t key[2];
                [[maybe_unused]] uint32_t keyL[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int kk = k + u;
                    if (kk >= R) { key[u] = 0; if constexpr (SW) keyL[u] = 0; break; }      // odd R: the last pair has one row (keys are >= CB > 0)
                    const uint32_t s = prmt(tA[kk], tB[kk], sel);
                    uint32_t h = __viaddmax_s16x2(hd, s, E[kk]);                 // max(Hdiag + s, E)
                    h = __vimax3_s16x2(h, F, CP);                        // max(., F, 0)
                    hd = H[kk]; H[kk] = h;
                    key[u] = h * one - (uint32_t)(kk * 0x00010001);              // value - rowInThread: a true IMAD (FMA pipe), no lane borrow
                    const uint32_t hg = h - goP;                                 // FMA-pipe IADD; no lane borrow: h >= 0x4000 > 16*go
                    E[kk] = __viaddmax_s16x2(E[kk], ngeP, hg);                   // max(E - ge, H - go)
                    if constexpr (!SW) {
                        F = __viaddmax_s16x2(F, ngeP, hg);                       // max(F - ge, H - go)
                    } else {
                        // the lower bound L: H without F first (E opens from it), then the vertical gap
                        uint32_t hn = __viaddmax_s16x2(hdL, s, EL[kk]);
                        hn = vmax2(hn, CP);
                        const uint32_t hL = vmax2(hn, FL);
                        hdL = HL[kk]; HL[kk] = hL;
                        keyL[u] = hL * one - (uint32_t)(kk * 0x00010001);
                        EL[kk] = __viaddmax_s16x2(EL[kk], ngeP, hn - goP);
                        // continuations of the two chains (every F lane is >= 0x4000 - 16*go > 16*ge: no borrow); L's is dropped
                        // when [contL, contU] meets W
                        const uint32_t contU = F - geP;
                        uint32_t contL = FL - geP;
                        const uint32_t z = (hiX - contL) & (contU + loY);
                        const uint32_t drop = prmt(z, 0u, 0xBB99u);              // 0xFFFF in the lanes where both sign bits are set
                        contL &= ~drop;
                        F = vmax2(contU, hg);
                        FL = vmax2(contL, hL - goP);
                    }
                }
                cm = k == 0 ? vmax2(key[0], key[1]) : __vimax3_s16x2(cm, key[0], key[1]);   // column best over keys, two rows per ALU instruction
                if constexpr (SW) cmL = k == 0 ? vmax2(keyL[0], keyL[1]) : __vimax3_s16x2(cmL, keyL[0], keyL[1]);
            }
            outH = H[R - 1]; outF = F;
            if constexpr (SW) { outHL = HL[R - 1]; outFL = FL; runL = vmax2(runL, cmL); }
            // local column best -> (value, absolute row); merge with the rows above (they win ties)
            const uint32_t lv = (cm + TAGS) & ~TAGS;
            uint32_t lrow = lv - cm + rowBase;
            // SW: bit 15 of the lane = "L differs from U at this thread's column-best cell" (equal keys <=> same value in the same row);
            // it travels with the row word, so the column's final word carries the flag of the cell that won
            if constexpr (SW) lrow += ((cm ^ cmL) + 0x7FFF7FFFu) & 0x80008000u;
            const uint32_t x = inV + 0x80008000u - lv;                       // lane bit15 set <=> inV >= lv (inV lanes are < 0x8000)
            const uint32_t keep = prmt(x, 0u, 0xBB99u);                 // 0xFFFF in lanes where the upstream value stays
            outV = vmax2(inV, lv);
            outRow = (inRow & keep) | (lrow & ~keep);
            if (g == G - 1) { colv[c] = outV; colr[c] = G == 8 ? (RowT)prmt(outRow, 0u, 0x4420u) : (RowT)outRow; }
        }
    };

    int t = 0;
    for (; t < min(G - 1, nsteps); ++t) step(t, std::true_type{});
#pragma unroll (R == 10 ? 8 : 4)
    for (; t < steadyEnd; ++t) step(t, std::false_type{});
    for (; t < nsteps; ++t) {
        step(t, std::true_type{});
        if (DIR == 1 && (t & 31) == 31) {
            // every 32 steps: have both lanes of this group seen a column whose maximum equals score1?
            __syncwarp();
            bool hitA = ln[0].p < 0 && ln[1].p < 0, hitB = true;
            const int hi = min(t - (G - 1), maxcols - 1);
            bool a = false, b = false;
            for (int c2 = g; c2 <= hi; c2 += G) {
                const uint32_t v = colv[c2] ^ targetV;
                if (c2 < ln[0].ncols && (v & 0xffffu) == 0) a = true;
                if (c2 < ln[1].ncols && (v >> 16) == 0) b = true;
            }
            const unsigned ba = __ballot_sync(FULL, a), bb = __ballot_sync(FULL, b);
            const unsigned gm = ((1u << G) - 1u) << ((lane / G) * G);
            hitA = hitA || (ba & gm) != 0 || ln[0].p < 0 || hi >= ln[0].ncols - 1;
            hitB = (bb & gm) != 0 || ln[1].p < 0 || hi >= ln[1].ncols - 1;
            done = !valid || (hitA && hitB);
            if (__all_sync(FULL, done)) break;
        }
    }
    __syncwarp();
    if (!valid) return;

    // ---- post-sweep scans (per lane) -------------------------------------------------------------------
    const unsigned GM = ((1u << G) - 1u) << ((lane / G) * G);      // shuffles below are group-scoped: the two groups of a warp may diverge
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const FastLane& q = ln[s];
        if (q.p < 0) continue;
        const int sh = 16 * s;
        const int p = q.p;
        swb_result& r = d.res[p];
        const bool wordSem = d.p_mode[p] != 0;
        // read-length bucket of the 16-thread-group lists this pair continues on (this instantiation's own for G = 16)
        const int BKT = G == 8 ? max(0, (((d.p_rlen[p] + 15) & ~15) + 31) / 32 - 1) : R / 2 - 1;
        if (DIR == 0 && SW && (d.p_state[p] & PST_HAVE_WORD)) {
            // overflow verification of a provisional 16-bit result: L <= H(8-bit), so L reaching 255-bias proves the overflow
            int mv = (int)((runL >> sh) & 0xffffu);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) mv = max(mv, __shfl_xor_sync(GM, mv, o, G));
            if (g == 0) {
                const int ml = mv > CB ? (((mv + (SC - 1)) & ~(SC - 1)) - CB) / SC : 0;
                if (ml >= 255 - d.bias) atomicAdd(d.counters + CNT_SW_VERIFIED, 1);
                else if (verifyX >= 0) list_push(d.list[verifyX], d.counters + verifyX, p);
            }
            continue;
        }
        if (DIR == 0) {
            // best score and first column reaching it
            int bv = 0, bc = 0x7fffffff;
            for (int c = g; c < q.ncols; c += G) {
                const int v = (int)((colv[c] >> sh) & 0xffffu);
                if (v > bv) { bv = v; bc = c; }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(GM, bv, o, G), oc = __shfl_xor_sync(GM, bc, o, G);
                if (ov > bv || (ov == bv && oc < bc)) { bv = ov; bc = oc; }
            }
            const int T = bv > CB ? (bv - CB) / SC : 0;
            int end_ref, end_read;
            if (T > 0) { end_ref = bc; end_read = min(rowOf(bc, s) - q.off, q.L - 1); }
            else { end_ref = wordSem ? 0 : -1; end_read = 0; }                      // ssw.c:427 / 220
            // sub-optimal score outside the mask (ssw.c:366-379 byte, 568-581 word)
            const int edgeL = max(end_ref - q.mask, 0);
            const int edgeR = min(end_ref + q.mask, q.ncols) + (wordSem ? 0 : 1);
            int sv = CB, si = 0x7fffffff;
            for (int c = g; c < q.ncols; c += G) {
                if (c < edgeL || c >= edgeR) {
                    const int v = (int)((colv[c] >> sh) & 0xffffu);
                    if (v > sv) { sv = v; si = c; }
                }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(GM, sv, o, G), oi = __shfl_xor_sync(GM, si, o, G);
                if (ov > sv || (ov == sv && oi < si)) { sv = ov; si = oi; }
            }
            // first column whose maximum reaches 128+go+ge: every F value before it is < 128+ge, so the signed lazy-F test
            // (ssw.c:311) is correct there and the 8-bit pass is exact Gotoh up to that column (used by the certificate)
            const int thr = CB + SC * (128 + q.go + q.ge);
            int cs = q.ncols;
            for (int c = g; c < q.ncols; c += G) {
                if ((int)((colv[c] >> sh) & 0xffffu) >= thr) { cs = c; break; }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) cs = min(cs, __shfl_xor_sync(GM, cs, o, G));
            const int s2 = sv > CB ? (sv - CB) / SC : 0;
            const int r2 = s2 > 0 ? si : 0;
            // SW: the 8-bit pass's column maxima lie between L's and U's, so its outputs are U's as soon as L equals U at the two
            // cells they are read from: the best cell of the best column (score1, ref_end1, read_end1; every earlier column and every
            // row above stay below it in U already) and the best cell of the sub-optimal column (score2, ref_end2)
            uint32_t fl = 0;
            if constexpr (SW) {
                if (T > 0) fl |= colr[bc];
                if (s2 > 0 && q.mask >= 15) fl |= colr[si];
                fl = (fl >> sh) & 0x8000u;
            }
            if (g == 0) {
                warp_count(d.counters + CNT_CELLS_FWD, (unsigned long long)q.Lp * q.ncols);
                const int limit = 255 - d.bias;                                    // 8-bit pass overflows at max + bias >= 255 (ssw.c:327)
                bool accept;
                if (wordSem) accept = d.score_size == 1 || T >= limit;            // else the result is a byte-mode one: exact path decides
                else accept = T < limit && T < 128 + q.go + q.ge;                  // safe zone of the signed lazy-F test
                const bool unsafeZone = !wordSem && T < limit && T >= 128 + q.go + q.ge;
                if (SW && unsafeZone) {                                            // ... or certified by the sandwich
                    accept = fl == 0;
                    atomicAdd(d.counters + (accept ? CNT_SW_CERTIFIED : CNT_SW_REJECTED), 1);
                }
                if (!accept) {
                    d.p_mode[p] = 0; d.p_state[p] = 0;
                    // a 16-bit-semantics sweep whose score stays below the 8-bit limit: the result is the 8-bit pass's.  It is swept again
                    // in 8-bit semantics (16-row padding, the other mask edge) by the sandwich flavour, which runs after this launch
                    if (!SW && wordSem && d.score_size == 2 && !(d.opt & 128)) list_push(d.list[LIST_SW_FWD + BKT], d.counters + CNT_SW_FWD + BKT, p);
                    else list_push(d.list[LIST_BYTE_FWD], d.counters + CNT_BYTE_FWD, p);
                } else {
                    r.score1 = (uint16_t)T; r.ref_end1 = end_ref; r.read_end1 = end_read;
                    d.p_csafe[p] = cs;
                    if (q.mask >= 15) { r.score2 = (uint16_t)s2; r.ref_end2 = r2; } else { r.score2 = 0; r.ref_end2 = -1; }
                    d.p_state[p] = (wordSem && d.score_size == 2) ? (PST_FAST | PST_NEED_CERT) : PST_FAST;
                    { const unsigned am = __activemask(); if ((int)(threadIdx.x & 31) == __ffs(am) - 1) atomicAdd(d.counters + CNT_FAST_DONE, __popc(am)); }
                    const bool scoreOnly = d.flag == 0 || (d.flag == 2 && T < (int)d.filters);     // ssw.c:872
                    if (!scoreOnly) {
                        // reverse pass: banded (swb_revband.cuh) when the score deficit B = mx*rows - T bounds the deviation
                        // of every path scoring T from the main diagonal to one of the band classes, else the wavefront sweep
                        int cls = -1;
                        if (T > 0 && q.ge > 0 && !(d.opt & 4) && !(SW && unsafeZone)) {
                            const int B = d.max_score * (end_read + 1) - T;
                            int wd = 0, wi = 0;
                            if (B >= q.go) { wd = (B - q.go) / q.ge + 1; wi = (B - q.go + q.ge) / (d.max_score + q.ge); }
                            cls = revb_class(wi, wd);
                        }
                        if (cls >= 0) list_push(d.list[LIST_REVB + cls], d.counters + LIST_REVB + cls, p);
                        else if (SW && unsafeZone) list_push(d.list[LIST_SW_REV + BKT], d.counters + CNT_SW_REV + BKT, p);   // the reverse pass needs its own certificate
                        else list_push(d.list[LIST_FAST_REV + BKT], d.counters + CNT_FAST_REV + BKT, p);
                    }
                    else d.p_state[p] |= PST_BAND_DONE;                            // no reverse pass / traceback will follow
                }
            }
        } else {
            // first column (scanning away from ref_end1) whose maximum equals score1 (ssw.c:337 / 539)
            const int tv = (int)((targetV >> sh) & 0xffffu);
            int hc = 0x7fffffff;
            for (int c = g; c < q.ncols; c += G) {
                const int v = (int)((colv[c] >> sh) & 0xffffu);
                if (v == tv) { hc = c; break; }
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) hc = min(hc, __shfl_xor_sync(GM, hc, o, G));
            // SW: the 8-bit reverse pass stops at the first column whose maximum equals score1; every earlier column stays below it in U
            // already, so L has to equal U only at the best cell of the hit column
            uint32_t fl = 0;
            if constexpr (SW) { if (hc != 0x7fffffff) fl = (colr[hc] >> sh) & 0x8000u; }
            if (g == 0) {
                if (hc == 0x7fffffff || q.target <= 0 || fl) {
                    // no column reaches score1 (or score 0 corner): let the exact path reproduce ssw.c literally
                    const int md = d.p_mode[p];
                    d.p_state[p] &= ~PST_FAST;
                    list_push(d.list[md ? LIST_WORD_REV2 : LIST_BYTE_REV2], d.counters + (md ? LIST_WORD_REV2 : LIST_BYTE_REV2), p);
                } else {
                    warp_count(d.counters + CNT_CELLS_REV, (unsigned long long)q.Lp * (hc + 1));
                    r.ref_begin1 = r.ref_end1 - hc;                                 // ssw.c:885-886
                    r.read_begin1 = r.read_end1 - (rowOf(hc, s) - q.off);
                    const int f = d.flag;
                    const bool noCigar = (7 & f) == 0 || ((2 & f) != 0 && (int)r.score1 < (int)d.filters) ||
                                         ((4 & f) != 0 && (r.ref_end1 - r.ref_begin1 > d.filterd || r.read_end1 - r.read_begin1 > d.filterd));
                    if (!noCigar) push_band(d, p, r); else d.p_state[p] |= PST_BAND_DONE;
                }
            }
        }
    }
}

// Grid: one block per FAST-group bundle of the list slice; a launch whose grid is smaller than the slice (the sandwich flavour is
// launched against an upper bound of its list) strides over it.
template <int R, int DIR, bool GCOLS, int SW = 0, int G = FAST_G>
__global__ void __launch_bounds__(128, (G == 8 && R > 16) ? 4 : 0)      // 19 rows per thread: cap at 128 registers (4 blocks per SM)
k_fast(SwbDev d, const int32_t* __restrict__ jobs, const int32_t* __restrict__ njobs_ptr, int colAlloc, int pairOffset, int pairLimit, int verifyX = -1)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_rowtab[SWB_MAX_N];
    // this launch serves the slice [pairOffset, pairOffset + pairLimit) of the job list (pairOffset is even)
    jobs += pairOffset;
    const int npairs = min(max(*njobs_ptr - pairOffset, 0), pairLimit);
    const int ngroups = (npairs + 1) >> 1;
    const int groupsPerBlock = blockDim.x / G;
    if (blockIdx.x * groupsPerBlock >= ngroups) return;
    // per-read-base score table: byte nt = 16 * mat[nt][rb]  (qP_word's profile cell, ssw.c:402, scaled)
    if (threadIdx.x < d.n) {
        uint32_t t = 0;
        for (int nt = 0; nt < 4; ++nt) t |= (uint32_t)(uint8_t)(int8_t)((G == 8 ? 32 : FAST_SCALE) * d.mat[nt * d.n + threadIdx.x]) << (8 * nt);
        s_rowtab[threadIdx.x] = t;
    }
    __syncthreads();
    if constexpr (SW) {
        for (int g0 = blockIdx.x * groupsPerBlock; g0 < ngroups; g0 += gridDim.x * groupsPerBlock) {
            const int grp = g0 + threadIdx.x / G;
            fast_group<R, DIR, GCOLS, SW, G>(d, jobs, npairs, grp, grp < ngroups, colAlloc, verifyX, smem_raw, s_rowtab);
            __syncwarp();                                       // the group's shared-memory columns are reused by its next lane pair
        }
    } else {
        // the plain sweep is launched with one block per bundle of its list (no loop: the loop state costs the hot kernel registers)
        const int grp = blockIdx.x * groupsPerBlock + threadIdx.x / G;
        fast_group<R, DIR, GCOLS, SW, G>(d, jobs, npairs, grp, grp < ngroups, colAlloc, verifyX, smem_raw, s_rowtab);
    }
}
