// swb_band.cuh — banded affine-gap DP + traceback -> BAM CIGAR (replaces banded_sw, ssw.c:588-772).
//
// One thread per alignment.  The DP is followed literally: the band coordinate macros set_u/set_d
// (ssw.c:92-95), the three rolling row buffers and which of their slots are cleared at each row start
// (ssw.c:627, 633), the tie rules of ssw.c:647-664, the running maximum that is NOT reset between band
// widenings (ssw.c:661, 669), the `while (i >= 0 && j > 0)` exit and the tail rules (ssw.c:679, 734-751).
// When the band is wider than the matrix those details decide what a cell reads as its upper neighbour,
// so a "clean" formulation would not be bit-exact.
//
// Direction information (3 bytes per cell in the reference) is packed into 4 bits per cell, 8 cells per
// 32-bit word, and kept in a global scratch arena (one word store per 8 cells instead of a 32-byte sector
// per cell); each thread bump-allocates ceil((2w+1)/8)*4*readLen bytes.  The three rolling rows live in
// shared memory laid out [slot][thread], which is bank-conflict free whatever slot each thread touches.
// Band doubling (ssw.c:668-669) happens inside the kernel with a fresh scratch allocation; a pair is only
// re-queued for a later launch when the scratch arena is exhausted or its band outgrows the shared-memory
// rows of its instantiation.  Jobs are bucketed by band-width class and launched widest-first in one grid, so
// the threads of a warp run bands of similar width and the short ones fill the tail.  The traceback is a single
// walk with the band rows held / prefetched in registers; its ops are buffered in registers and written,
// already reversed, into the output arena.
#pragma once
#include "swb_common.cuh"
#include "swb_cert.cuh"

// Instantiations: BW = largest half-width whose rolling rows fit the shared-memory layout (0 = rows in global
// memory, any width), T = threads per block, RING = circular buffer of window bases per thread (>= band width).
//   k_band<16, 128> : classes 0-3 (the bulk)   k_band<48, 64> : class 4   k_band<112, 32> : class 5   k_band<254, 32> : class 6
//   k_band<0, 128>  : anything needing more than 512 slots, or scores beyond 16 bits (global rows; latency bound, rare)
#define SWB_BAND_LOCAL_BW 16
#define SWB_BAND_MID_BW 48
#define SWB_BAND_WIDE_BW 112
#define SWB_BAND_HUGE_BW 254                           // 512 slots: every band over a window of up to 509 columns
#define SWB_BAND_THREADS 128
#define SWB_BAND_MID_THREADS 64
#define SWB_BAND_WIDE_THREADS 32
#define SWB_BAND_HUGE_THREADS 32
#define SWB_BAND_MAX16 30000                          // largest score the 16-bit shared-memory rows may hold
__host__ __device__ constexpr int band_rows_w(int BW) { return 2 * BW + 4; }
__host__ __device__ constexpr int band_ring(int BW) { return BW <= 16 ? 64 : (BW <= 48 ? 128 : (BW <= 112 ? 256 : 512)); }
__host__ __device__ constexpr int band_smem_bytes(int BW, int T) { return BW == 0 ? 0 : 3 * band_rows_w(BW) * T * 2 + band_ring(BW) * T; }

__device__ __forceinline__ int band_x(int w, int i) { int x = i - w; return x > 0 ? x : 0; }

struct BandGeom {
    int w, width_d, strideW, refLen, readLen;           // strideW: 32-bit words per band row
    __device__ __forceinline__ int beg(int i) const { int b = i - w; return b > 0 ? b : 0; }
    __device__ __forceinline__ int end(int i) const { int e = i + w; return e < refLen - 1 ? e : refLen - 1; }
};

// value the reference would read at direction_line[set_d(i, j, state)] (ssw.c:680-681); 0 = not a
// computed cell (the reference reads uninitialised memory there; we report a traceback error)
__device__ __forceinline__ int band_dir_at(const uint32_t* dir, const BandGeom& g, int i, int j, int state) {
    int x = j - band_x(g.w, i);
    int ii = i, st = state;
    if (x < 0 || x >= g.width_d) {
        // literal flat index into the 3-byte-per-cell matrix lands in another row
        long long flat = (long long)g.width_d * 3 * i + (long long)x * 3 + state;
        if (flat < 0 || flat >= (long long)g.width_d * 3 * g.readLen) return 0;
        long long cell = flat / 3; st = (int)(flat - cell * 3);
        ii = (int)(cell / g.width_d); x = (int)(cell - (long long)ii * g.width_d);
    }
    const int jj = x + band_x(g.w, ii);
    if (jj < g.beg(ii) || jj > g.end(ii)) return 0;
    const int b = (int)((dir[(size_t)ii * g.strideW + (x >> 3)] >> (4 * (x & 7))) & 15u);
    const int de = 2 + (b & 1), df = 4 + ((b >> 1) & 1);
    if (st == 0) return de;
    if (st == 1) return df;
    const int sel = (b >> 2) & 3;
    return sel == 0 ? 1 : (sel == 1 ? de : df);
}

// Walk the traceback (ssw.c:672-751).  The packed direction words of the row being visited are held in registers and
// the row above is prefetched while the current row is processed (a step moves up by at most one row), so the walk is
// not a chain of dependent L2 round trips.  Ops are produced last-to-first; the first `BAND_OPBUF` of them are kept in
// registers so the common case needs a single walk: the caller allocates arena space for the returned count and
// calls band_emit().  Returns the op count, or -1 for the reference's "Trace back error" (ssw.c:711-719).
#define BAND_OPBUF 12
#define BAND_ROWWORDS 5                               // words per band row held in registers (band width <= 40)

struct BandOps { uint32_t op[BAND_OPBUF]; int n; };

__device__ __forceinline__ void band_push(BandOps& o, uint32_t v, uint32_t* out, int total) {
    if (o.n < BAND_OPBUF) {
#pragma unroll
        for (int q = 0; q < BAND_OPBUF; ++q) if (q == o.n) o.op[q] = v;      // register-indexed store
    }
    ++o.n;
    if (out) out[total - o.n] = v;
}

template <bool REGROWS>
__device__ __forceinline__ int band_traceback(const uint32_t* dir, const BandGeom& g, BandOps& ops, uint32_t* out, int total) {
    int i = g.readLen - 1, j = g.refLen - 1;
    int e = 0, state = 2;
    int op = 0, prev_op = 0;                           // 0 M, 1 I, 2 D  (BAM op codes)
    ops.n = 0;
    uint32_t cur[BAND_ROWWORDS], nxt[BAND_ROWWORDS];
    int curRow = -1;
    if (REGROWS && i >= 0) {
#pragma unroll
        for (int q = 0; q < BAND_ROWWORDS; ++q) cur[q] = q < g.strideW ? dir[(size_t)i * g.strideW + q] : 0u;
#pragma unroll
        for (int q = 0; q < BAND_ROWWORDS; ++q) nxt[q] = (q < g.strideW && i >= 1) ? dir[(size_t)(i - 1) * g.strideW + q] : 0u;
        curRow = i;
    }
    while (i >= 0 && j > 0) {                          // ssw.c:679
        int dv;
        const int x = j - band_x(g.w, i);
        if (REGROWS && x >= 0 && x < g.width_d && j <= g.end(i)) {
            if (curRow != i) {                         // moved up: rotate the prefetched row in (or reload after a fallback step), prefetch the next
                const bool oneUp = curRow - 1 == i;
#pragma unroll
                for (int q = 0; q < BAND_ROWWORDS; ++q) cur[q] = oneUp ? nxt[q] : (q < g.strideW ? dir[(size_t)i * g.strideW + q] : 0u);
#pragma unroll
                for (int q = 0; q < BAND_ROWWORDS; ++q) nxt[q] = (q < g.strideW && i >= 1) ? dir[(size_t)(i - 1) * g.strideW + q] : 0u;
                curRow = i;
            }
            uint32_t w = cur[0];
#pragma unroll
            for (int q = 1; q < BAND_ROWWORDS; ++q) if ((x >> 3) == q) w = cur[q];
            const int b = (int)((w >> (4 * (x & 7))) & 15u);
            const int de = 2 + (b & 1), df = 4 + ((b >> 1) & 1);
            const int sel = (b >> 2) & 3;
            dv = state == 0 ? de : (state == 1 ? df : (sel == 0 ? 1 : (sel == 1 ? de : df)));
        } else {
            dv = band_dir_at(dir, g, i, j, state);     // out-of-band corner cases, literal index arithmetic
        }
        switch (dv) {
            case 1: --i; --j; state = 2; op = 0; break;
            case 2: --i;      state = 0; op = 1; break;
            case 3: --i;      state = 2; op = 1; break;
            case 4: --j;      state = 1; op = 2; break;
            case 5: --j;      state = 2; op = 2; break;
            default: return -1;                        // "Trace back error", ssw.c:711-719
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

// rolling-row accessor: shared memory [slot][thread] (LOCAL) or a private global array (wide bands)
template <bool LOCAL, int T> struct BandRow;
template <int T> struct BandRow<true, T>  { short* p; __device__ __forceinline__ short& operator[](int u) const { return p[u * T]; } };
template <int T> struct BandRow<false, T> { int* p;   __device__ __forceinline__ int& operator[](int u) const { return p[u]; } };

template <int BW, int T>
__global__ void __launch_bounds__(T)
k_band(SwbDev d, int listBase, int firstClass, int lastClass, int nextBase)
{
    constexpr bool LOCAL = BW > 0;
    constexpr int ROWS_W = band_rows_w(BW);
    constexpr int RING = band_ring(BW);
    __shared__ unsigned long long s_rowTab[8];         // per read base: its scores against every window base, 8 x int8 (n <= 8)
    if (d.n <= 8 && threadIdx.x < d.n) {
        unsigned long long tab = 0;
        for (int nt = 0; nt < d.n; ++nt) tab |= (unsigned long long)(uint8_t)d.mat[nt * d.n + threadIdx.x] << (8 * nt);
        s_rowTab[threadIdx.x] = tab;
    }
    __syncthreads();
    // one launch serves the classes firstClass..lastClass, widest first (blocks are scheduled in order, so the long
    // threads start first and the short ones fill the tail); thread t indexes the concatenation of their lists
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int p = -1;
    for (int k = lastClass; k >= firstClass; --k) {
        // each class starts at a block boundary so a block never mixes classes
        const int nk = d.counters[listBase + k];
        const int padded = (nk + T - 1) / T * T;
        if (t < padded) { if (t < nk) p = d.list[listBase + k][t]; break; }
        t -= padded;
    }
    if (p < 0) return;
    swb_result& r = d.res[p];

    BandGeom g;
    g.refLen = r.ref_end1 - r.ref_begin1 + 1;          // ssw.c:897-899
    g.readLen = r.read_end1 - r.read_begin1 + 1;
    int bw, best;
    if (d.t_bw[p] == 0) { int dl = g.refLen - g.readLen; bw = (dl < 0 ? -dl : dl) + 1; best = 0; }
    else { bw = d.t_bw[p]; best = d.t_best[p]; }       // re-queued: resume with the saved width and running maximum
    const int len = g.refLen > g.readLen ? g.refLen : g.readLen;
    const int score = r.score1;
    const int go = d.gap_open[p], ge = d.gap_ext[p];
    const int n = d.n;
    const int8_t* mat = d.mat;
    const bool fits16 = (long long)d.max_score * (g.readLen > 0 ? g.readLen : 1) <= SWB_BAND_MAX16;
    // score1 == 0 in byte mode leaves ref_begin1 == -1 and the reference reads ref[-1] (undefined);
    // the 1x1 problem it then solves gives "1M" whatever that byte is
    const bool ubRef = r.ref_begin1 < 0;
    const int8_t* ref = d.windows + d.p_woff[p] + (ubRef ? 0 : r.ref_begin1);
    const int8_t* read = d.reads + d.p_roff[p] + (r.read_begin1 < 0 ? 0 : r.read_begin1);
    const bool packRow = n <= 8;                       // substitution scores of one read base vs every window base in 8 x int8

    extern __shared__ __align__(16) unsigned char band_smem[];
    BandRow<LOCAL, T> hPrev, ePrev, hCur;
    unsigned char* refRing = nullptr;                  // LOCAL: window bases of the current band, slot = column & (RING-1)
    if constexpr (LOCAL) {
        hPrev.p = reinterpret_cast<short*>(band_smem) + threadIdx.x;
        ePrev.p = hPrev.p + ROWS_W * T;
        hCur.p = ePrev.p + ROWS_W * T;
        refRing = band_smem + 3 * ROWS_W * T * 2 + threadIdx.x;
    }
    uint32_t* dir = nullptr;
    long long cells = 0;

    for (;;) {                                         // band widening loop, ssw.c:612-669
        if (LOCAL && (band_rows_needed(bw, g.refLen) > ROWS_W || !fits16)) {
            // outgrew this instantiation's shared-memory rows: continue in the next wider one
            d.t_bw[p] = bw; d.t_best[p] = best;
            // the class whose instantiation holds these rows (7: global rows), or the warp-per-alignment kernel
            if (ubRef || r.read_begin1 < 0) { const int c = band_class(bw, g.refLen, fits16); list_push(d.list[nextBase + c], d.counters + nextBase + c, p); }
            else requeue_band(d, nextBase, p, bw, g.refLen, g.readLen, r, fits16);
            warp_count(d.counters + CNT_CELLS_BAND, (unsigned long long)cells);
            return;
        }
        g.w = bw; g.width_d = 2 * bw + 1; g.strideW = (g.width_d + 7) >> 3;
        const int width = 2 * bw + 3;
        // ---- scratch: packed direction words (+ the three rolling rows when they do not live in shared memory)
        const long long dirBytes = (long long)g.strideW * 4 * (g.readLen > 0 ? g.readLen : 1);
        const long long rowBytes = LOCAL ? 0 : 3ll * (width + 1) * 4;
        const long long need = dirBytes + rowBytes;
        const unsigned long long off = warp_bump(&d.bump[0], (unsigned long long)need);
        if ((long long)off + need > d.band_cap) {      // out of scratch: retry in a later launch
            d.t_bw[p] = bw; d.t_best[p] = best;
            atomicAdd(d.counters + CNT_BAND_OVERFLOW, 1);
            const int c = BW == 0 ? 7 : band_class(bw, g.refLen, fits16);
            list_push(d.list[nextBase + c], d.counters + nextBase + c, p);
            warp_count(d.counters + CNT_CELLS_BAND, (unsigned long long)cells);
            return;
        }
        dir = reinterpret_cast<uint32_t*>(d.band + off);
        if constexpr (!LOCAL) {
            hPrev.p = reinterpret_cast<int*>(d.band + off + dirBytes);
            ePrev.p = hPrev.p + (width + 1); hCur.p = ePrev.p + (width + 1);
        }
        // the reference's buffers are realloc'ed across widenings and not cleared; every slot it reads is
        // written first within an iteration except where it reads uninitialised memory -- start from zeros
        {
            // (only slots 0 .. refLen+1 are ever touched, see band_class)
            const int zmax = LOCAL ? min(width, ROWS_W - 1) : width;
            for (int j = 0; j <= zmax; ++j) { hPrev[j] = 0; ePrev[j] = 0; hCur[j] = 0; }
        }

        int ringHi = -1;                               // last window column already in refRing
        int rbNext = g.readLen > 0 ? read[0] : 0;      // software prefetch: next row's read base and the next window base
        int refNext = (!ubRef && g.refLen > 0) ? ref[0] : 0;
        for (int i = 0; i < g.readLen; ++i) {          // ssw.c:628-667
            const int beg = g.beg(i), end = g.end(i);
            if constexpr (LOCAL) {
                for (; ringHi < end; ) {
                    ++ringHi;
                    refRing[(ringHi & (RING - 1)) * T] = (unsigned char)refNext;
                    refNext = (!ubRef && ringHi + 1 < g.refLen) ? ref[ringHi + 1] : 0;
                }
            }
            int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            int f = 0, u = 0;
            hPrev[0] = 0; ePrev[0] = 0; hPrev[edge] = 0; ePrev[edge] = 0; hCur[0] = 0;     // ssw.c:633
            const int xi = band_x(bw, i), xp = band_x(bw, i - 1);
            uint32_t* line = dir + (size_t)i * g.strideW;
            uint32_t word = 0;
            const int rb = rbNext;
            rbNext = i + 1 < g.readLen ? read[i + 1] : 0;
            const unsigned long long rowTab = packRow ? s_rowTab[rb] : 0ull;
            if constexpr (LOCAL) {
                // Shared-memory rows: walk the slots with running pointers and carry the left (hCur[u-1]) and diagonal
                // (hPrev[up-1] = the previous cell's upper neighbour) values in registers.  The rows start zeroed, so the
                // reference's i == 0 special case (ssw.c:644-645) yields the same values and needs no branch.
                const int up0 = beg - xp + 1;                          // set_u(e, w, i-1, beg); u starts at 1, lf at 0
                const short* ph = hPrev.p + up0 * T;                   // hPrev[up]
                short* peUp = ePrev.p + up0 * T;                       // ePrev[up] (read)
                short* peU = ePrev.p + T;                              // ePrev[u]  (written)
                short* pc = hCur.p + T;                                // hCur[u]
                const unsigned char* pr = refRing;
                int slot = beg & (RING - 1);
                int hDiag = hPrev[up0 - 1], hLeft = 0;                 // hCur[0] = 0 (ssw.c:633)
                for (int j = beg; j <= end; ++j) {
                    const int hUp = *ph, eUp = *peUp;
                    int a = hUp - go, b = eUp - ge;                    // ssw.c:644-648
                    const int ev = a > b ? a : b;
                    *peU = (short)ev;
                    const int bitE = a > b ? 1 : 0;
                    a = hLeft - go; b = f - ge;                        // ssw.c:650-653
                    f = a > b ? a : b;
                    const int bitF = a > b ? 1 : 0;
                    const int e1 = ev > 0 ? ev : 0, f1 = f > 0 ? f : 0;   // ssw.c:655-659
                    const int gmax = e1 > f1 ? e1 : f1;
                    const int rc = pr[slot * T];
                    const int sc = packRow ? (int)(int8_t)(rowTab >> (8 * rc)) : (int)mat[rc * n + rb];
                    const int m = hDiag + sc;
                    const int h = gmax > m ? gmax : m;
                    *pc = (short)h;
                    if (h > best) best = h;                            // ssw.c:661
                    const int sel = gmax <= m ? 0 : (e1 > f1 ? 1 : 2); // ssw.c:663-664
                    const int x = j - xi;
                    word |= (uint32_t)(bitE | (bitF << 1) | (sel << 2)) << (4 * (x & 7));
                    if ((x & 7) == 7) { line[x >> 3] = word; word = 0; }
                    hDiag = hUp; hLeft = h;
                    ph += T; peUp += T; peU += T; pc += T; slot = (slot + 1) & (RING - 1);
                }
                u = end - xi + 1;
            } else
            for (int j = beg; j <= end; ++j) {
                u = j - xi + 1;                        // set_u(u, w, i, j)
                const int up = j - xp + 1;             // set_u(e, w, i-1, j)
                const int lf = u - 1;                  // set_u(b, w, i, j-1)
                const int dg = up - 1;                 // set_u(d, w, i-1, j-1)
                int a = i == 0 ? -go : hPrev[up] - go; // ssw.c:644-648
                int b = i == 0 ? -ge : ePrev[up] - ge;
                const int ev = a > b ? a : b;
                ePrev[u] = (LOCAL ? (short)ev : ev);
                const int bitE = a > b ? 1 : 0;
                a = hCur[lf] - go;                     // ssw.c:650-653
                b = f - ge;
                f = a > b ? a : b;
                const int bitF = a > b ? 1 : 0;
                const int e1 = ev > 0 ? ev : 0;        // ssw.c:655-659
                const int f1 = f > 0 ? f : 0;
                const int gmax = e1 > f1 ? e1 : f1;
                int rc;
                if constexpr (LOCAL) rc = refRing[(j & (RING - 1)) * T];
                else rc = ubRef ? 0 : ref[j];
                const int sc = packRow ? (int)(int8_t)(rowTab >> (8 * rc)) : (int)mat[rc * n + rb];
                const int m = hPrev[dg] + sc;
                const int h = gmax > m ? gmax : m;
                hCur[u] = (LOCAL ? (short)h : h);
                if (h > best) best = h;                // ssw.c:661
                const int sel = gmax <= m ? 0 : (e1 > f1 ? 1 : 2);       // ssw.c:663-664
                const int x = j - xi;
                word |= (uint32_t)(bitE | (bitF << 1) | (sel << 2)) << (4 * (x & 7));
                if ((x & 7) == 7) { line[x >> 3] = word; word = 0; }
            }
            if (end >= beg && ((end - xi) & 7) != 7) line[(end - xi) >> 3] = word;
            cells += end - beg + 1;
            // ssw.c:666 copies the row just computed (slots 1..u) into the previous-row buffer.  Every slot the next row
            // reads is either one of those or is zeroed at its start (ssw.c:633), so exchanging the two buffers is
            // equivalent and saves a shared-memory round trip per cell.
            { auto* tp = hPrev.p; hPrev.p = hCur.p; hCur.p = tp; }
        }
        if (best < score && bw * 2 <= len) { bw *= 2; continue; }       // ssw.c:668-669: widen and redo
        break;
    }
    warp_count(d.counters + CNT_CELLS_BAND, (unsigned long long)cells);

    // ---- traceback: one walk (ops buffered in registers), allocate, emit reversed ---------------------------------
    BandOps ops;
    const bool regRows = g.strideW <= BAND_ROWWORDS;
    const int l = regRows ? band_traceback<true>(dir, g, ops, nullptr, 0) : band_traceback<false>(dir, g, ops, nullptr, 0);
    if (l < 0) { r.flag = 1; r.cigar_len = 0; r.cigar_off = 0; d.p_state[p] |= PST_BAND_DONE; return; }      // ssw.c:911
    const unsigned long long coff = warp_bump(&d.bump[1], (unsigned long long)l);
    r.cigar_len = l; r.cigar_off = (int64_t)coff;
    if ((long long)coff + l > d.cigar_cap) { atomicAdd(d.counters + CNT_CIGAR_OVERFLOW, 1); return; }
    if (l <= BAND_OPBUF) {
#pragma unroll
        for (int q = 0; q < BAND_OPBUF; ++q) if (q < l) d.cigar[coff + (l - 1 - q)] = ops.op[q];   // reverse (ssw.c:753-762)
    } else {
        // more ops than the register buffer holds: walk again, writing straight into the arena
        if (regRows) band_traceback<true>(dir, g, ops, d.cigar + coff, l); else band_traceback<false>(dir, g, ops, d.cigar + coff, l);
    }
    d.p_state[p] |= PST_BAND_DONE;                     // the certificate pass (k_certify_rest) may now look at this pair
}
