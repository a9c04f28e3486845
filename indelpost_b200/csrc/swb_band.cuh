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
// shared memory laid out [slot][thread], which is bank-conflict free whatever slot each thread touches.  Band doubling is done in
// rounds: a pair whose DP maximum is still below score1 is re-queued with twice the width (about 0.5 %
// of realistic pairs).  The traceback is walked twice (count, then emit) so the CIGAR
// can be written, already reversed, straight into the output arena.
#pragma once
#include "swb_common.cuh"

#define SWB_BAND_LOCAL_BW 16                         // bands up to this half-width keep their rows in shared memory
#define SWB_BAND_LOCAL_W (2 * SWB_BAND_LOCAL_BW + 4)
#define SWB_BAND_THREADS 128
#define SWB_BAND_REFRING 64                           // circular buffer of window bases per thread (>= band width)
#define SWB_BAND_SMEM (3 * SWB_BAND_LOCAL_W * SWB_BAND_THREADS * 2 + SWB_BAND_REFRING * SWB_BAND_THREADS)
#define SWB_BAND_MAX16 30000                          // largest score the 16-bit shared-memory rows may hold

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

// walk the traceback; if out != nullptr write the ops reversed into out[0..total)
__device__ __forceinline__ int band_traceback(const uint32_t* dir, const BandGeom& g, uint32_t* out, int total) {
    int i = g.readLen - 1, j = g.refLen - 1;
    int e = 0, l = 0, state = 2;
    int op = 0, prev_op = 0;                           // 0 M, 1 I, 2 D  (BAM op codes)
    while (i >= 0 && j > 0) {                          // ssw.c:679
        const int dv = band_dir_at(dir, g, i, j, state);
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
            ++l;
            if (out) out[total - l] = ((uint32_t)e << 4) | (uint32_t)prev_op;
            prev_op = op; e = 1;
        }
    }
    if (op == 0) {                                     // ssw.c:734-751
        ++l;
        if (out) out[total - l] = ((uint32_t)(e + 1) << 4) | 0u;
    } else {
        l += 2;
        if (out) { out[total - (l - 1)] = ((uint32_t)e << 4) | (uint32_t)op; out[total - l] = (1u << 4) | 0u; }
    }
    return l;
}

// rolling-row accessor: shared memory [slot][thread] (LOCAL) or a private global array (wide bands)
template <bool LOCAL> struct BandRow;
template <> struct BandRow<true>  { short* p; __device__ __forceinline__ short& operator[](int u) const { return p[u * SWB_BAND_THREADS]; } };
template <> struct BandRow<false> { int* p;   __device__ __forceinline__ int& operator[](int u) const { return p[u]; } };

template <bool LOCAL>
__global__ void __launch_bounds__(SWB_BAND_THREADS)
k_band(SwbDev d, const int32_t* __restrict__ jobs, const int32_t* __restrict__ njobs_ptr, int32_t* nextList, int32_t* nextCount, int round)
{
    const int njobs = *njobs_ptr;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= njobs) return;
    const int p = jobs[t];
    swb_result& r = d.res[p];

    BandGeom g;
    g.refLen = r.ref_end1 - r.ref_begin1 + 1;          // ssw.c:897-899
    g.readLen = r.read_end1 - r.read_begin1 + 1;
    int bw, best;
    if (round == 0 || d.t_bw[p] == 0) { int dl = g.refLen - g.readLen; bw = (dl < 0 ? -dl : dl) + 1; best = 0; }
    else { bw = d.t_bw[p]; best = d.t_best[p]; }
    const bool wantLocal = bw <= SWB_BAND_LOCAL_BW && (long long)d.max_score * (g.readLen > 0 ? g.readLen : 1) <= SWB_BAND_MAX16;
    if (wantLocal != LOCAL) return;                    // the other instantiation handles it
    g.w = bw; g.width_d = 2 * bw + 1; g.strideW = (g.width_d + 7) >> 3;
    const int width = 2 * bw + 3;
    const int len = g.refLen > g.readLen ? g.refLen : g.readLen;
    const int score = r.score1;
    const int go = d.gap_open[p], ge = d.gap_ext[p];
    const int n = d.n;
    const int8_t* mat = d.mat;
    // score1 == 0 in byte mode leaves ref_begin1 == -1 and the reference reads ref[-1] (undefined);
    // the 1x1 problem it then solves gives "1M" whatever that byte is
    const bool ubRef = r.ref_begin1 < 0;
    const int8_t* ref = d.windows + d.p_woff[p] + (ubRef ? 0 : r.ref_begin1);
    const int8_t* read = d.reads + d.p_roff[p] + (r.read_begin1 < 0 ? 0 : r.read_begin1);

    // ---- scratch: packed direction words (+ the three rolling rows when the band is too wide for shared memory)
    const long long dirBytes = (long long)g.strideW * 4 * (g.readLen > 0 ? g.readLen : 1);
    const long long rowBytes = LOCAL ? 0 : 3ll * (width + 1) * 4;
    const long long need = dirBytes + rowBytes;
    d.t_bw[p] = bw; d.t_best[p] = best;
    const unsigned long long off = atomicAdd(&d.bump[0], (unsigned long long)need);
    if ((long long)off + need > d.band_cap) {          // out of scratch: retry in the next round
        atomicAdd(d.counters + CNT_BAND_OVERFLOW, 1);
        list_push(nextList, nextCount, p);
        return;
    }
    uint32_t* dir = reinterpret_cast<uint32_t*>(d.band + off);
    extern __shared__ __align__(16) unsigned char band_smem[];
    BandRow<LOCAL> hPrev, ePrev, hCur;
    unsigned char* refRing = nullptr;                  // LOCAL: window bases of the current band, slot = column & 63
    if constexpr (LOCAL) {
        hPrev.p = reinterpret_cast<short*>(band_smem) + threadIdx.x;
        ePrev.p = hPrev.p + SWB_BAND_LOCAL_W * SWB_BAND_THREADS;
        hCur.p = ePrev.p + SWB_BAND_LOCAL_W * SWB_BAND_THREADS;
        refRing = band_smem + 3 * SWB_BAND_LOCAL_W * SWB_BAND_THREADS * 2 + threadIdx.x;
    } else {
        hPrev.p = reinterpret_cast<int*>(d.band + off + dirBytes);
        ePrev.p = hPrev.p + (width + 1); hCur.p = ePrev.p + (width + 1);
    }
    // substitution scores of one read base against every window base, packed 8 x int8 (n <= 8)
    const bool packRow = n <= 8;
    // the reference's buffers are realloc'ed across widenings and not cleared; every slot it reads is
    // written first within an iteration except where it reads uninitialised memory -- start from zeros
    for (int j = 0; j <= width; ++j) { hPrev[j] = 0; ePrev[j] = 0; hCur[j] = 0; }

    long long cells = 0;
    int ringHi = -1;                                   // last window column already in refRing
    for (int i = 0; i < g.readLen; ++i) {              // ssw.c:628-667
        const int beg = g.beg(i), end = g.end(i);
        if constexpr (LOCAL) {
            for (; ringHi < end; ) { ++ringHi; refRing[(ringHi & (SWB_BAND_REFRING - 1)) * SWB_BAND_THREADS] = (unsigned char)(ubRef ? 0 : ref[ringHi]); }
        }
        int edge = end + 1 < width - 1 ? end + 1 : width - 1;
        int f = 0, u = 0;
        hPrev[0] = 0; ePrev[0] = 0; hPrev[edge] = 0; ePrev[edge] = 0; hCur[0] = 0;     // ssw.c:633
        const int xi = band_x(bw, i), xp = band_x(bw, i - 1);
        uint32_t* line = dir + (size_t)i * g.strideW;
        uint32_t word = 0;
        const int rb = read[i];
        unsigned long long rowTab = 0;
        if (packRow) { for (int nt = 0; nt < n; ++nt) rowTab |= (unsigned long long)(uint8_t)mat[nt * n + rb] << (8 * nt); }
        for (int j = beg; j <= end; ++j) {
            u = j - xi + 1;                            // set_u(u, w, i, j)
            const int up = j - xp + 1;                 // set_u(e, w, i-1, j)
            const int lf = u - 1;                      // set_u(b, w, i, j-1)
            const int dg = up - 1;                     // set_u(d, w, i-1, j-1)
            int a = i == 0 ? -go : hPrev[up] - go;     // ssw.c:644-648
            int b = i == 0 ? -ge : ePrev[up] - ge;
            const int ev = a > b ? a : b;
            ePrev[u] = (LOCAL ? (short)ev : ev);
            const int bitE = a > b ? 1 : 0;
            a = hCur[lf] - go;                         // ssw.c:650-653
            b = f - ge;
            f = a > b ? a : b;
            const int bitF = a > b ? 1 : 0;
            const int e1 = ev > 0 ? ev : 0;            // ssw.c:655-659
            const int f1 = f > 0 ? f : 0;
            const int gmax = e1 > f1 ? e1 : f1;
            int rc;
            if constexpr (LOCAL) rc = refRing[(j & (SWB_BAND_REFRING - 1)) * SWB_BAND_THREADS];
            else rc = ubRef ? 0 : ref[j];
            const int sc = packRow ? (int)(int8_t)(rowTab >> (8 * rc)) : (int)mat[rc * n + rb];
            const int m = hPrev[dg] + sc;
            const int h = gmax > m ? gmax : m;
            hCur[u] = (LOCAL ? (short)h : h);
            if (h > best) best = h;                    // ssw.c:661
            const int sel = gmax <= m ? 0 : (e1 > f1 ? 1 : 2);       // ssw.c:663-664
            const int x = j - xi;
            word |= (uint32_t)(bitE | (bitF << 1) | (sel << 2)) << (4 * (x & 7));
            if ((x & 7) == 7) { line[x >> 3] = word; word = 0; }
        }
        if (end >= beg && ((end - xi) & 7) != 7) line[(end - xi) >> 3] = word;
        cells += end - beg + 1;
        for (int j = 1; j <= u; ++j) hPrev[j] = hCur[j];             // ssw.c:666
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(d.counters + CNT_CELLS_BAND), (unsigned long long)cells);

    if (best < score && bw * 2 <= len) {               // ssw.c:668-669: widen and redo
        d.t_bw[p] = bw * 2; d.t_best[p] = best;
        list_push(nextList, nextCount, p);
        return;
    }

    // ---- traceback: count, allocate, emit ---------------------------------------------------------
    const int l = band_traceback(dir, g, nullptr, 0);
    if (l < 0) { r.flag = 1; r.cigar_len = 0; r.cigar_off = 0; return; }      // ssw.c:911
    const unsigned long long coff = atomicAdd(&d.bump[1], (unsigned long long)l);
    r.cigar_len = l; r.cigar_off = (int64_t)coff;
    if ((long long)coff + l > d.cigar_cap) { atomicAdd(d.counters + CNT_CIGAR_OVERFLOW, 1); return; }
    band_traceback(dir, g, d.cigar + coff, l);
}
