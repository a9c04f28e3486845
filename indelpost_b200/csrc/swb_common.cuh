// swb_common.cuh — shared device/host definitions of libswb200 (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/swb200.h"

#define SWB_MAX_N 32            // largest substitution-matrix edge the kernels stage in shared memory
#define SWB_NBUCKETS 8          // fast-path read-length buckets: bucket b holds padded lengths <= 32*(b+1) (R = 2*(b+1) rows per thread)
#define SWB_NF8 6               // forward-only read-length buckets of the 8-thread-per-lane-pair sweep: padded lengths <= 32, 56, 80, 104, 128, 152 (R = 4 .. 19 rows per thread)
#define SWB_NFWD (SWB_NBUCKETS + SWB_NF8)   // forward list families: [0, 8) 16-thread groups, [8, 14) 8-thread groups
#define SWB_NLISTS 160
#define SWB_NCOUNTERS 200

// ---------------------------------------------------------------------------------------------
// Device-resident batch ("workspace").  One per context, grown on demand.
// Layout in HBM: structure-of-arrays indexed by pair; sequences stay in their de-duplicated
// tables (codes, 1 byte/base) and every pair points into them, so a locus' window is uploaded once.
// ---------------------------------------------------------------------------------------------
struct SwbDev {
    // tables
    int8_t*  reads;     int64_t* read_off; int32_t* read_len;
    int8_t*  windows;   int64_t* win_off;  int32_t* win_len;
    // raw per-pair inputs
    int32_t* pair_read; int32_t* pair_win; int32_t* ref_beg; int32_t* ref_len;   // ref_beg/ref_len may be null
    uint8_t* gap_open;  uint8_t* gap_ext;  int32_t* mask_len;                    // mask_len may be null
    int8_t*  mat;       // n*n
    // derived per-pair (k_prepare)
    int64_t* p_roff;    // read start in reads[]
    int64_t* p_woff;    // window start (incl. ref_beg) in windows[]
    int32_t* p_rlen;
    int32_t* p_wlen;
    int32_t* p_mask;
    uint8_t* p_mode;    // 0: result is byte-mode, 1: word-mode (semantic chosen by k_prepare for the fast path, final mode after forward)
    uint8_t* p_state;   // PST_* flags
    int32_t* p_csafe;   // fast path: first window column whose column maximum reaches 128+go+ge (wlen if none): up to there the 8-bit pass is exact Gotoh
    // results
    swb_result* res;
    // job lists (indices of pairs) + counters
    int32_t* list[SWB_NLISTS];
    int32_t* counters;  // [SWB_NCOUNTERS]
    // per-pair column maxima scratch for the sub-optimal score (u16, stride = max window length)
    uint16_t* colmax;   int32_t colmax_stride;
    // banded traceback scratch + CIGAR arena
    uint8_t*  band;     int64_t band_cap;
    uint32_t* cigar;    int64_t cigar_cap;
    unsigned long long* bump;   // [0]: band scratch bump pointer, [1]: cigar arena bump pointer
    int32_t*  t_bw;     // current band width per pair (banded rounds)
    int32_t*  t_best;   // running DP maximum per pair (not reset between widenings, ssw.c:600,661)
    int32_t   warp_next;   // list that takes the re-queued jobs fit for the warp-per-alignment kernel this round (-1: none)

    int32_t n_pairs, n_reads, n_windows;
    int32_t n;          // matrix edge
    int32_t bias;       // |min(mat)| (ssw.c:795-799)
    int8_t  score_size; uint8_t flag; uint16_t filters; int32_t filterd;
    int32_t seq_encoding;
    int32_t seq_shift;  // device blobs hold one code per byte; a caller offset o (bytes, packed input) maps to (o - byte_base) << seq_shift
    int32_t max_rlen, max_wlen;
    int32_t max_score;  // max(mat): bound on the score gained per read base
    int32_t fast_ok;    // batch-level eligibility of the DPX fast path (matrix range, n >= 4, score_size)
    int32_t fast_max_cols;
    int32_t fast8_ok;   // the 8-thread-group sweep may be used: every matrix entry in [-4, 3] (its scores are scaled by 32 into one signed byte)
    // chunk view: the uploaded tables are slices of the caller's tables (pipelined swb_align_batch)
    int32_t ridx_base, widx_base;     // first read / window index present in the slice
    int32_t n_reads_total, n_windows_total;
    int64_t rbyte_base, wbyte_base;   // byte offset of the slice inside the caller's blobs
    uint32_t* fast_cols;   // global column-best storage of the fast path when windows are too long for shared memory (null otherwise)
    int32_t one;        // always 1, but opaque to the compiler: x * one + c compiles to a real IMAD (FMA pipe) instead of an ALU-pipe add
    int32_t opt;        // experiment switches (SWB200_OPT): bit0 = certificate inline in k_band instead of the separate pass, bit1 = scalar-lane exact kernel, bit2 = no banded reverse pass, bit4 = no register-band kernel, bit5 = single traceback phase, bit6 = no warp-per-alignment band kernel, bit7 = no sandwich sweep (unsafe-zone 8-bit pairs and overflow verifications go to the exact kernels)
};

// list[] slots; counters[i] is the length of list[i] for i < SWB_NLISTS
#define SWB_NBANDCLASS 8         // band jobs are bucketed by the rolling-buffer slots they need, see band_class()
#define SWB_BAND_CLS_MID 4      // first class of k_band<48,64>; 5: k_band<112,32>; 6: k_band<254,32>; 7: k_band<0,128> (see band_class)
// LIST_BANDW*: register-band jobs per exact half-width 1..SWB_BANDW_MAX (swb_bandreg.cuh); _FIRST: the traceback phase that is
// certified first; _NEXT: jobs that kernel widened once.  LIST_BANDWARP*: wide bands, one warp per alignment (swb_bandwarp.cuh);
// _NEXT is two lists used alternately by the re-queue rounds.
enum { LIST_BYTE_FWD = 0, LIST_WORD_FWD = 1, LIST_BYTE_REV = 2, LIST_WORD_REV = 3, LIST_VERIFY = 4, LIST_VERIFY2 = 5,
       LIST_BYTE_REV2 = 6, LIST_WORD_REV2 = 7,      // fast-path pairs handed to the exact reverse pass (the exact path itself uses LIST_*_REV, concurrently)
       LIST_FAST_FWD = 8, LIST_FAST_REV = 16, LIST_BAND = 24, LIST_BAND_NEXT = 32, LIST_BAND_FIRST = 40, LIST_REVB = 48,
       LIST_BANDW = 56, LIST_BANDW_FIRST = 80, LIST_BANDW_NEXT = 104,
       LIST_BANDWARP = 128, LIST_BANDWARP_FIRST = 129, LIST_BANDWARP_NEXT = 130,
       // sandwich sweep (swb_fast.cuh, SW = 1): 8-bit-final pairs whose scores can pass 128+go+ge, per read-length bucket, forward / reverse;
       // LIST_VERIFYX (+1): pairs whose overflow neither the CIGAR certificate nor the sandwich lower bound could prove (exact 8-bit pass)
       LIST_SW_FWD = 132, LIST_SW_REV = 140, LIST_VERIFYX = 148,
       LIST_F8_FWD = 150,
       LIST_BANDQ_TMP = 156, LIST_BANDW_TMP = 157,
       LIST_LATE = 158 };       // one-shot path: pairs whose records change after the first traceback round (early download)      // a round's warp-kernel jobs split by schedule (swb_bandwarp.cuh): 8 lanes / 32 lanes per alignment      // SWB_NF8 lists: forward sweep with 8 threads per lane pair (swb_fast.cuh, G = 8)
enum { CNT_BYTE_FWD = 0, CNT_WORD_FWD = 1, CNT_BYTE_REV = 2, CNT_WORD_REV = 3,
       CNT_FAST_FWD = 8, CNT_FAST_REV = 16, CNT_BAND = 24, CNT_BAND_NEXT = 32,
       CNT_SW_FWD = 132, CNT_SW_REV = 140, CNT_F8_FWD = 150,
       CNT_CELLS_FWD = 160, CNT_CELLS_REV = 162, CNT_CELLS_BAND = 164, CNT_BAND_OVERFLOW = 166, CNT_CIGAR_OVERFLOW = 167,
       CNT_FAST_DONE = 168, CNT_CERT_FAIL = 169, CNT_VERIFY_BYTE = 170, CNT_EXACT_JOBS = 171,
       CNT_SW_CERTIFIED = 172,     // unsafe-zone pairs the sandwich certified (forward); CNT_SW_REJECTED: those it sent to the exact path
       CNT_SW_REJECTED = 173, CNT_SW_VERIFIED = 174,   // overflow verifications settled by the sandwich lower bound
       CNT_FAST_MAXCOLS = 176 };   // [SWB_NFWD] longest window among the fast-path pairs of each forward list family
#define SWB_BANDW_MAX 24           // widest half-width the register-band kernel is instantiated for
#define SWB_BANDREG_MAXROWS 320    // longest read segment it stages in shared memory
// banded reverse pass (swb_revband.cuh): band classes as (rows below, columns right of) the main diagonal
#define SWB_NREVB 6
#define SWB_REVB_CLASSES(X) X(0, 2, 5) X(1, 3, 12) X(2, 5, 18) X(3, 6, 25) X(4, 10, 37) X(5, 13, 50)
__host__ __device__ __forceinline__ int revb_class(int wi, int wd) {
#define SWB_REVB_PICK(c, WI, WD) if (wi <= WI && wd <= WD) return c;
    SWB_REVB_CLASSES(SWB_REVB_PICK)
#undef SWB_REVB_PICK
    return -1;
}
// bucket of the 8-thread-group forward sweep for a padded read length (rows the result counts: 8-padded in 16-bit semantics, 16-padded in
// 8-bit semantics); -1: too long
__host__ __device__ __forceinline__ int f8_bucket(int lp) { return lp <= 32 ? 0 : lp <= 56 ? 1 : lp <= 80 ? 2 : lp <= 104 ? 3 : lp <= 128 ? 4 : lp <= 152 ? 5 : -1; }
// p_state flags
enum { PST_FAST = 1,        // forward result produced by the DPX fast path
       PST_NEED_CERT = 2,   // word-mode result accepted provisionally: the 8-bit pass still has to be shown to overflow
       PST_HAVE_WORD = 4,   // 16-bit result already in place: the exact 8-bit pass only verifies the overflow
       PST_BAND_DONE = 8 }; // traceback finished (or no CIGAR requested): the certificate may look at the pair

__device__ __forceinline__ void list_push(int32_t* list, int32_t* counter, int32_t v) {
    // warp-aggregated append; lanes are grouped by destination list (one call site may feed several lists)
    const unsigned act = __activemask();
    const unsigned m = __match_any_sync(act, (unsigned long long)(uintptr_t)counter);
    const int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    list[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1))] = v;
}

// warp-aggregated bump allocation: the active lanes of a warp get consecutive ranges from ONE atomicAdd
__device__ __forceinline__ unsigned long long warp_bump(unsigned long long* ptr, unsigned long long need) {
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31;
    unsigned long long pre = 0, tot = 0;
    for (unsigned rest = m; rest; rest &= rest - 1) {
        const int src = __ffs(rest) - 1;
        const unsigned long long v = __shfl_sync(m, need, src);
        if (src < lane) pre += v;
        tot += v;
    }
    const int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(ptr, tot);
    base = __shfl_sync(m, base, leader);
    return base + pre;
}

// warp-aggregated 64-bit statistics counter
__device__ __forceinline__ void warp_count(int32_t* counter64, unsigned long long v) {
    const unsigned m = __activemask();
    unsigned long long tot = 0;
    for (unsigned rest = m; rest; rest &= rest - 1) tot += __shfl_sync(m, v, __ffs(rest) - 1);
    if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(reinterpret_cast<unsigned long long*>(counter64), tot);
}

// f(idx, byte) for every byte of the sequence [ptr, ptr + len), read as aligned 16-byte chunks (one thread walks its own
// sequence: a dozen independent vector loads instead of a dependent byte stream).  Reads up to 15 bytes before / after
// the sequence: every sequence blob is a cudaMalloc allocation (256-byte aligned) with 16 bytes of slack at its end.
template <typename F>
__device__ __forceinline__ void for_each_byte16(const int8_t* ptr, int len, F f) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
    const int mis = (int)(a & 15);
    const uint4* base = reinterpret_cast<const uint4*>(a - mis);
    const int nch = (mis + len + 15) >> 4;
    // four chunks in flight: the thread-per-alignment kernels run at low occupancy, so the staging is as long as its loads' latency
    for (int ch0 = 0; ch0 < nch; ch0 += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (ch0 + u < nch) v[u] = __ldg(base + ch0 + u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (ch0 + u >= nch) break;
            const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
            const int i0 = (ch0 + u) * 16 - mis;
            if (i0 >= 0 && i0 + 16 <= len) {
#pragma unroll
                for (int q = 0; q < 16; ++q) f(i0 + q, (w[q >> 2] >> (8 * (q & 3))) & 0xffu);
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) if ((unsigned)(i0 + q) < (unsigned)len) f(i0 + q, (w[q >> 2] >> (8 * (q & 3))) & 0xffu);
            }
        }
    }
}

// two sequences at once: the loads of both are issued before either is consumed (eight 16-byte loads in flight)
template <typename FA, typename FB>
__device__ __forceinline__ void for_each_byte16_pair(const int8_t* pa, int la, FA fa, const int8_t* pb, int lb, FB fb) {
    const uintptr_t aa = reinterpret_cast<uintptr_t>(pa), ab = reinterpret_cast<uintptr_t>(pb);
    const int ma = (int)(aa & 15), mb = (int)(ab & 15);
    const uint4* ba = reinterpret_cast<const uint4*>(aa - ma);
    const uint4* bb = reinterpret_cast<const uint4*>(ab - mb);
    const int na = (ma + la + 15) >> 4, nb = (mb + lb + 15) >> 4;
    const int nmax = na > nb ? na : nb;
    for (int ch0 = 0; ch0 < nmax; ch0 += 4) {
        uint4 va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { if (ch0 + u < na) va[u] = __ldg(ba + ch0 + u); if (ch0 + u < nb) vb[u] = __ldg(bb + ch0 + u); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (ch0 + u < na) {
                const uint32_t w[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
                const int i0 = (ch0 + u) * 16 - ma;
#pragma unroll
                for (int q = 0; q < 16; ++q) if ((unsigned)(i0 + q) < (unsigned)la) fa(i0 + q, (w[q >> 2] >> (8 * (q & 3))) & 0xffu);
            }
            if (ch0 + u < nb) {
                const uint32_t w[4] = {vb[u].x, vb[u].y, vb[u].z, vb[u].w};
                const int i0 = (ch0 + u) * 16 - mb;
#pragma unroll
                for (int q = 0; q < 16; ++q) if ((unsigned)(i0 + q) < (unsigned)lb) fb(i0 + q, (w[q >> 2] >> (8 * (q & 3))) & 0xffu);
            }
        }
    }
}

// band job class from the half-width (see SWB_NBANDCLASS)
// The rolling buffers of banded_sw are 2*bw+3 slots wide, but no slot beyond refLen+1 is ever read (set_u gives u <= j+1 <= refLen and
// the upper neighbour is u or u+1): a job needs min(2*bw+4, refLen+3) slots, which decides the instantiation of k_band that can
// keep its rows in shared memory -- also for the very wide bands a free gap extension produces (bw > refLen).
//   classes 0-3: rows <= 36  (k_band<16,128>; bucketed by width 1-2 | 3-4 | 5-8 | wider so a warp runs similar bands)
//   class 4: rows <= 100 (k_band<48,64>)   5: <= 228 (k_band<112,32>)   6: <= 512 (k_band<254,32>)   7: global-memory rows
__host__ __device__ __forceinline__ int band_rows_needed(int bw, int refLen) { const int a = 2 * bw + 4, b = refLen + 3; return a < b ? a : b; }
__host__ __device__ __forceinline__ int band_class(int bw, int refLen, bool fits16 = true) {
    if (!fits16) return 7;
    const int need = band_rows_needed(bw, refLen);
    if (need <= 36) return bw <= 2 ? 0 : bw <= 4 ? 1 : bw <= 8 ? 2 : 3;
    return need <= 100 ? 4 : need <= 228 ? 5 : need <= 512 ? 6 : 7;
}

// bands the warp-per-alignment kernel (swb_bandwarp.cuh) takes: wider than the register-band kernel's, and either narrower than the
// matrix (regular) or never sliding
__device__ __forceinline__ bool bandwarp_ok(const SwbDev& d, int bw, int refLen, int readLen, const swb_result& r) {
    return bw > SWB_BANDW_MAX && (refLen >= 2 * bw + 2 || bw >= readLen - 1) && refLen <= 512 && readLen <= SWB_BANDREG_MAXROWS && d.n <= 8 &&
           r.ref_begin1 >= 0 && r.read_begin1 >= 0 && !(d.opt & 64);
}

// re-queue a job whose band has grown to bw (t_bw / t_best already stored) for the next round
__device__ __forceinline__ void requeue_band(const SwbDev& d, int nextBase, int p, int bw, int refLen, int readLen, const swb_result& r, bool fits16 = true) {
    if (d.warp_next >= 0 && bandwarp_ok(d, bw, refLen, readLen, r)) { list_push(d.list[d.warp_next], d.counters + d.warp_next, p); return; }
    const int c = band_class(bw, refLen, fits16);
    list_push(d.list[nextBase + c], d.counters + nextBase + c, p);
}

// queue a pair for the banded traceback in the class of its initial band width |refLen - readLen| + 1 (ssw.c:899)
__device__ __forceinline__ void push_band(const SwbDev& d, int p, const swb_result& r) {
    const int refLen = r.ref_end1 - r.ref_begin1 + 1, readLen = r.read_end1 - r.read_begin1 + 1;
    const int dl = refLen - readLen;
    const int bw = (dl < 0 ? -dl : dl) + 1;
    // provisional 16-bit results whose alignment has a net insertion are the only ones that can fail the overflow
    // certificate (swb_cert.cuh): they are traced back first so their verification overlaps the rest of the stage
    const bool first = (d.p_state[p] & PST_NEED_CERT) && dl < 0 && !(d.opt & 32);
    // regular jobs (band narrower than the matrix, see swb_bandreg.cuh) go to the register-band kernel of their exact width
    if (bw <= SWB_BANDW_MAX && refLen >= 2 * bw + 2 && readLen <= SWB_BANDREG_MAXROWS && d.n <= 8 && r.ref_begin1 >= 0 && r.read_begin1 >= 0 && !(d.opt & 16)) {
        const int slot = (first ? LIST_BANDW_FIRST : LIST_BANDW) + bw - 1;
        list_push(d.list[slot], d.counters + slot, p);
        return;
    }
    // wide regular bands (what a free gap extension produces): one warp per alignment
    if (bandwarp_ok(d, bw, refLen, readLen, r)) {
        const int slot = first ? LIST_BANDWARP_FIRST : LIST_BANDWARP;
        list_push(d.list[slot], d.counters + slot, p);
        return;
    }
    const int c = band_class(bw, refLen, (long long)d.max_score * (readLen > 0 ? readLen : 1) <= 30000);
    const int base = first ? LIST_BAND_FIRST : LIST_BAND;
    list_push(d.list[base + c], d.counters + base + c, p);
}

// sswpy.pyx:16-29 DNA_BASE_LUT
__host__ __device__ __forceinline__ int8_t swb_dna_code(unsigned char c) {
    switch (c) {
        case 'A': case 'a': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}
