// swb_host.h — host-side context of libswb200 shared by the translation units (one per kernel family, so that the
// instantiations compile in parallel; device code never crosses a unit: every kernel is launched by the unit that defines it).
#pragma once
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <algorithm>
#include <atomic>
#include <thread>
#include <climits>
#include <chrono>
#include "swb_common.cuh"

#define SWB_MAX_DEVICES 64
#define SWB_NSIDE 6
// fast path: windows longer than this keep their column bests in global memory (written once per step by a group's last thread) and
// only the 2-byte selectors in shared memory -- beyond it the per-column shared memory, not the registers, would set the occupancy
// (four 128-thread blocks per SM: 8 groups x 10 bytes or 16 groups x 8 bytes per column)
#define SWB_FAST_SMEM_COLS 704
#define SWB_FAST8_SMEM_COLS 448
static inline int fast_smem_cols(int fam) { return fam >= SWB_NBUCKETS ? SWB_FAST8_SMEM_COLS : SWB_FAST_SMEM_COLS; }

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

enum { EV_START = 0, EV_PREP, EV_FWD, EV_REV, EV_BAND, EV_H2D0, EV_H2D1, EV_D2H0, EV_D2H1, EV_BAND_R0, EV_BAND_ALL, EV_COUNT };

#define SWB_MAX_PARTS 4
// a contiguous range of the batch's pairs with its own job lists and counters (see compute_setup)
struct Part { int32_t p0 = 0, p1 = 0; int32_t* lists = nullptr; size_t perList = 0; int32_t* counters = nullptr; };

struct swb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;   // main; wide band classes; overflow verification
    cudaStream_t stream4 = nullptr; cudaEvent_t ev_join3; cudaEvent_t ev_x_fork, ev_x_join;     // exact path beside the fast reverse pass (on stream3)
    unsigned verify_pending = 0;   // second verification stream
    cudaStream_t bulk_stream = nullptr;                                     // lowest priority: the forward DPX sweep (yields SM slots to the short, latency-bound kernels of the other lane)
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_copy;              // streamed path: host->device copies back to back on their own stream
    cudaStream_t bulk_stream2 = nullptr;                                    // second one: consecutive forward slices of the streamed one-shot path overlap their tails
    cudaEvent_t ev_bulk_fork, ev_bulk_join, ev_bulk_join2, ev_piece;
    cudaEvent_t ev_fork, ev_join, ev_join2, ev_fork3;
    cudaStream_t side_stream[SWB_NSIDE] = {}; cudaEvent_t ev_side_join[SWB_NSIDE], ev_side_split;      // wide band kernels of a traceback round, one stream each
    cudaStream_t rev_stream[SWB_NREVB];                                     // banded reverse pass: one stream per band class
    cudaEvent_t ev_rev_fork, ev_rev_join[SWB_NREVB];
    cudaStream_t bandw_stream[SWB_BANDW_MAX]; cudaEvent_t ev_bandw_join[SWB_BANDW_MAX];   // register-band kernels: one stream per half-width
    unsigned bandreg_used = 0; int bandreg_base = 0;                       // side streams the register-band kernels of the current round run on
    cudaEvent_t ev[EV_COUNT];
    std::string err;
    SwbDev d;
    bool have_batch = false, computed = false;
    // device buffers
    DevBuf b_reads_pk, b_windows_pk;              // packed input (SWB_SEQ_PACKED4 / PACKED2) as uploaded, before k_unpack
    DevBuf b_reads, b_read_off, b_read_len, b_windows, b_win_off, b_win_len;
    DevBuf b_pair_read, b_pair_win, b_ref_beg, b_ref_len, b_go, b_ge, b_mask, b_mat;
    DevBuf b_roff, b_woff, b_rlen, b_wlen, b_pmask, b_mode, b_res, b_lists, b_counters, b_colmax, b_band, b_cigar, b_bump;
    DevBuf b_tbw, b_tbest, b_rbad, b_wbad, b_state, b_csafe, b_fastcols;
    DevBuf b_ind_off, b_ind_cnt, b_ind_rend, b_ind_recs, b_ind_misc, b_ind_cig, b_ind_coff, b_ind_clen, b_ind_rs, b_ind_qs;   // indel extraction
    // early download of the one-shot path: the caller's output arrays, known while the traceback rounds are still running
    struct { bool active = false, started = false; swb_result* results = nullptr; uint32_t* arena = nullptr; int64_t cap = 0, arena_done = 0; } early;
    cudaEvent_t ev_early = nullptr; DevBuf b_late; swb_result* h_late_rec = nullptr; int32_t* h_late_idx = nullptr; size_t h_late_cap = 0;
    Part parts[SWB_MAX_PARTS]; int nparts = 1, cur_part = 0, force_parts = 0;
    cudaEvent_t ev_part_fwd[SWB_MAX_PARTS] = {}; cudaEvent_t ev_fwd_end = nullptr;      // behind a part's forward sweeps / behind the last one (timed)
    int fastMaxCols[SWB_NFWD] = {};              // longest window per forward list family
    int swCounts[SWB_NBUCKETS] = {};             // sandwich forward list lengths of the current batch
    int32_t* h_counters = nullptr;              // pinned mirror of counters
    int32_t* h_snap[2] = {nullptr, nullptr};    // streamed path: counter snapshots of the piece in flight and the one before
    cudaEvent_t ev_snap[2];
    unsigned long long* h_bump = nullptr;       // pinned mirror of bump
    swb_timing tm;
    int smem_optin = 0;
    int n_sm = 0;
    int64_t chunk_pairs = 0;
    std::vector<std::pair<const char*, double>> trace;
    swb_ctx* sibling = nullptr;                 // second lane, created on demand by the pipelined swb_align_batch
    bool pipelined_last = false;
};

// SWB200_TRACE=1: host wall-clock marks (after the host-side synchronisation points) dumped to stderr per swb_align_batch call
extern const bool g_trace;
static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TR(ctx, name) do { if (g_trace) (ctx)->trace.push_back(std::make_pair((const char*)(name), now_ms())); } while (0)

#define CUDA_TRY(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } } while (0)

// SWB200_DEBUG_SYNC=1: synchronise after every stage and name the one that faulted
static inline int stage_check(swb_ctx* c, const char* name) {
    static const bool dbg = getenv("SWB200_DEBUG_SYNC") != nullptr;
    if (!dbg) return 0;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { c->err = std::string("stage ") + name + ": " + cudaGetErrorString(e); fprintf(stderr, "libswb200: %s\n", c->err.c_str()); return -1; }
    return 0;
}

// ---- launchers, one translation unit per kernel family ------------------------------------------------------------------
// swb_l_exact.cu: k_exact / k_exact2 (MODE 0 byte, 1 word; DIR 0 forward, 1 reverse)
int swb_launch_exact(swb_ctx* c, int mode, int dir, int listSlot, int upperBound, cudaStream_t st, bool fewJobsLikely);
cudaError_t swb_exact_set_attrs(int smem_optin);
// swb_l_fast.cu (compiled once per direction): one launch per non-empty read-length bucket over the list ranges [first[b], counts[b])
int swb_launch_fast_range_fwd(swb_ctx* c, const int* first, const int* counts, cudaStream_t st);
int swb_launch_fast_range_rev(swb_ctx* c, const int* first, const int* counts, cudaStream_t st);
// (-DSWB_FAST_G8) the forward sweep with 8 threads per lane pair; first / counts hold SWB_NFWD entries, this family reads [SWB_NBUCKETS, SWB_NFWD)
int swb_launch_fast8_range_fwd(swb_ctx* c, const int* first, const int* counts, cudaStream_t st);
// (-DSWB_FAST_SW=1) the sandwich flavour: 8-bit-final pairs whose scores can pass 128+go+ge, and overflow verifications
int swb_launch_sandwich_fwd(swb_ctx* c, const int* counts, cudaStream_t st);
int swb_launch_sandwich_rev(swb_ctx* c, const int* counts, cudaStream_t st);
int swb_launch_sandwich_verify(swb_ctx* c, int listSlot, int xSlot, int upperBound, cudaStream_t st);
// swb_l_revband.cu
int swb_launch_rev_band(swb_ctx* c, int upperBoundPairs);
// swb_l_bandreg.cu (compiled in two halves of the width range): the register-band kernel of half-width w
int swb_launch_band_reg_lo(swb_ctx* c, int w, int listSlot, int njobs, int nextBase, int nextBaseW, int resume, cudaStream_t st);
int swb_launch_band_reg_hi(swb_ctx* c, int w, int listSlot, int njobs, int nextBase, int nextBaseW, int resume, cudaStream_t st);
// swb_l_band.cu: the literal banded_sw kernel (which = 0: rows in global memory, 1: local, 2: mid, 3: wide, 4: huge) and the warp-per-alignment one
int swb_launch_band(swb_ctx* c, int which, int blocks, int listBase, int firstClass, int lastClass, int nextBase, cudaStream_t st);
int swb_launch_band_warp(swb_ctx* c, int listSlot, int njobs, int nextBase, cudaStream_t st, cudaStream_t st2, cudaEvent_t evSplit);
cudaError_t swb_band_set_attrs();
