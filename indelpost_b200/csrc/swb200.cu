// swb200.cu — libswb200: host pipeline + C ABI (include/swb200.h).
//
// One swb_ctx per GPU: streams, timing events and grow-only device buffers.  A batch moves through
//   upload  : one cudaMemcpyAsync per input array (tables + per-pair arrays)
//   prepare : encode/validate sequences, derive per-pair pointers, route each pair to the DPX fast path
//             (per read-length bucket) or to the exact striped emulation, seed the job lists
//   forward : score / end position / sub-optimal score        (ssw.c:842-871)   k_fast<R,0>, k_exact2<.,0>
//   reverse : begin position                                   (ssw.c:875-891)   k_fast<R,1>, k_exact2<.,1>
//   band    : banded DP + traceback -> CIGAR                   (ssw.c:897-916)   k_band, in two phases so that the
//             overflow certificate (swb_cert.cuh) of the pairs that can fail it and their exact 8-bit
//             verification overlap the bulk of the traceback
//   download: results + CIGAR arena
// Stage-to-stage hand-over is through device-side job lists (append with warp-aggregated atomics), so
// the host only synchronises to read the few counters that size the next launch.  swb_align_batch additionally
// pipelines large batches over two such contexts ("lanes") with chunk views of the caller's tables.
// There is no CPU implementation behind any entry point: without a usable GPU every call fails.
#include "swb_host.h"
#define SWB_WITH_CERT_KERNEL          // k_certify_rest is defined (and launched) by this unit only
#include "swb_cert.cuh"
#include "swb_band.cuh"               // launch geometry constants only: the band kernels are instantiated by swb_l_band.cu
#include "swb_exact2.cuh"             // SWB_EXACT_SPARSE_MAX
#include "swb_bandreg.cuh"            // SWB_BANDREG_* (instantiated by swb_l_bandreg.cu)
#include "swb_indels.cuh"

#define SWB_VERSION "swb200 0.1 (sm_100a)"

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------

// one warp per sequence: ASCII -> code (in place) and range check; bad[s] = 1 if any code is outside [0, n)
__global__ void k_encode_validate(int8_t* blob, const int64_t* off, const int32_t* len, int32_t nseq, int n, int ascii, uint8_t* bad, int64_t byte_base, int shift)
{
    // bad[s]: bit0 = code outside [0, n) (invalid input), bit1 = some code >= 4 (not usable by the DPX fast path as a window)
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nseq) return;
    int8_t* s = blob + ((off[w] - byte_base) << shift);
    const int L = len[w];
    bool b = false, hi = false;
    for (int i = lane; i < L; i += 32) {
        int c = s[i];
        if (ascii) { c = swb_dna_code((unsigned char)c); s[i] = (int8_t)c; }
        if (c < 0 || c >= n) b = true;
        if (c >= 4) hi = true;
    }
    b = __any_sync(0xffffffffu, b);
    hi = __any_sync(0xffffffffu, hi);
    if (lane == 0) bad[w] = (b ? 1 : 0) | (hi ? 2 : 0);
}

// codes-only variant of k_encode_validate (nothing to rewrite): 8 lanes per sequence, aligned 16-byte loads, byte-parallel
// range tests.  bad[s] as above.
__global__ void k_validate_codes(const int8_t* __restrict__ blob, const int64_t* __restrict__ off, const int32_t* __restrict__ len, int32_t nseq, int n, uint8_t* bad, int64_t byte_base, int shift)
{
    const int sidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const unsigned gm = 0xffu << ((threadIdx.x & 31) & ~7);
    bool b = false, hi = false;
    if (sidx < nseq) {
        const int8_t* s = blob + ((off[sidx] - byte_base) << shift);
        const int L = len[sidx];
        const uintptr_t a = reinterpret_cast<uintptr_t>(s);
        const int mis = (int)(a & 15);
        const uint4* base = reinterpret_cast<const uint4*>(a - mis);
        const int nch = (mis + L + 15) >> 4;
        const uint32_t nn = (uint32_t)n * 0x01010101u;
        for (int ch = sub; ch < nch; ch += 8) {
            const uint4 v = __ldg(base + ch);
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
            const int i0 = ch * 16 - mis;                   // sequence index of this chunk's first byte
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // keep only the bytes that belong to the sequence (the others read as code 0)
                uint32_t keep = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) if ((unsigned)(i0 + 4 * q + k) < (unsigned)L) keep |= 0xffu << (8 * k);
                const uint32_t x = w[q] & keep;
                if (__vcmpgeu4(x, nn)) b = true;            // unsigned compare: negative codes are >= 128
                if (__vcmpgeu4(x, 0x04040404u)) hi = true;
            }
        }
    }
    b = (__ballot_sync(0xffffffffu, b) & gm) != 0;
    hi = (__ballot_sync(0xffffffffu, hi) & gm) != 0;
    if (sidx < nseq && sub == 0) bad[sidx] = (b ? 1 : 0) | (hi ? 2 : 0);
}

// SWB_SEQ_PACKED4 / PACKED2 -> one code per byte: packed byte i of the range becomes codes [i << shift, (i + 1) << shift)
// (shift 1: two nibbles, low first; shift 2: four 2-bit fields, bits 0-1 first).  One thread per packed byte: a warp reads 32
// consecutive bytes and writes 64 / 128 consecutive ones.
__global__ void k_unpack(const uint8_t* __restrict__ src, int8_t* __restrict__ dst, int64_t nbytes, int shift)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbytes) return;
    const uint32_t b = src[i];
    if (shift == 1) *reinterpret_cast<uint16_t*>(dst + 2 * i) = (uint16_t)((b & 15u) | ((b >> 4) << 8));
    else *reinterpret_cast<uint32_t*>(dst + 4 * i) = (b & 3u) | (((b >> 2) & 3u) << 8) | (((b >> 4) & 3u) << 16) | ((b >> 6) << 24);
}

// early download (one-shot path): the pairs re-queued by the first traceback round -- the only ones whose records still change --
// are remembered in LIST_LATE (one block per source list) ...
__global__ void k_collect_late(SwbDev d, int nxtBase, int nxtW, int warpNxt)
{
    const int b = blockIdx.x;
    const int src = b < SWB_NBANDCLASS ? nxtBase + b : b < SWB_NBANDCLASS + SWB_BANDW_MAX ? nxtW + (b - SWB_NBANDCLASS) : warpNxt;
    const int n = d.counters[src];
    __shared__ int base;
    if (threadIdx.x == 0) base = n > 0 ? atomicAdd(d.counters + LIST_LATE, n) : 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) d.list[LIST_LATE][base + i] = d.list[src][i];
}
// ... and their final records are gathered into a compact array once the step is complete
__global__ void k_gather_late(const swb_result* __restrict__ res, const int32_t* __restrict__ late, int n, swb_result* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = res[late[i]];
}

__global__ void k_prepare(SwbDev d, const uint8_t* read_bad, const uint8_t* win_bad, int32_t p0, int32_t p1)
{
    const int p = p0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    swb_result r;
    r.score1 = 0; r.score2 = 0; r.ref_begin1 = -1; r.ref_end1 = 0; r.read_begin1 = -1; r.read_end1 = 0;
    r.ref_end2 = 0; r.cigar_len = 0; r.flag = 0; r.status = SWB_OK; r.cigar_off = 0;
    const int riAbs = d.pair_read[p], wiAbs = d.pair_win[p];
    const int ri = riAbs - d.ridx_base, wi = wiAbs - d.widx_base;      // position inside the uploaded table slices
    bool ok = riAbs >= 0 && riAbs < d.n_reads_total && wiAbs >= 0 && wiAbs < d.n_windows_total &&
              ri >= 0 && ri < d.n_reads && wi >= 0 && wi < d.n_windows;
    int rl = 0, wl = 0, rb = 0;
    if (ok) {
        rl = d.read_len[ri];
        rb = d.ref_beg ? d.ref_beg[p] : 0;
        wl = d.ref_len ? d.ref_len[p] : d.win_len[wi] - rb;
        ok = rl > 0 && wl > 0 && rb >= 0 && (long long)rb + wl <= d.win_len[wi] && !(read_bad[ri] & 1) && !(win_bad[wi] & 1);
    }
    if (ok && d.score_size != 0 && d.score_size != 1 && d.score_size != 2) ok = false;   // ssw.c:856-859: no profile
    d.p_mode[p] = 0; d.p_state[p] = 0;
    d.t_bw[p] = 0; d.t_best[p] = 0;
    if (!ok) {
        r.status = SWB_ERR_BAD_INPUT;
        d.p_roff[p] = 0; d.p_woff[p] = 0; d.p_rlen[p] = 0; d.p_wlen[p] = 0; d.p_mask[p] = 0;
        d.res[p] = r;
        return;
    }
    d.p_roff[p] = (d.read_off[ri] - d.rbyte_base) << d.seq_shift;
    d.p_woff[p] = ((d.win_off[wi] - d.wbyte_base) << d.seq_shift) + rb;
    d.p_rlen[p] = rl;
    d.p_wlen[p] = wl;
    d.p_mask[p] = d.mask_len ? d.mask_len[p] : (rl / 2 < 15 ? 15 : rl / 2);      // sswpy.pyx:209-211
    d.res[p] = r;
    // ---- route: DPX fast path when the striped result is provably plain Gotoh, else exact emulation
    const int go = d.gap_open[p], ge = d.gap_ext[p];
    const int lp16 = (rl + 15) & ~15;
    const bool fast = d.fast_ok && go > ge && !(win_bad[wi] & 2) && lp16 <= 32 * SWB_NBUCKETS && wl <= d.fast_max_cols &&
                      (long long)d.max_score * rl <= FAST_MAX_SCORE;
    if (fast) {
        // which padding the result will be reported in: 16-bit semantics if the 8-bit pass can overflow at all
        const bool wordSem = d.score_size == 1 || (d.max_score * rl + d.bias >= 255);
        d.p_mode[p] = wordSem ? 1 : 0;
        const int b = (lp16 + 31) / 32 - 1;
        // 8-bit-final pairs whose scores can pass 128+go+ge (where the 8-bit pass may deviate from Gotoh) take the sandwich sweep
        const bool sw = !wordSem && !(d.opt & 128) && d.max_score * rl >= 128 + go + ge;
        // forward family: the 8-thread-group sweep (rows per thread = padded length / 8: no pad rows for 150 bp) when the matrix fits its
        // scale and the read its buckets, else the 16-thread-group sweep of the read's 32-row bucket
        const int b8 = (d.fast8_ok && !sw && d.max_score * rl <= 767) ? f8_bucket(wordSem ? ((rl + 7) & ~7) : lp16) : -1;
        int fam = b;
        if (sw) list_push(d.list[LIST_SW_FWD + b], d.counters + CNT_SW_FWD + b, p);
        else if (b8 >= 0) { fam = SWB_NBUCKETS + b8; list_push(d.list[LIST_F8_FWD + b8], d.counters + CNT_F8_FWD + b8, p); }
        else list_push(d.list[LIST_FAST_FWD + b], d.counters + CNT_FAST_FWD + b, p);
        if (wl > *(volatile int32_t*)(d.counters + CNT_FAST_MAXCOLS + fam)) atomicMax(d.counters + CNT_FAST_MAXCOLS + fam, wl);   // test first: one hot address
        // the pair may continue on its 32-row bucket's lists (reverse sweep, sandwich re-sweep): that bucket's launches size their columns by it too
        if (fam != b && wl > *(volatile int32_t*)(d.counters + CNT_FAST_MAXCOLS + b)) atomicMax(d.counters + CNT_FAST_MAXCOLS + b, wl);
    }
    else if (d.score_size == 1) list_push(d.list[LIST_WORD_FWD], d.counters + CNT_WORD_FWD, p);
    else list_push(d.list[LIST_BYTE_FWD], d.counters + CNT_BYTE_FWD, p);
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------

const bool g_trace = getenv("SWB200_TRACE") != nullptr;
static std::string g_create_err;
static std::mutex g_mu;
static void destroy_ctx(swb_ctx* c);


extern "C" int swb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" const char* swb_version(void) { return SWB_VERSION; }

extern "C" swb_ctx* swb_create(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_err = std::string("no CUDA device usable: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + " (libswb200 has no CPU fallback)";
        cudaGetLastError();
        return nullptr;
    }
    if (device < 0 || device >= n) { g_create_err = "device index out of range"; return nullptr; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return nullptr; }
    if (prop.major != 10) {
        g_create_err = std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "; libswb200 is built for sm_100a only";
        return nullptr;
    }
    swb_ctx* c = new swb_ctx();
    c->device = device;
    c->n_sm = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    memset(&c->d, 0, sizeof c->d);
    c->d.warp_next = -1;
    memset(&c->tm, 0, sizeof c->tm);
    // stream priorities: the long forward sweep runs at the lowest priority, the per-class side streams in the middle and
    // everything on the critical path (small launches between host round trips) at the highest, so that with two pipeline
    // lanes one lane's short kernels are not queued behind the other lane's thousands of sweep blocks
    int prevDevice = -1;
    cudaGetDevice(&prevDevice);
    int prLeast = 0, prGreatest = 0;
    if (cudaSetDevice(device) != cudaSuccess) {
        g_create_err = std::string("cannot select device: ") + cudaGetErrorString(cudaGetLastError());
        delete c; return nullptr;
    }
    cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest);
    const int prMid = (prLeast + prGreatest) / 2;
    // every creation / allocation is checked: a context with a missing stream or a null pinned mirror must not be handed out
    cudaError_t bad = cudaSuccess; const char* what = "";
    auto chk = [&](cudaError_t e2, const char* w) { if (e2 != cudaSuccess && bad == cudaSuccess) { bad = e2; what = w; } };
    auto mkStream = [&](cudaStream_t* st, int prio) { chk(cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, prio), "cudaStreamCreateWithPriority"); };
    auto mkEvent = [&](cudaEvent_t* ev) { chk(cudaEventCreateWithFlags(ev, cudaEventDisableTiming), "cudaEventCreateWithFlags"); };
    mkStream(&c->stream, prGreatest);
    for (int i = 0; i < EV_COUNT; ++i) chk(cudaEventCreate(&c->ev[i]), "cudaEventCreate");
    mkStream(&c->stream2, prGreatest);
    mkStream(&c->stream3, prGreatest);
    mkStream(&c->stream4, prGreatest); mkEvent(&c->ev_join3);
    mkStream(&c->bulk_stream, prLeast);
    mkStream(&c->bulk_stream2, prLeast);
    mkStream(&c->copy_stream, prGreatest); mkEvent(&c->ev_copy);
    mkEvent(&c->ev_bulk_join2); mkEvent(&c->ev_piece);
    mkEvent(&c->ev_bulk_fork); mkEvent(&c->ev_bulk_join);
    mkEvent(&c->ev_fork3); mkEvent(&c->ev_x_fork); mkEvent(&c->ev_x_join);
    mkEvent(&c->ev_rev_fork);
    for (int i = 0; i < SWB_BANDW_MAX; ++i) { mkStream(&c->bandw_stream[i], prMid); mkEvent(&c->ev_bandw_join[i]); }
    for (int i = 0; i < SWB_NREVB; ++i) { mkStream(&c->rev_stream[i], prMid); mkEvent(&c->ev_rev_join[i]); }
    for (int i = 0; i < SWB_NSIDE; ++i) { mkStream(&c->side_stream[i], prMid); mkEvent(&c->ev_side_join[i]); }
    mkEvent(&c->ev_side_split); mkEvent(&c->ev_early);
    for (int i = 0; i < SWB_MAX_PARTS; ++i) mkEvent(&c->ev_part_fwd[i]);
    chk(cudaEventCreate(&c->ev_fwd_end), "cudaEventCreate");
    mkEvent(&c->ev_fork); mkEvent(&c->ev_join); mkEvent(&c->ev_join2);
    chk(cudaMallocHost((void**)&c->h_counters, SWB_NCOUNTERS * sizeof(int32_t)), "cudaMallocHost");
    for (int i = 0; i < 2; ++i) { chk(cudaMallocHost((void**)&c->h_snap[i], SWB_NCOUNTERS * sizeof(int32_t)), "cudaMallocHost"); mkEvent(&c->ev_snap[i]); }
    chk(cudaMallocHost((void**)&c->h_bump, 2 * sizeof(unsigned long long)), "cudaMallocHost");
    // opt in to large dynamic shared memory for the exact kernels
    chk(swb_exact_set_attrs(c->smem_optin), "cudaFuncSetAttribute");
    chk(swb_band_set_attrs(), "cudaFuncSetAttribute");
    if (prevDevice >= 0 && prevDevice != device) cudaSetDevice(prevDevice);      // leave the caller's current device as it was
    if (bad != cudaSuccess) {
        g_create_err = std::string(what) + ": " + cudaGetErrorString(bad);
        cudaGetLastError();
        destroy_ctx(c);
        return nullptr;
    }
    return c;
}

static void destroy_ctx(swb_ctx* c) {
    if (!c) return;
    if (c->sibling) { destroy_ctx(c->sibling); c->sibling = nullptr; }
    int prevDevice = -1;
    cudaGetDevice(&prevDevice);
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    DevBuf* all[] = { &c->b_reads, &c->b_read_off, &c->b_read_len, &c->b_windows, &c->b_win_off, &c->b_win_len, &c->b_pair_read, &c->b_pair_win,
                      &c->b_ref_beg, &c->b_ref_len, &c->b_go, &c->b_ge, &c->b_mask, &c->b_mat, &c->b_roff, &c->b_woff, &c->b_rlen, &c->b_wlen,
                      &c->b_pmask, &c->b_mode, &c->b_res, &c->b_lists, &c->b_counters, &c->b_colmax, &c->b_band, &c->b_cigar, &c->b_bump,
                      &c->b_tbw, &c->b_tbest, &c->b_rbad, &c->b_wbad, &c->b_state, &c->b_csafe, &c->b_fastcols,
                      &c->b_ind_off, &c->b_ind_cnt, &c->b_ind_rend, &c->b_ind_recs, &c->b_ind_misc, &c->b_ind_cig, &c->b_ind_coff, &c->b_ind_clen, &c->b_ind_rs, &c->b_ind_qs, &c->b_reads_pk, &c->b_windows_pk, &c->b_late };
    for (DevBuf* b : all) b->release();
    auto dS = [](cudaStream_t st) { if (st) cudaStreamDestroy(st); };
    auto dE = [](cudaEvent_t ev) { if (ev) cudaEventDestroy(ev); };
    for (int i = 0; i < EV_COUNT; ++i) dE(c->ev[i]);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    if (c->h_bump) cudaFreeHost(c->h_bump);
    for (int i = 0; i < 2; ++i) { if (c->h_snap[i]) cudaFreeHost(c->h_snap[i]); dE(c->ev_snap[i]); }
    dS(c->stream);
    dS(c->stream3); dE(c->ev_fork3); dE(c->ev_x_fork); dE(c->ev_x_join);
    dS(c->stream4); dE(c->ev_join3);
    dS(c->bulk_stream); dE(c->ev_bulk_fork); dE(c->ev_bulk_join);
    dS(c->bulk_stream2); dE(c->ev_bulk_join2); dE(c->ev_piece);
    dS(c->copy_stream); dE(c->ev_copy);
    dE(c->ev_rev_fork);
    for (int i = 0; i < SWB_NREVB; ++i) { dS(c->rev_stream[i]); dE(c->ev_rev_join[i]); }
    for (int i = 0; i < SWB_NSIDE; ++i) { dS(c->side_stream[i]); dE(c->ev_side_join[i]); }
    dE(c->ev_side_split); dE(c->ev_early);
    if (c->h_late_rec) cudaFreeHost(c->h_late_rec);
    if (c->h_late_idx) cudaFreeHost(c->h_late_idx);
    for (int i = 0; i < SWB_MAX_PARTS; ++i) dE(c->ev_part_fwd[i]);
    dE(c->ev_fwd_end);
    for (int i = 0; i < SWB_BANDW_MAX; ++i) { dS(c->bandw_stream[i]); dE(c->ev_bandw_join[i]); }
    dS(c->stream2); dE(c->ev_fork); dE(c->ev_join); dE(c->ev_join2);
    cudaGetLastError();
    if (prevDevice >= 0 && prevDevice != c->device) cudaSetDevice(prevDevice);
    delete c;
}

extern "C" void swb_destroy(swb_ctx* c) {
    std::lock_guard<std::mutex> lk(g_mu);
    destroy_ctx(c);
}

extern "C" const char* swb_last_error(const swb_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

extern "C" void* swb_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, (size_t)(bytes > 0 ? bytes : 1)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void swb_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int64_t swb_pack_table(const int8_t* blob, const int64_t* off, const int32_t* len, int32_t n, int src_ascii, int bits, uint8_t* dst, int64_t* dst_off) {
    if ((bits != 2 && bits != 4) || n < 0 || (n > 0 && (!blob || !off || !len || !dst || !dst_off))) return -1;
    const int per = 8 / bits;
    const unsigned lim = bits == 2 ? 4u : 16u;
    int64_t pos = 0;
    for (int32_t i = 0; i < n; ++i) {
        const int8_t* s = blob + off[i];
        const int32_t l = len[i];
        if (l < 0 || off[i] < 0) return -1;
        dst_off[i] = pos;
        for (int32_t k = 0; k < l; k += per) {
            unsigned byte = 0;
            for (int q = 0; q < per && k + q < l; ++q) {
                const unsigned code = src_ascii ? (unsigned)swb_dna_code((unsigned char)s[k + q]) : (unsigned)(unsigned char)s[k + q];
                if (code >= lim) return -1;
                byte |= code << (bits * q);
            }
            dst[pos++] = (uint8_t)byte;
        }
    }
    return pos;
}

extern "C" void swb_encode_dna(const char* ascii, int8_t* codes, int64_t len) {
    for (int64_t i = 0; i < len; ++i) codes[i] = swb_dna_code((unsigned char)ascii[i]);
}

// ------------------------------------------------------------------------------------------------
// upload
// ------------------------------------------------------------------------------------------------

template <typename T>
static int up(swb_ctx* c, DevBuf& b, const T* src, size_t count, T** dst) {
    if (!src) { *dst = nullptr; return 0; }
    CUDA_TRY(c, b.ensure(count * sizeof(T) + 16));
    if (count) CUDA_TRY(c, cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    c->tm.h2d_bytes += (int64_t)(count * sizeof(T));
    *dst = reinterpret_cast<T*>(b.p);
    return 0;
}

// chunk view of a batch: `b` is already sliced (table pointers/counts and per-pair pointers offset); the bases tell
// the kernels how to translate the caller's indices / byte offsets into the slice
struct ChunkView { int32_t ridx_base = 0, widx_base = 0, n_reads_total = 0, n_windows_total = 0; int64_t rbyte_base = 0, wbyte_base = 0; bool sliced = false; };

static int upload_view(swb_ctx* c, const swb_batch* b, const ChunkView& v);

extern "C" int swb_upload(swb_ctx* c, const swb_batch* b) {
    if (!c) return -1;
    c->pipelined_last = false;
    ChunkView v;
    if (b) { v.n_reads_total = b->n_reads; v.n_windows_total = b->n_windows; }
    return upload_view(c, b, v);
}

static inline int seq_shift_of(int enc) { return enc == SWB_SEQ_PACKED4 ? 1 : enc == SWB_SEQ_PACKED2 ? 2 : 0; }
// bytes a sequence of len bases occupies in the caller's blob
static inline int64_t seq_bytes(int32_t len, int shift) { return ((int64_t)len + (1 << shift) - 1) >> shift; }
static int launch_unpack(swb_ctx* c, const uint8_t* src, int8_t* dst, int64_t nbytes, int shift, cudaStream_t st) {
    if (nbytes <= 0) return 0;
    k_unpack<<<(unsigned)((nbytes + 255) / 256), 256, 0, st>>>(src, dst, nbytes, shift);
    c->tm.n_launches++;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

// batch-level scalars of SwbDev (everything but the device pointers)
static void set_batch_scalars(swb_ctx* c, const swb_batch* b, const ChunkView& v, int32_t max_rl, int32_t max_wl) {
    SwbDev& d = c->d;
    d.n_pairs = b->n_pairs; d.n_reads = b->n_reads; d.n_windows = b->n_windows;
    d.ridx_base = v.ridx_base; d.widx_base = v.widx_base; d.n_reads_total = v.n_reads_total; d.n_windows_total = v.n_windows_total;
    d.rbyte_base = v.rbyte_base; d.wbyte_base = v.wbyte_base;
    d.n = b->n; d.score_size = b->score_size; d.flag = b->flag; d.filters = b->filters; d.filterd = b->filterd;
    d.seq_encoding = b->seq_encoding;
    d.seq_shift = seq_shift_of(b->seq_encoding);
    d.max_rlen = max_rl; d.max_wlen = max_wl;
    int bias = 0;
    for (int i = 0; i < b->n * b->n; ++i) if (b->mat[i] < bias) bias = b->mat[i];       // ssw.c:795-797
    d.bias = (b->score_size == 0 || b->score_size == 2) ? std::abs(bias) : 0;
    int mx = 0; bool small = true;
    for (int i = 0; i < b->n * b->n; ++i) { mx = std::max<int>(mx, b->mat[i]); if (b->mat[i] > 7 || b->mat[i] < -7) small = false; }
    d.max_score = mx;
    { const char* o = getenv("SWB200_OPT"); d.opt = o ? atoi(o) : 0; }
    // one traceback phase.  (The two-phase split -- pairs that can fail the overflow certificate traced back first, so that their
    // exact 8-bit verification hides behind the bulk of the traceback -- paid while that verification took ~3 ms; the sandwich
    // lower bound settles it in ~0.1 ms now and the second phase only costs a second low-occupancy tail: 4.35 -> 3.79 ms on
    // config 2.  SWB200_TWO_PHASE=1 brings it back.)
    if (!getenv("SWB200_TWO_PHASE")) d.opt |= 32;
    d.one = 1;
    { bool f8 = true; for (int i = 0; i < b->n * b->n; ++i) if (b->mat[i] > 3 || b->mat[i] < -4) f8 = false;
      d.fast8_ok = (f8 && !getenv("SWB200_NO_G8")) ? 1 : 0; }
    d.fast_ok = (small && b->n >= 4 && mx > 0 && (b->score_size == 1 || b->score_size == 2) && !getenv("SWB200_NO_FAST")) ? 1 : 0;
}

static int upload_view(swb_ctx* c, const swb_batch* b, const ChunkView& v) {
    if (!c) return -1;
    c->err.clear();
    c->have_batch = false; c->computed = false;
    if (!b || b->n_pairs < 0 || b->n_reads < 0 || b->n_windows < 0) { c->err = "bad batch header"; return -1; }
    if (b->n < 1 || b->n > SWB_MAX_N || !b->mat) { c->err = "substitution matrix edge n must be in [1, 32]"; return -1; }
    if (b->n_pairs && (!b->pair_read || !b->pair_win || !b->gap_open || !b->gap_ext)) { c->err = "missing per-pair arrays"; return -1; }
    if ((b->n_reads && (!b->reads || !b->read_off || !b->read_len)) || (b->n_windows && (!b->windows || !b->win_off || !b->win_len))) { c->err = "missing sequence tables"; return -1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    SwbDev& d = c->d;
    memset(&c->tm, 0, sizeof c->tm);

    // table extents (host scan of the small length arrays; also the max lengths that size shared memory)
    int64_t reads_bytes = 0, win_bytes = 0; int32_t max_rl = 0, max_wl = 0;
    // ASCII tables are encoded IN PLACE on the device and DNA_BASE_LUT is not idempotent ('A' -> 0 -> 4): entries that share
    // blob bytes would be encoded twice.  ASCII entries must therefore be ascending and disjoint (include/swb200.h).
    const bool ascii_in = b->seq_encoding == SWB_SEQ_ASCII;
    if (b->seq_encoding < SWB_SEQ_CODES || b->seq_encoding > SWB_SEQ_PACKED2) { c->err = "unknown seq_encoding"; return -1; }
    const int shift = seq_shift_of(b->seq_encoding);
    for (int32_t i = 0; i < b->n_reads; ++i) {
        if (b->read_len[i] < 0 || b->read_off[i] < v.rbyte_base) { c->err = "negative read offset/length"; return -1; }
        if (ascii_in && b->read_off[i] - v.rbyte_base < reads_bytes) { c->err = "ASCII read table entries must be ascending and must not overlap"; return -1; }
        reads_bytes = std::max<int64_t>(reads_bytes, b->read_off[i] - v.rbyte_base + seq_bytes(b->read_len[i], shift)); max_rl = std::max(max_rl, b->read_len[i]);
    }
    for (int32_t i = 0; i < b->n_windows; ++i) {
        if (b->win_len[i] < 0 || b->win_off[i] < v.wbyte_base) { c->err = "negative window offset/length"; return -1; }
        if (ascii_in && b->win_off[i] - v.wbyte_base < win_bytes) { c->err = "ASCII window table entries must be ascending and must not overlap"; return -1; }
        win_bytes = std::max<int64_t>(win_bytes, b->win_off[i] - v.wbyte_base + seq_bytes(b->win_len[i], shift)); max_wl = std::max(max_wl, b->win_len[i]);
    }

    CUDA_TRY(c, cudaEventRecord(c->ev[EV_H2D0], c->stream));
    const size_t np = (size_t)b->n_pairs;
    if (shift) {
        // packed input: the bytes land in a staging buffer and are unpacked on the device into the one-code-per-byte blobs
        int8_t *pr = nullptr, *pw = nullptr;
        if (up(c, c->b_reads_pk, b->reads, (size_t)reads_bytes, &pr)) return -1;
        if (up(c, c->b_windows_pk, b->windows, (size_t)win_bytes, &pw)) return -1;
        CUDA_TRY(c, c->b_reads.ensure(((size_t)reads_bytes << shift) + 16));   d.reads = (int8_t*)c->b_reads.p;
        CUDA_TRY(c, c->b_windows.ensure(((size_t)win_bytes << shift) + 16));   d.windows = (int8_t*)c->b_windows.p;
        if (launch_unpack(c, (const uint8_t*)pr, d.reads, reads_bytes, shift, c->stream)) return -1;
        if (launch_unpack(c, (const uint8_t*)pw, d.windows, win_bytes, shift, c->stream)) return -1;
    } else {
    if (up(c, c->b_reads, b->reads, (size_t)reads_bytes, &d.reads)) return -1;
    if (up(c, c->b_windows, b->windows, (size_t)win_bytes, &d.windows)) return -1;
    }
    if (up(c, c->b_read_off, b->read_off, (size_t)b->n_reads, &d.read_off)) return -1;
    if (up(c, c->b_read_len, b->read_len, (size_t)b->n_reads, &d.read_len)) return -1;
    if (up(c, c->b_win_off, b->win_off, (size_t)b->n_windows, &d.win_off)) return -1;
    if (up(c, c->b_win_len, b->win_len, (size_t)b->n_windows, &d.win_len)) return -1;
    if (up(c, c->b_pair_read, b->pair_read, np, &d.pair_read)) return -1;
    if (up(c, c->b_pair_win, b->pair_win, np, &d.pair_win)) return -1;
    if (up(c, c->b_ref_beg, b->ref_beg, np, &d.ref_beg)) return -1;
    if (up(c, c->b_ref_len, b->ref_len, np, &d.ref_len)) return -1;
    if (up(c, c->b_go, b->gap_open, np, &d.gap_open)) return -1;
    if (up(c, c->b_ge, b->gap_ext, np, &d.gap_ext)) return -1;
    if (up(c, c->b_mask, b->mask_len, np, &d.mask_len)) return -1;
    if (up(c, c->b_mat, b->mat, (size_t)b->n * b->n, &d.mat)) return -1;
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_H2D1], c->stream));

    set_batch_scalars(c, b, v, max_rl, max_wl);
    c->have_batch = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// compute
// ------------------------------------------------------------------------------------------------

static int read_counters(swb_ctx* c) {
    CUDA_TRY(c, cudaMemcpyAsync(c->h_counters, c->d.counters, SWB_NCOUNTERS * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_bump, c->d.bump, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    TR(c, "counters");
    return 0;
}

// sequence table -> codes in place (ASCII input) + per-sequence validity flags
static int launch_validate(swb_ctx* c, int8_t* blob, const int64_t* off, const int32_t* len, int32_t nseq, int ascii, uint8_t* bad, int64_t byte_base, cudaStream_t st) {
    if (nseq <= 0) return 0;
    if (ascii) k_encode_validate<<<(nseq + 3) / 4, 128, 0, st>>>(blob, off, len, nseq, c->d.n, 1, bad, byte_base, c->d.seq_shift);
    else k_validate_codes<<<(nseq + 15) / 16, 128, 0, st>>>(blob, off, len, nseq, c->d.n, bad, byte_base, c->d.seq_shift);
    c->tm.n_launches++;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

// counts: plain forward list lengths per family (SWB_NFWD entries: 16-thread-group buckets, then 8-thread-group buckets);
// c->swCounts: what k_prepare put on the sandwich lists.  The plain forward sweeps append to a bucket's sandwich list (16-bit-semantics
// pairs whose result turned out to be the 8-bit pass's) and to its reverse list, so the follow-up launches are sized by an upper
// bound per 32-row bucket (the kernels read the real list length)
static void bucket_upper_bounds(const swb_ctx* c, const int* counts, int* ub) {
    // 8-thread-group bucket j (padded length <= 32, 56, 80, 104, 128, 152) -> the 32-row buckets its reads can belong to
    static const int map8[SWB_NF8][2] = {{0, 0}, {1, 1}, {1, 2}, {2, 3}, {3, 3}, {4, 4}};
    for (int b = 0; b < SWB_NBUCKETS; ++b) ub[b] = counts[b] + c->swCounts[b];
    for (int j = 0; j < SWB_NF8; ++j) for (int b = map8[j][0]; b <= map8[j][1]; ++b) ub[b] += counts[SWB_NBUCKETS + j];
}

// forward sweeps of the current part on the low-priority stream (see swb_create); the caller orders that stream behind the prepare
// kernels and records the event the part's tail waits for
static int launch_fast_fwd(swb_ctx* c, const int* counts) {
    int ub[SWB_NBUCKETS];
    bucket_upper_bounds(c, counts, ub);
    int rc = swb_launch_fast8_range_fwd(c, nullptr, counts, c->bulk_stream);
    if (!rc) rc = swb_launch_fast_range_fwd(c, nullptr, counts, c->bulk_stream);
    if (!rc && !(c->d.opt & 128)) rc = swb_launch_sandwich_fwd(c, ub, c->bulk_stream);
    if (rc) return rc;
    return stage_check(c, "fast fwd");
}
// reverse sweeps (wavefront) of the current part on the main stream
static int launch_fast_rev(swb_ctx* c, const int* counts) {
    int ub[SWB_NBUCKETS];
    bucket_upper_bounds(c, counts, ub);
    int rc = swb_launch_fast_range_rev(c, nullptr, ub, c->stream);
    if (!rc && !(c->d.opt & 128)) rc = swb_launch_sandwich_rev(c, ub, c->stream);
    if (rc) return rc;
    return stage_check(c, "fast rev");
}

// the main stream waits for the band classes (after it has queued the wavefront sweep of the remaining pairs)
static int join_rev_band(swb_ctx* c, int upperBoundPairs) {
    if (upperBoundPairs <= 0 || (c->d.opt & 4)) return 0;
    for (int i = 0; i < SWB_NREVB; ++i) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_rev_join[i], 0));
    return stage_check(c, "rev band");
}

// certificate pass over the pairs whose traceback is done, then the exact 8-bit verification of those that failed it,
// on the verification stream (no host round trip: the grid is sized by an upper bound, the kernel reads the real count)
static int certify_and_verify_async(swb_ctx* c, int verifyList, int upperBound, int which = 0) {
    SwbDev& d = c->d;
    cudaStream_t s = c->stream;
    cudaStream_t vs = which ? c->stream4 : c->stream3;       // the two verifications of a compute run side by side (each is a long, serial kernel over a handful of pairs)
    const Part& pt = c->parts[c->cur_part];
    const size_t np = (size_t)std::max<int32_t>(pt.p1 - pt.p0, 1);
    k_certify_rest<<<(unsigned)((np + 127) / 128), 128, 0, s>>>(d, pt.p0, pt.p1, verifyList);
    c->tm.n_launches++;
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(c->ev_fork3, s));
    CUDA_TRY(c, cudaStreamWaitEvent(vs, c->ev_fork3, 0));
    // first the sandwich lower bound (a DPX sweep, microseconds for a handful of pairs); what it cannot settle goes on to the exact
    // 8-bit pass, which confirms the overflow or produces the byte-mode result
    int xList = verifyList;
    if (!(d.opt & 128)) {
        const int xs = LIST_VERIFYX + which;
        CUDA_TRY(c, cudaMemsetAsync(d.counters + xs, 0, 4, vs));
        const int rc = swb_launch_sandwich_verify(c, verifyList, xs, upperBound, vs);
        if (rc < 0) return -1;
        if (rc == 0) xList = xs;
    }
    if (swb_launch_exact(c, 0, 0, xList, upperBound, vs, /*fewJobsLikely=*/true)) return -1;
    CUDA_TRY(c, cudaEventRecord(which ? c->ev_join3 : c->ev_join2, vs));
    c->verify_pending |= 1u << which;
    return 0;
}
static int certify_phase2_hook(swb_ctx* c, int total) { return total > 0 ? certify_and_verify_async(c, LIST_VERIFY2, total, 1) : 0; }

static int launch_band_reg(swb_ctx* c, int baseW, const int* njobsW, int nextBase, int nextBaseW, int resume) {
    int any = 0;
    for (int w = 1; w <= SWB_BANDW_MAX; ++w) any += njobsW[w - 1];
    if (!any) return 0;
    CUDA_TRY(c, cudaEventRecord(c->ev_rev_fork, c->stream));
    c->bandreg_used = 0;
    for (int w = SWB_BANDW_MAX; w >= 1; --w) {              // widest (longest threads) first
        const int n = njobsW[w - 1];
        if (n <= 0) continue;
        cudaStream_t st = c->bandw_stream[w - 1];
        CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_rev_fork, 0));
        if ((w <= 12 ? swb_launch_band_reg_lo : swb_launch_band_reg_hi)(c, w, baseW + w - 1, n, nextBase, nextBaseW, resume, st)) return -1;
        CUDA_TRY(c, cudaEventRecord(c->ev_bandw_join[w - 1], st));
        c->bandreg_used |= 1u << (w - 1);
    }
    c->bandreg_base = baseW;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}
// the main stream waits for the register-band kernels (after it has queued the literal kernel for the other jobs)
static int join_band_reg(swb_ctx* c) {
    if (!c->bandreg_used) return 0;
    for (int i = 0; i < SWB_BANDW_MAX; ++i) if (c->bandreg_used & (1u << i)) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_bandw_join[i], 0));
    CUDA_TRY(c, cudaMemsetAsync(c->d.counters + c->bandreg_base, 0, 4 * SWB_BANDW_MAX, c->stream));      // lists consumed
    c->bandreg_used = 0;
    return 0;
}

// banded DP + traceback (ssw.c:897-916) over the four band-class lists; a launch round per class, repeated only
// for pairs the kernel re-queued (scratch exhausted, or band outgrew the shared-memory rows)
// firstRoundOnly: launch the jobs of `firstBase` and leave what they re-queue in LIST_BAND_NEXT for a later call;
// keepNext: LIST_BAND_NEXT already holds such re-queued jobs, append to them in the first round instead of clearing
static int run_band_rounds(swb_ctx* c, bool record, int firstBase = LIST_BAND, int* firstJobs = nullptr, bool firstRoundOnly = false, bool keepNext = false,
                           int (*afterFirstRound)(swb_ctx*, int) = nullptr, bool haveCounters = false) {
    SwbDev& d = c->d;
    cudaStream_t s = c->stream;
    if (!haveCounters && read_counters(c)) return -1;       // haveCounters: the host copy already holds this phase's list sizes
    int cur = firstBase, nxt = LIST_BAND_NEXT;
    int round = 0, stalls = 0;
    if (firstJobs) *firstJobs = 0;
    for (;;) {
        int njobs[SWB_NBANDCLASS], total = 0;
        for (int k = 0; k < SWB_NBANDCLASS; ++k) { njobs[k] = c->h_counters[cur + k]; total += njobs[k]; }
        // round 0 also serves the register-band lists that belong to this phase (their re-queues land in `nxt`)
        // later rounds: the jobs a register-band kernel widened once and that are still regular (LIST_BANDW_NEXT)
        int njobsW[SWB_BANDW_MAX] = {};
        const int baseW = round == 0 ? (firstBase == LIST_BAND_FIRST ? LIST_BANDW_FIRST : LIST_BANDW) : LIST_BANDW_NEXT;
        if (round == 0 || !firstRoundOnly) for (int k = 0; k < SWB_BANDW_MAX; ++k) { njobsW[k] = c->h_counters[baseW + k]; total += njobsW[k]; }
        // round 0: the wide regular bands of this phase (one warp per alignment, swb_bandwarp.cuh)
        // later rounds: the widened jobs that fit it (LIST_BANDWARP_NEXT, two lists alternating like cur / nxt)
        const int baseWarp = round == 0 ? (firstBase == LIST_BAND_FIRST ? LIST_BANDWARP_FIRST : LIST_BANDWARP) : LIST_BANDWARP_NEXT + ((round - 1) & 1);
        const int warpNxt = LIST_BANDWARP_NEXT + (round & 1);
        const int nWarp = c->h_counters[baseWarp];
        total += nWarp;
        if (round == 0 && firstJobs) *firstJobs = total;
        if (g_trace) { fprintf(stderr, "TRACE band round %d base %d: classes", round, cur); for (int k = 0; k < SWB_NBANDCLASS; ++k) fprintf(stderr, " %d", njobs[k]); fprintf(stderr, " | reg"); for (int k = 0; k < SWB_BANDW_MAX; ++k) fprintf(stderr, " %d", njobsW[k]); fprintf(stderr, "\n"); }
        if (total <= 0 && !(round == 0 && keepNext)) break;
        if (!(round == 0 && keepNext)) {
            CUDA_TRY(c, cudaMemsetAsync(d.counters + nxt, 0, 4 * SWB_NBANDCLASS, s));
            CUDA_TRY(c, cudaMemsetAsync(d.counters + warpNxt, 0, 4, s));
        }
        d.warp_next = warpNxt;
        CUDA_TRY(c, cudaMemsetAsync(d.counters + CNT_BAND_OVERFLOW, 0, 4, s));
        CUDA_TRY(c, cudaMemsetAsync(d.bump, 0, 8, s));
        // the wider classes go to the side stream so their long, latency-bound threads overlap the bulk
        const bool side = njobs[4] > 0 || njobs[5] > 0 || njobs[6] > 0 || njobs[7] > 0 || nWarp > 0;
        if (side) {
            // every wide kernel is a chain of dependent steps as long as its longest job, whatever the job count: one side stream each,
            // so that a round costs the longest of them and not their sum
            CUDA_TRY(c, cudaEventRecord(c->ev_fork, s));
            for (int k = 0; k < SWB_NSIDE; ++k) CUDA_TRY(c, cudaStreamWaitEvent(c->side_stream[k], c->ev_fork, 0));
            if (nWarp > 0) {
                if (swb_launch_band_warp(c, baseWarp, nWarp, nxt, c->side_stream[0], c->side_stream[1], c->ev_side_split)) return -1;
                CUDA_TRY(c, cudaMemsetAsync(d.counters + baseWarp, 0, 4, c->side_stream[0]));      // list consumed (by the split kernel)
            }
            if (swb_launch_band(c, 0, (njobs[7] + SWB_BAND_THREADS - 1) / SWB_BAND_THREADS, cur, 7, 7, nxt, c->side_stream[2])) return -1;
            if (swb_launch_band(c, 4, (njobs[6] + SWB_BAND_HUGE_THREADS - 1) / SWB_BAND_HUGE_THREADS, cur, 6, 6, nxt, c->side_stream[3])) return -1;
            if (swb_launch_band(c, 3, (njobs[5] + SWB_BAND_WIDE_THREADS - 1) / SWB_BAND_WIDE_THREADS, cur, 5, 5, nxt, c->side_stream[4])) return -1;
            if (swb_launch_band(c, 2, (njobs[4] + SWB_BAND_MID_THREADS - 1) / SWB_BAND_MID_THREADS, cur, 4, 4, nxt, c->side_stream[5])) return -1;
            for (int k = 0; k < SWB_NSIDE; ++k) CUDA_TRY(c, cudaEventRecord(c->ev_side_join[k], c->side_stream[k]));
        }
        // widened jobs: the register-band kernels double once in place (up to half-width 14); what doubles beyond that but still fits a
        // register band (16 .. 24) is re-run in the next round by the kernel of the doubled width, whose latency is half the literal
        // kernel's; the literal / warp kernels take the rest and keep doubling in place
        const int regNext = round == 0 ? LIST_BANDW_NEXT : -1;
        if (launch_band_reg(c, baseW, njobsW, nxt, regNext, round == 0 ? 0 : 1)) return -1;
        int blocks = 0;
        for (int k = 0; k < SWB_BAND_CLS_MID; ++k) blocks += (njobs[k] + SWB_BAND_THREADS - 1) / SWB_BAND_THREADS;
        if (swb_launch_band(c, 1, blocks, cur, 0, SWB_BAND_CLS_MID - 1, nxt, s)) return -1;
        if (side) for (int k = 0; k < SWB_NSIDE; ++k) CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_side_join[k], 0));
        if (join_band_reg(c)) return -1;
        CUDA_TRY(c, cudaGetLastError());
        if (stage_check(c, "band")) return -1;
        if (record && round == 0) CUDA_TRY(c, cudaEventRecord(c->ev[EV_BAND_R0], s));
        TR(c, "band_round_enqueued");
        if (round == 0 && afterFirstRound && afterFirstRound(c, total)) return -1;
        if (firstRoundOnly) {
            // no host round trip: whatever this round re-queued (band outgrown, scratch full) is picked up by the next call
            CUDA_TRY(c, cudaMemsetAsync(d.counters + firstBase, 0, 4 * SWB_NBANDCLASS, s));
            c->tm.band_rounds++;
            return 0;
        }
        if (read_counters(c)) return -1;
        if (round == 0 && record && c->early.active && !c->early.started && c->nparts == 1) {
            // Early download: every record except those of the pairs just re-queued is final now (the certificate only touches
            // p_state; a verification that overturns a 16-bit result is detected at the end and falls back to a full download).
            // The records and the CIGARs emitted so far cross PCIe beside the re-queue rounds instead of behind them.
            const int64_t arenaNow = (int64_t)c->h_bump[1];
            if (arenaNow <= c->early.cap && c->h_counters[CNT_CIGAR_OVERFLOW] == 0) {
                k_collect_late<<<SWB_NBANDCLASS + SWB_BANDW_MAX + 1, 128, 0, s>>>(d, nxt, LIST_BANDW_NEXT, warpNxt);
                c->tm.n_launches++;
                CUDA_TRY(c, cudaEventRecord(c->ev_copy, s));
                CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
                CUDA_TRY(c, cudaEventRecord(c->ev[EV_D2H0], c->copy_stream));
                if (d.n_pairs) CUDA_TRY(c, cudaMemcpyAsync(c->early.results, d.res, (size_t)d.n_pairs * sizeof(swb_result), cudaMemcpyDeviceToHost, c->copy_stream));
                if (arenaNow) CUDA_TRY(c, cudaMemcpyAsync(c->early.arena, d.cigar, (size_t)arenaNow * 4, cudaMemcpyDeviceToHost, c->copy_stream));
                CUDA_TRY(c, cudaEventRecord(c->ev_early, c->copy_stream));
                c->early.started = true; c->early.arena_done = arenaNow;
            }
        }
        if (total > 0 && c->h_counters[CNT_BAND_OVERFLOW] >= total) {
            // nothing fitted: the scratch is smaller than a single band; grow it
            if (++stalls > 8 || c->b_band.cap >= ((size_t)64 << 30)) { c->err = "banded traceback scratch exhausted"; return -1; }
            size_t want = c->b_band.cap * 4;
            CUDA_TRY(c, c->b_band.ensure(want));
            d.band = (uint8_t*)c->b_band.p; d.band_cap = (int64_t)c->b_band.cap;
        }
        std::swap(cur, nxt);
        c->tm.band_rounds++;
        if (++round > 64) { c->err = "banded traceback did not converge"; return -1; }
    }
    if (record && round == 0) CUDA_TRY(c, cudaEventRecord(c->ev[EV_BAND_R0], s));
    // leave the list sets this call used empty for a later phase
    CUDA_TRY(c, cudaMemsetAsync(d.counters + firstBase, 0, 4 * SWB_NBANDCLASS, s));
    CUDA_TRY(c, cudaMemsetAsync(d.counters + LIST_BAND_NEXT, 0, 4 * SWB_NBANDCLASS, s));
    CUDA_TRY(c, cudaMemsetAsync(d.counters + LIST_BANDWARP_NEXT, 0, 8, s));
    return 0;
}

// the part of the workspace whose size depends on the longest read / window of the batch (only the stages after the forward
// sweeps use it, so the streamed path can size it once it has seen every table entry)
static int compute_setup_lens(swb_ctx* c) {
    SwbDev& d = c->d;
    const size_t np = (size_t)d.n_pairs;
    d.colmax_stride = (d.max_wlen + 7) & ~7;
    CUDA_TRY(c, c->b_colmax.ensure(np * (size_t)d.colmax_stride * 2 + 16)); d.colmax = (uint16_t*)c->b_colmax.p;
    // the fast path keeps 10 bytes per window column per lane-pair in shared memory
    d.fast_max_cols = 16384;     // selectors (2 B/column/lane pair) must fit shared memory; column bests move to global memory beyond 1024 columns
    {
        // direction-byte scratch: enough for a typical band on every pair; pairs that do not fit are
        // deferred to the next round by the kernel itself
        size_t want = std::max<size_t>((size_t)64 << 20, np * (size_t)(d.max_rlen + 8) * 12);
        want = std::min<size_t>(want, (size_t)8 << 30);
        if (c->b_band.cap < want) CUDA_TRY(c, c->b_band.ensure(want));
        d.band = (uint8_t*)c->b_band.p; d.band_cap = (int64_t)c->b_band.cap;
        size_t cw = std::max<size_t>(4096, np * 16);
        if (c->b_cigar.cap < cw * 4) CUDA_TRY(c, c->b_cigar.ensure(cw * 4));
        d.cigar = (uint32_t*)c->b_cigar.p; d.cigar_cap = (int64_t)(c->b_cigar.cap / 4);
    }

    return 0;
}

// make part k's job lists and counters the current ones (for the launches queued from now on)
static void use_part(swb_ctx* c, int k) {
    const Part& pt = c->parts[k];
    for (int i = 0; i < SWB_NLISTS; ++i) c->d.list[i] = pt.lists + (size_t)i * pt.perList;
    c->d.counters = pt.counters;
    c->cur_part = k;
}

// workspace for the batch described by c->d (grow-only buffers), counters cleared, start event recorded
static int compute_setup(swb_ctx* c) {
    CUDA_TRY(c, cudaSetDevice(c->device));
    SwbDev& d = c->d;
    const size_t np = (size_t)d.n_pairs;
    swb_timing& tm = c->tm;
    tm.ms_total = tm.ms_prepare = tm.ms_forward = tm.ms_reverse = tm.ms_traceback = 0;
    tm.cells_forward = tm.cells_reverse = tm.cells_band = 0; tm.n_fast = tm.n_exact = 0; tm.n_launches = 0;

    // workspace
    CUDA_TRY(c, c->b_roff.ensure(np * 8 + 16));   d.p_roff = (int64_t*)c->b_roff.p;
    CUDA_TRY(c, c->b_woff.ensure(np * 8 + 16));   d.p_woff = (int64_t*)c->b_woff.p;
    CUDA_TRY(c, c->b_rlen.ensure(np * 4 + 16));   d.p_rlen = (int32_t*)c->b_rlen.p;
    CUDA_TRY(c, c->b_wlen.ensure(np * 4 + 16));   d.p_wlen = (int32_t*)c->b_wlen.p;
    CUDA_TRY(c, c->b_pmask.ensure(np * 4 + 16));  d.p_mask = (int32_t*)c->b_pmask.p;
    CUDA_TRY(c, c->b_mode.ensure(np + 16));       d.p_mode = (uint8_t*)c->b_mode.p;
    CUDA_TRY(c, c->b_state.ensure(np + 16));      d.p_state = (uint8_t*)c->b_state.p;
    CUDA_TRY(c, c->b_csafe.ensure(np * 4 + 16));  d.p_csafe = (int32_t*)c->b_csafe.p;
    CUDA_TRY(c, c->b_res.ensure(np * sizeof(swb_result) + 16)); d.res = (swb_result*)c->b_res.p;
    // Parts: a large batch is cut into contiguous pair ranges with a job-list / counter set each, so that the latency-bound stages
    // (reverse, traceback) of one part run beside the forward sweep of the next instead of behind the whole sweep.  Kernels get SwbDev
    // by value, so a launch keeps the set that was current when it was queued (use_part).
    {
        const char* e = getenv("SWB200_PARTS");
        // Default ONE part.  Measured on config 2 (1 M pairs): 18.2 ms with one part, 19.6 / 21.1 / 22.3 ms with 2 / 3 / 4 -- the forward
        // sweep keeps every register of an SM busy, so the latency-bound kernels of the other part wait for block slots and then
        // share the issue ports; their stages take 8-11 ms instead of 5.7.  SWB200_PARTS=k keeps the experiment reachable.
        int want = c->force_parts ? c->force_parts : e ? atoi(e) : 1;
        c->nparts = std::max(1, std::min(want, SWB_MAX_PARTS));
        size_t perList = 0;
        for (int k = 0; k < c->nparts; ++k) {
            Part& pt = c->parts[k];
            pt.p0 = (int32_t)((np * k / c->nparts) & ~(size_t)1); pt.p1 = k + 1 == c->nparts ? (int32_t)np : (int32_t)((np * (k + 1) / c->nparts) & ~(size_t)1);
            perList = std::max(perList, (size_t)(pt.p1 - pt.p0) + 32);
        }
        CUDA_TRY(c, c->b_lists.ensure((size_t)c->nparts * SWB_NLISTS * perList * 4));
        CUDA_TRY(c, c->b_counters.ensure((size_t)c->nparts * SWB_NCOUNTERS * 4));
        for (int k = 0; k < c->nparts; ++k) {
            c->parts[k].lists = (int32_t*)c->b_lists.p + (size_t)k * SWB_NLISTS * perList;
            c->parts[k].perList = perList;
            c->parts[k].counters = (int32_t*)c->b_counters.p + (size_t)k * SWB_NCOUNTERS;
        }
        use_part(c, 0);
    }
    CUDA_TRY(c, c->b_bump.ensure(2 * 8));         d.bump = (unsigned long long*)c->b_bump.p;
    CUDA_TRY(c, c->b_tbw.ensure(np * 4 + 16));    d.t_bw = (int32_t*)c->b_tbw.p;
    CUDA_TRY(c, c->b_tbest.ensure(np * 4 + 16));  d.t_best = (int32_t*)c->b_tbest.p;
    CUDA_TRY(c, c->b_rbad.ensure((size_t)d.n_reads + 16));
    CUDA_TRY(c, c->b_wbad.ensure((size_t)d.n_windows + 16));
    if (compute_setup_lens(c)) return -1;

    cudaStream_t s = c->stream;
    TR(c, "compute_begin");
    CUDA_TRY(c, cudaMemsetAsync(c->b_counters.p, 0, (size_t)c->nparts * SWB_NCOUNTERS * 4, s));
    CUDA_TRY(c, cudaMemsetAsync(d.bump, 0, 16, s));
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_START], s));
    tm.ms_band_round0 = tm.ms_band_rest = tm.ms_certify = 0; tm.band_rounds = 0;
    tm.n_sw_certified = tm.n_sw_rejected = tm.n_sw_verified = 0;
    return 0;
}

static int swb_compute_impl(swb_ctx* c);

// everything after the fast-path forward sweeps of the CURRENT part (pairs [pt.p0, pt.p1)): exact forward passes, reverse, banded
// traceback, certificate; its stage times and statistics are added to c->tm.  fwdDone: recorded behind the part's forward sweeps.
static int compute_tail(swb_ctx* c, const int* fwdCounts, int nFastTotal, cudaEvent_t fwdDone) {
    SwbDev& d = c->d;
    const Part& pt = c->parts[c->cur_part];
    const size_t np = (size_t)(pt.p1 - pt.p0);               // upper bound of every list of this part
    swb_timing& tm = c->tm;
    cudaStream_t s = c->stream;
    if (fwdDone) CUDA_TRY(c, cudaStreamWaitEvent(s, fwdDone, 0));
    // ---- exact striped emulation, forward and reverse, on a side stream: these kernels serve a minority (pairs the fast path cannot
    //      decide) but each is a long chain of dependent steps whatever the count, so they run beside the reverse pass of the fast-path
    //      pairs instead of in front of and behind it.  Their reverse passes read LIST_BYTE_REV / LIST_WORD_REV, which only they fill;
    //      fast-path pairs whose reverse sweep gives up go to LIST_BYTE_REV2 / LIST_WORD_REV2 and are served after the join.
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_FWD], s));
    cudaStream_t xs = c->stream3;
    CUDA_TRY(c, cudaEventRecord(c->ev_x_fork, s));
    CUDA_TRY(c, cudaStreamWaitEvent(xs, c->ev_x_fork, 0));
    if (swb_launch_exact(c, 0, 0, LIST_BYTE_FWD, (int)np, xs, false)) return -1;
    if (swb_launch_exact(c, 1, 0, LIST_WORD_FWD, (int)np, xs, false)) return -1;
    if (swb_launch_exact(c, 0, 1, LIST_BYTE_REV, (int)np, xs, false)) return -1;
    if (swb_launch_exact(c, 1, 1, LIST_WORD_REV, (int)np, xs, false)) return -1;
    CUDA_TRY(c, cudaEventRecord(c->ev_x_join, xs));

    // ---- reverse (ssw.c:875-891) ----------------------------------------------------------------
    if (swb_launch_rev_band(c, nFastTotal)) return -1;          // the pairs the forward sweep put into a band class
    if (launch_fast_rev(c, fwdCounts)) return -1;           // the rest; rev bucket sizes are bounded by the fwd ones
    if (join_rev_band(c, nFastTotal)) return -1;
    CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_x_join, 0));
    // fast-path pairs handed to the exact reverse pass (rare: no column reached score1, or the sandwich rejected the reverse sweep):
    // their lists are sized on the host, one more counter read only when there are any
    if (read_counters(c)) return -1;
    {
        const int n2b = c->h_counters[LIST_BYTE_REV2], n2w = c->h_counters[LIST_WORD_REV2];
        if (n2b > 0 && swb_launch_exact(c, 0, 1, LIST_BYTE_REV2, n2b, nullptr, false)) return -1;
        if (n2w > 0 && swb_launch_exact(c, 1, 1, LIST_WORD_REV2, n2w, nullptr, false)) return -1;
        if ((n2b > 0 || n2w > 0) && read_counters(c)) return -1;
    }
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_REV], s));
    TR(c, "fwd_rev_enqueued");

    // ---- banded DP + traceback (ssw.c:897-916), overlapped with the overflow verification -----------------------
    // phase 1: the pairs that can fail the certificate (provisional 16-bit result + net insertion), then the certificate
    //          pass over them; their exact 8-bit verification is launched on the side stream (grid sized by an upper
    //          bound, the kernel reads the real count) while
    // phase 2: the bulk of the traceback runs on the main stream.
    c->verify_pending = 0;
    const bool certify = nFastTotal > 0 && d.score_size == 2;
    CUDA_TRY(c, cudaMemsetAsync(d.counters + CNT_BYTE_REV, 0, 4, s));      // consumed by the reverse stage; reused by the verification
    CUDA_TRY(c, cudaMemsetAsync(d.counters + CNT_WORD_FWD, 0, 4, s));
    CUDA_TRY(c, cudaMemsetAsync(d.counters + CNT_BYTE_FWD, 0, 4, s));      // reused as the leftovers' verify list
    int nFirst = 0;
    if (run_band_rounds(c, false, LIST_BAND_FIRST, &nFirst, /*firstRoundOnly=*/true, false, nullptr, /*haveCounters=*/true)) return -1;      // its re-queues join phase 2's rounds
    if (certify && nFirst > 0 && certify_and_verify_async(c, LIST_VERIFY, nFirst)) return -1;
    // phase 2; its first round is followed at once by the certificate + verification of its own pairs (hook), which
    // then overlap the re-queue rounds
    //          (its lists were complete when phase 1 read the counters: no host round trip between the phases)
    if (run_band_rounds(c, true, LIST_BAND, nullptr, false, /*keepNext=*/nFirst > 0, certify ? certify_phase2_hook : nullptr, /*haveCounters=*/true)) return -1;
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_BAND_ALL], s));
    if (c->verify_pending & 1) CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_join2, 0));      // both verifications done
    if (c->verify_pending & 2) CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_join3, 0));
    c->verify_pending = 0;

    // ---- leftovers: certificate for the pairs finished in re-queue rounds (rarely fails), byte-mode redo for verified
    //      pairs whose 8-bit pass did not overflow (never observed in practice)
    if (certify) {
        k_certify_rest<<<(unsigned)((np + 127) / 128), 128, 0, s>>>(d, pt.p0, pt.p1, LIST_BYTE_FWD);
        tm.n_launches++;
        CUDA_TRY(c, cudaGetLastError());
        if (stage_check(c, "certify")) return -1;
        if (read_counters(c)) return -1;
        const int nverify3 = c->h_counters[LIST_BYTE_FWD];
        if (nverify3 > 0 && swb_launch_exact(c, 0, 0, LIST_BYTE_FWD, nverify3, nullptr, /*fewJobsLikely=*/true)) return -1;
        if (nverify3 > 0 && read_counters(c)) return -1;
        const int nbyte = c->h_counters[CNT_BYTE_REV];
        if (nbyte > 0) {
            if (swb_launch_exact(c, 0, 1, LIST_BYTE_REV, nbyte, nullptr, false)) return -1;
            if (run_band_rounds(c, false, LIST_BAND)) return -1;
        }
    }
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_BAND], s));
    if (read_counters(c)) return -1;
    if (c->h_counters[CNT_CIGAR_OVERFLOW] > 0) {
        // device arena too small: grow to the exact requirement and redo (rare; the default is 16 ops/pair)
        // (sized from what this part's pairs needed so far, scaled to the whole batch; the redo repeats if that is still short)
        size_t need = (size_t)((double)c->h_bump[1] * (double)d.n_pairs / (double)std::max<int32_t>(1, pt.p1)) + (size_t)c->h_bump[1] / 4 + 1024;
        cudaStreamSynchronize(c->bulk_stream); cudaStreamSynchronize(c->bulk_stream2);      // later parts' sweeps still run
        CUDA_TRY(c, c->b_cigar.ensure(need * 4));
        const int rc = swb_compute_impl(c);                 // the whole batch again, every part, from the device-resident inputs
        return rc ? rc : 1;                                 // 1: the step is complete, the caller must not go on with its own parts
    }
    // this part's stage times (its reverse / traceback stages overlap the forward sweep of the next part, so the stages of a
    // multi-part step do not add up to ms_total) and statistics
    {
        float t = 0;
        cudaEventElapsedTime(&t, c->ev[EV_FWD], c->ev[EV_REV]); tm.ms_reverse += t;
        cudaEventElapsedTime(&t, c->ev[EV_REV], c->ev[EV_BAND]); tm.ms_traceback += t;
        cudaEventElapsedTime(&t, c->ev[EV_REV], c->ev[EV_BAND_R0]); tm.ms_band_round0 += t;
        cudaEventElapsedTime(&t, c->ev[EV_BAND_R0], c->ev[EV_BAND_ALL]); tm.ms_band_rest += t;
        cudaEventElapsedTime(&t, c->ev[EV_BAND_ALL], c->ev[EV_BAND]); tm.ms_certify += t;
        int64_t v = 0;
        memcpy(&v, c->h_counters + CNT_CELLS_FWD, 8); tm.cells_forward += v;
        memcpy(&v, c->h_counters + CNT_CELLS_REV, 8); tm.cells_reverse += v;
        memcpy(&v, c->h_counters + CNT_CELLS_BAND, 8); tm.cells_band += v;
        tm.n_fast += c->h_counters[CNT_FAST_DONE] - c->h_counters[CNT_VERIFY_BYTE];
        tm.n_exact += c->h_counters[CNT_EXACT_JOBS];
        tm.n_sw_certified += c->h_counters[CNT_SW_CERTIFIED]; tm.n_sw_rejected += c->h_counters[CNT_SW_REJECTED]; tm.n_sw_verified += c->h_counters[CNT_SW_VERIFIED];
    }
    return 0;
}

// after the last part's tail: whole-step times
static int compute_finish(swb_ctx* c) {
    swb_timing& tm = c->tm;
    cudaEventElapsedTime(&tm.ms_prepare, c->ev[EV_START], c->ev[EV_PREP]);
    cudaEventElapsedTime(&tm.ms_forward, c->ev[EV_PREP], c->ev_fwd_end);      // first forward launch .. end of the last part's sweeps
    cudaEventElapsedTime(&tm.ms_total, c->ev[EV_START], c->ev[EV_BAND]);
    c->computed = true;
    TR(c, "compute_end");
    return 0;
}

// counters of the current part (already in c->h_counters) -> forward list lengths per family, sandwich list lengths, column maxima
static int part_counts(swb_ctx* c, const int32_t* hc, int* fwdCounts) {
    int nFast = 0;
    for (int f = 0; f < SWB_NFWD; ++f) {
        fwdCounts[f] = hc[f < SWB_NBUCKETS ? CNT_FAST_FWD + f : CNT_F8_FWD + (f - SWB_NBUCKETS)];
        nFast += fwdCounts[f]; c->fastMaxCols[f] = std::max(c->fastMaxCols[f], hc[CNT_FAST_MAXCOLS + f]);
    }
    for (int b = 0; b < SWB_NBUCKETS; ++b) { c->swCounts[b] = hc[CNT_SW_FWD + b]; nFast += c->swCounts[b]; }
    return nFast;
}

static int swb_compute_impl(swb_ctx* c) {
    if (compute_setup(c)) return -1;
    SwbDev& d = c->d;
    swb_timing& tm = c->tm;
    cudaStream_t s = c->stream;
    // ---- prepare ------------------------------------------------------------------------------
    if (launch_validate(c, d.reads, d.read_off, d.read_len, d.n_reads, d.seq_encoding == SWB_SEQ_ASCII, (uint8_t*)c->b_rbad.p, d.rbyte_base, s)) return -1;
    if (launch_validate(c, d.windows, d.win_off, d.win_len, d.n_windows, d.seq_encoding == SWB_SEQ_ASCII, (uint8_t*)c->b_wbad.p, d.wbyte_base, s)) return -1;
    d.seq_encoding = SWB_SEQ_CODES;                         // tables are codes from now on (repeat computes must not re-encode)
    for (int k = 0; k < c->nparts; ++k) {
        const Part& pt = c->parts[k];
        if (pt.p1 <= pt.p0) continue;
        use_part(c, k);
        k_prepare<<<(unsigned)((pt.p1 - pt.p0 + 255) / 256), 256, 0, s>>>(d, (uint8_t*)c->b_rbad.p, (uint8_t*)c->b_wbad.p, pt.p0, pt.p1);
        tm.n_launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    if (stage_check(c, "prepare")) return -1;
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_PREP], s));
    // list lengths of every part in one round trip
    std::vector<int32_t> hc((size_t)c->nparts * SWB_NCOUNTERS);
    CUDA_TRY(c, cudaMemcpyAsync(hc.data(), c->b_counters.p, hc.size() * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    for (int f = 0; f < SWB_NFWD; ++f) c->fastMaxCols[f] = 0;
    int fwdCounts[SWB_MAX_PARTS][SWB_NFWD], swc[SWB_MAX_PARTS][SWB_NBUCKETS], nFast[SWB_MAX_PARTS];
    bool global = false;
    for (int k = 0; k < c->nparts; ++k) {
        nFast[k] = part_counts(c, hc.data() + (size_t)k * SWB_NCOUNTERS, fwdCounts[k]);
        memcpy(swc[k], c->swCounts, sizeof c->swCounts);
    }
    for (int f = 0; f < SWB_NFWD; ++f) if (c->fastMaxCols[f] > fast_smem_cols(f)) global = true;
    if (global && c->nparts > 1) {
        // long windows keep their column bests in ONE global scratch: a part's reverse sweep must not run beside the next part's forward
        // sweep.  Redo as a single part (rare shape for a large batch; costs one more prepare).
        c->force_parts = 1;
        const int rc = swb_compute_impl(c);
        c->force_parts = 0;
        return rc;
    }

    // ---- forward (ssw.c:842-860), part after part on the low-priority stream; then the tails, each behind its own part's sweeps ----
    //   fast path: one 16-bit Gotoh sweep per pair (pairs it cannot decide are appended to the exact lists)
    //   exact path: 8-bit pass, then 16-bit pass for the pairs that overflowed
    CUDA_TRY(c, cudaEventRecord(c->ev_bulk_fork, s));
    CUDA_TRY(c, cudaStreamWaitEvent(c->bulk_stream, c->ev_bulk_fork, 0));
    for (int k = 0; k < c->nparts; ++k) {
        use_part(c, k);
        memcpy(c->swCounts, swc[k], sizeof c->swCounts);
        if (launch_fast_fwd(c, fwdCounts[k])) return -1;
        CUDA_TRY(c, cudaEventRecord(c->ev_part_fwd[k], c->bulk_stream));
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_fwd_end, c->bulk_stream));
    for (int k = 0; k < c->nparts; ++k) {
        use_part(c, k);
        memcpy(c->swCounts, swc[k], sizeof c->swCounts);
        if (k > 0 && c->parts[k].p1 <= c->parts[k].p0) continue;
        const int rc = compute_tail(c, fwdCounts[k], nFast[k], c->ev_part_fwd[k]);
        if (rc < 0) return rc;
        if (rc == 1) return 0;                              // redone with a larger CIGAR arena
    }
    return compute_finish(c);
}

extern "C" int swb_compute(swb_ctx* c) {
    if (!c) return -1;
    if (!c->have_batch) { c->err = "swb_compute: no batch uploaded"; return -1; }
    if (g_trace) c->trace.clear();
    const double t0 = now_ms();
    const int rc = swb_compute_impl(c);
    if (g_trace) { fprintf(stderr, "TRACE compute:"); for (auto& e : c->trace) fprintf(stderr, " %s@%.2f", e.first, e.second - t0); fprintf(stderr, "\n"); }
    return rc;
}

// ------------------------------------------------------------------------------------------------
// download
// ------------------------------------------------------------------------------------------------

extern "C" int swb_download(swb_ctx* c, swb_result* results, uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used) {
    if (!c) return -1;
    if (!c->computed) { c->err = "swb_download: nothing computed"; return -1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    const SwbDev& d = c->d;
    const int64_t used = (int64_t)c->h_bump[1];
    if (cigar_used) *cigar_used = used;
    if (used > cigar_cap || (used > 0 && !cigar_arena)) { c->err = "cigar arena too small"; return -2; }
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_D2H0], c->stream));
    if (d.n_pairs) CUDA_TRY(c, cudaMemcpyAsync(results, d.res, (size_t)d.n_pairs * sizeof(swb_result), cudaMemcpyDeviceToHost, c->stream));
    if (used) CUDA_TRY(c, cudaMemcpyAsync(cigar_arena, d.cigar, (size_t)used * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_D2H1], c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->tm.d2h_bytes = (int64_t)d.n_pairs * (int64_t)sizeof(swb_result) + used * 4;
    cudaEventElapsedTime(&c->tm.ms_d2h, c->ev[EV_D2H0], c->ev[EV_D2H1]);
    cudaEventElapsedTime(&c->tm.ms_h2d, c->ev[EV_H2D0], c->ev[EV_H2D1]);
    return 0;
}


// ------------------------------------------------------------------------------------------------
// streamed one-shot path: ONE pass over the batch, but the host->device copies are cut into pieces of
// consecutive pairs and the fast-path forward sweep of a piece is queued as soon as its sequences are on the
// device, so PCIe runs underneath the forward stage (which is longer than the whole upload) instead of in front
// of it.  Reverse, traceback and certificate then run once over the whole batch.
// ------------------------------------------------------------------------------------------------

// byte extent and longest entry of a sequence table; false on a negative offset / length
static bool scan_table_range(const int64_t* off, const int32_t* len, int32_t i0, int32_t i1, int64_t& extent, int32_t& maxlen, int shift) {
    int64_t ext = 0, minoff = 0; int32_t ml = 0, minlen = 0;
    for (int32_t i = i0; i < i1; ++i) {
        const int64_t o = off[i]; const int32_t l = len[i];
        ext = std::max<int64_t>(ext, o + seq_bytes(l, shift)); ml = std::max(ml, l); minoff = std::min(minoff, o); minlen = std::min(minlen, l);
    }
    extent = ext; maxlen = ml;
    return minoff >= 0 && minlen >= 0;
}
// both tables, split over a few host threads when they are large (two million entries take ~3 ms on one core)
static bool scan_tables(const swb_batch* b, int64_t& reads_bytes, int32_t& max_rl, int64_t& win_bytes, int32_t& max_wl) {
    const int NT = ((int64_t)b->n_reads + b->n_windows >= 400000) ? 6 : 1;      // 3 slices per table
    const int shift = seq_shift_of(b->seq_encoding);
    struct Part { int64_t ext = 0; int32_t ml = 0; bool ok = true; };
    Part pr[3], pw[3];
    auto work = [&](int k) {
        const int which = k / 3, sl = k % 3;
        const int32_t n = which ? b->n_windows : b->n_reads;
        const int32_t i0 = (int32_t)((int64_t)n * sl / 3), i1 = (int32_t)((int64_t)n * (sl + 1) / 3);
        Part& q = which ? pw[sl] : pr[sl];
        q.ok = which ? scan_table_range(b->win_off, b->win_len, i0, i1, q.ext, q.ml, shift) : scan_table_range(b->read_off, b->read_len, i0, i1, q.ext, q.ml, shift);
    };
    if (NT == 1) { for (int k = 0; k < 6; ++k) work(k); }
    else {
        std::thread th[5];
        for (int k = 1; k < 6; ++k) th[k - 1] = std::thread(work, k);
        work(0);
        for (auto& t : th) t.join();
    }
    reads_bytes = 0; win_bytes = 0; max_rl = 0; max_wl = 0; bool ok = true;
    for (int k = 0; k < 3; ++k) {
        reads_bytes = std::max(reads_bytes, pr[k].ext); max_rl = std::max(max_rl, pr[k].ml); ok = ok && pr[k].ok;
        win_bytes = std::max(win_bytes, pw[k].ext); max_wl = std::max(max_wl, pw[k].ml); ok = ok && pw[k].ok;
    }
    return ok;
}

struct TableStream {                 // upload frontier of one sequence table
    const int8_t* blob; const int64_t* off; const int32_t* len; int32_t n;
    int8_t* d_blob; int64_t* d_off; int32_t* d_len; uint8_t* d_bad;
    uint8_t* d_pk = nullptr; int shift = 0;      // packed input: the caller's bytes land in d_pk and are unpacked into d_blob (<< shift)
    int64_t cap = 0;                 // caller-side blob bytes the device buffers can hold
    int32_t maxlen = 0;              // longest entry seen so far
    int32_t front = 0;               // entries [0, front) are on the device
    int64_t ulo = 0, uhi = 0;        // blob bytes [ulo, uhi) are on the device
    bool ascii = false;              // encoded in place on the device: entries must be ascending and disjoint
    int64_t prev_end = 0;            // end of the last entry seen (ASCII rule)
};

// make entries [front, upto] resident: their table rows, the blob bytes they cover (the resident byte interval stays
// contiguous), then encode / validate them
// copies on the copy stream; the encode / validate kernel of the new entries is queued by the caller on the main stream
// (after it waits for the copies) through table_encode()
struct TableStep { int32_t i0 = 0, n = 0; };
static int table_advance(swb_ctx* c, TableStream& t, int32_t upto, TableStep& st) {
    st.i0 = t.front; st.n = 0;
    if (upto < t.front) return 0;
    const int32_t i0 = t.front, n = upto - t.front + 1;
    int64_t lo = INT64_MAX, hi = 0;
    int64_t minoff = 0; int32_t minlen = 0, ml = t.maxlen;
    bool overlap = false; int64_t pe = t.prev_end;
    for (int32_t i = i0; i <= upto; ++i) {
        const int64_t o = t.off[i]; const int32_t l = t.len[i];
        const int64_t e = o + seq_bytes(l, t.shift);
        lo = std::min<int64_t>(lo, o); hi = std::max<int64_t>(hi, e); minoff = std::min(minoff, o); minlen = std::min(minlen, l); ml = std::max(ml, l);
        overlap |= o < pe; pe = std::max<int64_t>(pe, e);
    }
    if (minoff < 0 || minlen < 0) { c->err = "negative sequence offset/length"; return -1; }
    if (t.ascii && overlap) { c->err = "ASCII sequence table entries must be ascending and must not overlap"; return -1; }
    t.prev_end = pe;
    t.maxlen = ml;
    if (hi > t.cap) return -3;                              // the blob outgrew the buffer sized from the previous call: the caller restarts with a full scan
    cudaStream_t s = c->copy_stream;
    auto copy = [&](int64_t a, int64_t b2) -> int {
        if (b2 <= a) return 0;
        if (t.shift) {
            CUDA_TRY(c, cudaMemcpyAsync(t.d_pk + a, t.blob + a, (size_t)(b2 - a), cudaMemcpyHostToDevice, s));
            if (launch_unpack(c, t.d_pk + a, t.d_blob + (a << t.shift), b2 - a, t.shift, s)) return -1;      // on the copy stream, right behind its bytes
        } else CUDA_TRY(c, cudaMemcpyAsync(t.d_blob + a, t.blob + a, (size_t)(b2 - a), cudaMemcpyHostToDevice, s));
        c->tm.h2d_bytes += b2 - a;
        return 0;
    };
    if (hi > lo) {
        if (t.uhi == t.ulo) { if (copy(lo, hi)) return -1; t.ulo = lo; t.uhi = hi; }
        else {
            if (lo < t.ulo) { if (copy(lo, t.ulo)) return -1; t.ulo = lo; }
            if (hi > t.uhi) { if (copy(t.uhi, hi)) return -1; t.uhi = hi; }
        }
    }
    CUDA_TRY(c, cudaMemcpyAsync(t.d_off + i0, t.off + i0, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(c, cudaMemcpyAsync(t.d_len + i0, t.len + i0, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    c->tm.h2d_bytes += (int64_t)n * 12;
    st.n = n;
    t.front = upto + 1;
    return 0;
}
static int table_encode(swb_ctx* c, TableStream& t, const TableStep& st, int ascii) {
    if (st.n <= 0) return 0;
    return launch_validate(c, t.d_blob, t.d_off + st.i0, t.d_len + st.i0, st.n, ascii, t.d_bad + st.i0, 0, c->stream);
}

// second half of the early download (see run_band_rounds): the records of the late pairs, compacted on the device, and the CIGARs
// emitted after the first round.  Returns 0 when the caller's arrays are complete, 1 when the full download must be used instead.
static int finish_early_download(swb_ctx* c, swb_result* results, uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used) {
    const SwbDev& d = c->d;
    cudaStream_t s = c->stream;
    const int64_t used = (int64_t)c->h_bump[1];
    const int nLate = c->h_counters[LIST_LATE];
    if (c->h_counters[CNT_VERIFY_BYTE] > 0 || nLate > d.n_pairs / 8 || used > cigar_cap || c->h_counters[CNT_CIGAR_OVERFLOW] > 0) return 1;
    if (cigar_used) *cigar_used = used;
    if (nLate > 0) {
        CUDA_TRY(c, c->b_late.ensure((size_t)nLate * sizeof(swb_result) + 16));
        if ((size_t)nLate > c->h_late_cap) {
            if (c->h_late_rec) cudaFreeHost(c->h_late_rec);
            if (c->h_late_idx) cudaFreeHost(c->h_late_idx);
            c->h_late_rec = nullptr; c->h_late_idx = nullptr; c->h_late_cap = 0;
            const size_t cap = (size_t)nLate + (size_t)nLate / 2 + 1024;
            CUDA_TRY(c, cudaMallocHost((void**)&c->h_late_rec, cap * sizeof(swb_result)));
            CUDA_TRY(c, cudaMallocHost((void**)&c->h_late_idx, cap * sizeof(int32_t)));
            c->h_late_cap = cap;
        }
        k_gather_late<<<(nLate + 255) / 256, 256, 0, s>>>(d.res, d.list[LIST_LATE], nLate, (swb_result*)c->b_late.p);
        c->tm.n_launches++;
        CUDA_TRY(c, cudaMemcpyAsync(c->h_late_rec, c->b_late.p, (size_t)nLate * sizeof(swb_result), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaMemcpyAsync(c->h_late_idx, d.list[LIST_LATE], (size_t)nLate * 4, cudaMemcpyDeviceToHost, s));
    }
    if (used > c->early.arena_done) CUDA_TRY(c, cudaMemcpyAsync(cigar_arena + c->early.arena_done, d.cigar + c->early.arena_done, (size_t)(used - c->early.arena_done) * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_D2H1], s));
    CUDA_TRY(c, cudaEventSynchronize(c->ev_early));           // the bulk of the records (copy stream)
    CUDA_TRY(c, cudaStreamSynchronize(s));
    for (int i = 0; i < nLate; ++i) results[c->h_late_idx[i]] = c->h_late_rec[i];
    c->tm.d2h_bytes = (int64_t)d.n_pairs * (int64_t)sizeof(swb_result) + used * 4 + (int64_t)nLate * (int64_t)(sizeof(swb_result) + 4);
    cudaEventElapsedTime(&c->tm.ms_h2d, c->ev[EV_H2D0], c->ev[EV_H2D1]);
    c->tm.ms_d2h = 0;                                        // overlapped with the re-queue rounds
    return 0;
}

// forceScan = false: trust the buffer capacities left by the previous call (steady state: the same kind of batch again) and
// check every piece against them; returns -3 if a blob does not fit, the caller then repeats the call with forceScan = true
static int align_batch_streamed(swb_ctx* c, const swb_batch* b, swb_result* results, uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used, bool forceScan) {
    c->err.clear();
    c->have_batch = false; c->computed = false; c->pipelined_last = false;
    if (b->n < 1 || b->n > SWB_MAX_N || !b->mat) { c->err = "substitution matrix edge n must be in [1, 32]"; return -1; }
    if (!b->pair_read || !b->pair_win || !b->gap_open || !b->gap_ext) { c->err = "missing per-pair arrays"; return -1; }
    if (!b->reads || !b->read_off || !b->read_len || !b->windows || !b->win_off || !b->win_len) { c->err = "missing sequence tables"; return -1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    SwbDev& d = c->d;
    memset(&c->tm, 0, sizeof c->tm);
    TR(c, "stream_begin");
    int64_t reads_bytes = 0, win_bytes = 0; int32_t max_rl = 0, max_wl = 0;
    if (b->seq_encoding < SWB_SEQ_CODES || b->seq_encoding > SWB_SEQ_PACKED2) { c->err = "unknown seq_encoding"; return -1; }
    const int shift = seq_shift_of(b->seq_encoding);
    const bool scan = forceScan || c->b_reads.cap < 64 || c->b_windows.cap < 64 || (shift && (c->b_reads_pk.cap < 64 || c->b_windows_pk.cap < 64));
    if (scan) {
        if (!scan_tables(b, reads_bytes, max_rl, win_bytes, max_wl)) { c->err = "negative sequence offset/length"; return -1; }
    } else {
        reads_bytes = (int64_t)c->b_reads.cap - 16; win_bytes = (int64_t)c->b_windows.cap - 16;      // what is already there; lengths follow from the pieces
        if (shift) {                                         // caller-side (packed) bytes both the staging and the unpacked buffer can take
            reads_bytes = std::min<int64_t>(reads_bytes >> shift, (int64_t)c->b_reads_pk.cap - 16);
            win_bytes = std::min<int64_t>(win_bytes >> shift, (int64_t)c->b_windows_pk.cap - 16);
        }
    }
    TR(c, "tables_scanned");
    const size_t np = (size_t)b->n_pairs, nr = (size_t)b->n_reads, nw = (size_t)b->n_windows;

    // device input buffers (filled piece by piece below)
    if (shift) { CUDA_TRY(c, c->b_reads_pk.ensure((size_t)reads_bytes + 16)); CUDA_TRY(c, c->b_windows_pk.ensure((size_t)win_bytes + 16)); }
    CUDA_TRY(c, c->b_reads.ensure(((size_t)reads_bytes << shift) + 16));   d.reads = (int8_t*)c->b_reads.p;
    CUDA_TRY(c, c->b_read_off.ensure(nr * 8 + 16));             d.read_off = (int64_t*)c->b_read_off.p;
    CUDA_TRY(c, c->b_read_len.ensure(nr * 4 + 16));             d.read_len = (int32_t*)c->b_read_len.p;
    CUDA_TRY(c, c->b_windows.ensure(((size_t)win_bytes << shift) + 16));   d.windows = (int8_t*)c->b_windows.p;
    CUDA_TRY(c, c->b_win_off.ensure(nw * 8 + 16));              d.win_off = (int64_t*)c->b_win_off.p;
    CUDA_TRY(c, c->b_win_len.ensure(nw * 4 + 16));              d.win_len = (int32_t*)c->b_win_len.p;
    CUDA_TRY(c, c->b_pair_read.ensure(np * 4 + 16));            d.pair_read = (int32_t*)c->b_pair_read.p;
    CUDA_TRY(c, c->b_pair_win.ensure(np * 4 + 16));             d.pair_win = (int32_t*)c->b_pair_win.p;
    d.ref_beg = nullptr; d.ref_len = nullptr; d.mask_len = nullptr;
    if (b->ref_beg) { CUDA_TRY(c, c->b_ref_beg.ensure(np * 4 + 16)); d.ref_beg = (int32_t*)c->b_ref_beg.p; }
    if (b->ref_len) { CUDA_TRY(c, c->b_ref_len.ensure(np * 4 + 16)); d.ref_len = (int32_t*)c->b_ref_len.p; }
    if (b->mask_len) { CUDA_TRY(c, c->b_mask.ensure(np * 4 + 16)); d.mask_len = (int32_t*)c->b_mask.p; }
    CUDA_TRY(c, c->b_go.ensure(np + 16));                       d.gap_open = (uint8_t*)c->b_go.p;
    CUDA_TRY(c, c->b_ge.ensure(np + 16));                       d.gap_ext = (uint8_t*)c->b_ge.p;
    ChunkView v; v.n_reads_total = b->n_reads; v.n_windows_total = b->n_windows;
    set_batch_scalars(c, b, v, max_rl, max_wl);
    const int ascii = b->seq_encoding == SWB_SEQ_ASCII;
    cudaStream_t s = c->stream;
    if (up(c, c->b_mat, b->mat, (size_t)b->n * b->n, &d.mat)) return -1;
    if (compute_setup(c)) return -1;
    // the copy stream starts after whatever the main stream still had queued on these buffers
    CUDA_TRY(c, cudaEventRecord(c->ev_copy, s));
    CUDA_TRY(c, cudaStreamWaitEvent(c->copy_stream, c->ev_copy, 0));
    CUDA_TRY(c, cudaEventRecord(c->ev[EV_H2D0], c->copy_stream));
    swb_timing& tm = c->tm;

    TableStream tr, tw;
    tr.blob = b->reads; tr.off = b->read_off; tr.len = b->read_len; tr.n = b->n_reads;
    tr.d_blob = d.reads; tr.d_off = d.read_off; tr.d_len = d.read_len; tr.d_bad = (uint8_t*)c->b_rbad.p;
    tw.blob = b->windows; tw.off = b->win_off; tw.len = b->win_len; tw.n = b->n_windows;
    tw.d_blob = d.windows; tw.d_off = d.win_off; tw.d_len = d.win_len; tw.d_bad = (uint8_t*)c->b_wbad.p;
    tr.cap = reads_bytes; tw.cap = win_bytes;
    tr.shift = tw.shift = shift; tr.d_pk = (uint8_t*)c->b_reads_pk.p; tw.d_pk = (uint8_t*)c->b_windows_pk.p;
    tr.ascii = tw.ascii = b->seq_encoding == SWB_SEQ_ASCII;

    // piece boundaries: small first pieces (the sweep starts early), growing by 1.3x -- slower than the ratio of the sweep's to
    // the copy's throughput, so the copies stay ahead -- up to an eighth of the batch, and shrinking again at the end: the sweep
    // of the last piece cannot start before the last byte has arrived, so that piece should be short
    std::vector<int64_t> bounds(1, 0);
    {
        const double lo = std::max<double>(16384.0, (double)np / 48.0), cap = std::max<double>(32768.0, (double)np / 8.0);
        std::vector<int64_t> ramp;
        int64_t rampSum = 0;
        for (double sz = lo; sz < cap; sz *= 1.3) { ramp.push_back((int64_t)sz & ~(int64_t)1); rampSum += ramp.back(); }
        std::vector<int64_t> sizes;
        if (2 * rampSum >= (int64_t)np) {                      // small batch: a single ramp up
            int64_t left = (int64_t)np;
            for (size_t k = 0; left > 0; ++k) { int64_t szk = k < ramp.size() ? ramp[k] : (int64_t)cap; if (left - szk < szk / 2) szk = left; sizes.push_back(szk); left -= szk; }
        } else {
            sizes = ramp;
            int64_t mid = (int64_t)np - 2 * rampSum;
            const int nmid = (int)std::max<int64_t>(1, (mid + (int64_t)cap - 1) / (int64_t)cap);
            for (int k = 0; k < nmid; ++k) { const int64_t szk = k + 1 == nmid ? mid : ((mid / (nmid - k)) & ~(int64_t)1); sizes.push_back(szk); mid -= szk; }
            for (size_t k = ramp.size(); k-- > 0; ) sizes.push_back(ramp[k]);
        }
        for (int64_t szk : sizes) if (szk > 0) bounds.push_back(std::min<int64_t>(bounds.back() + szk, (int64_t)np));
        bounds.back() = (int64_t)np;
        // a piece belongs to one part (compute_setup): cut the pieces that straddle a part boundary
        for (int k = 1; k < c->nparts; ++k) {
            const int64_t cut = c->parts[k].p0;
            auto it = std::lower_bound(bounds.begin(), bounds.end(), cut);
            if (it == bounds.end() || *it != cut) bounds.insert(it, cut);
        }
    }
    const int npieces = (int)bounds.size() - 1;
    auto part_of = [&](int k) { int q = 0; while (q + 1 < c->nparts && bounds[k] >= c->parts[q + 1].p0) ++q; return q; };

    int done[SWB_MAX_PARTS][SWB_NFWD] = {};
    int swc[SWB_MAX_PARTS][SWB_NBUCKETS] = {};
    int nFastPart[SWB_MAX_PARTS] = {};
    for (int q = 0; q < SWB_NFWD; ++q) c->fastMaxCols[q] = 0;
    bool used2 = false, serialize = false;
    // The host runs one piece ahead: piece k+1's copies and prepare kernel are queued before it waits for piece k's
    // fast-list lengths, so PCIe never idles during a host round trip.
    auto enqueue_piece = [&](int k) -> int {
        const int32_t p0 = (int32_t)bounds[k], p1 = (int32_t)bounds[k + 1];
        use_part(c, part_of(k));                             // the piece's prepare kernel and counter snapshot go to its part's set
        int32_t rmax = -1, wmax = -1;
        for (int32_t p = p0; p < p1; ++p) {
            const int32_t r = b->pair_read[p], w = b->pair_win[p];
            if ((uint32_t)r < (uint32_t)b->n_reads) rmax = std::max(rmax, r);
            if ((uint32_t)w < (uint32_t)b->n_windows) wmax = std::max(wmax, w);
        }
        TableStep sr, sw;
        { const int e1 = table_advance(c, tr, rmax, sr); if (e1) return e1; const int e2 = table_advance(c, tw, wmax, sw); if (e2) return e2; }
        const size_t n = (size_t)(p1 - p0);
        cudaStream_t cs = c->copy_stream;
        CUDA_TRY(c, cudaMemcpyAsync(d.pair_read + p0, b->pair_read + p0, n * 4, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(c, cudaMemcpyAsync(d.pair_win + p0, b->pair_win + p0, n * 4, cudaMemcpyHostToDevice, cs));
        if (b->ref_beg) CUDA_TRY(c, cudaMemcpyAsync(d.ref_beg + p0, b->ref_beg + p0, n * 4, cudaMemcpyHostToDevice, cs));
        if (b->ref_len) CUDA_TRY(c, cudaMemcpyAsync(d.ref_len + p0, b->ref_len + p0, n * 4, cudaMemcpyHostToDevice, cs));
        if (b->mask_len) CUDA_TRY(c, cudaMemcpyAsync(d.mask_len + p0, b->mask_len + p0, n * 4, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(c, cudaMemcpyAsync(d.gap_open + p0, b->gap_open + p0, n, cudaMemcpyHostToDevice, cs));
        CUDA_TRY(c, cudaMemcpyAsync(d.gap_ext + p0, b->gap_ext + p0, n, cudaMemcpyHostToDevice, cs));
        tm.h2d_bytes += (int64_t)n * (10 + (b->ref_beg ? 4 : 0) + (b->ref_len ? 4 : 0) + (b->mask_len ? 4 : 0));
        if (k + 1 == npieces) CUDA_TRY(c, cudaEventRecord(c->ev[EV_H2D1], cs));
        // the piece's kernels wait for its copies; the copy stream itself runs on to the next piece
        CUDA_TRY(c, cudaEventRecord(c->ev_copy, cs));
        CUDA_TRY(c, cudaStreamWaitEvent(s, c->ev_copy, 0));
        if (table_encode(c, tr, sr, ascii) || table_encode(c, tw, sw, ascii)) return -1;
        k_prepare<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d, (uint8_t*)c->b_rbad.p, (uint8_t*)c->b_wbad.p, p0, p1);
        tm.n_launches++;
        CUDA_TRY(c, cudaGetLastError());
        if (k + 1 == npieces) CUDA_TRY(c, cudaEventRecord(c->ev[EV_PREP], s));
        // fast-list lengths after this piece, into the snapshot buffer of its parity; the sweep streams wait on the same event
        CUDA_TRY(c, cudaMemcpyAsync(c->h_snap[k & 1], d.counters, SWB_NCOUNTERS * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaEventRecord(c->ev_snap[k & 1], s));
        return 0;
    };
    if (npieces > 0) { const int e0 = enqueue_piece(0); if (e0) return e0; }
    for (int k = 0; k < npieces; ++k) {
        const bool last = k + 1 == npieces;
        const int pk = part_of(k);
        const bool lastOfPart = last || part_of(k + 1) != pk;
        if (!last) { const int e1 = enqueue_piece(k + 1); if (e1) return e1; }
        CUDA_TRY(c, cudaEventSynchronize(c->ev_snap[k & 1]));
        TR(c, "piece_ready");
        use_part(c, pk);
        const int32_t* hc = c->h_snap[k & 1];
        int upper[SWB_NFWD]; bool any = false, global = false;
        for (int q = 0; q < SWB_NFWD; ++q) {
            const int cnt = hc[q < SWB_NBUCKETS ? CNT_FAST_FWD + q : CNT_F8_FWD + (q - SWB_NBUCKETS)];
            upper[q] = lastOfPart ? cnt : (cnt & ~1);            // lane pairs stay intact: an odd leftover waits for the part's next piece
            c->fastMaxCols[q] = std::max(c->fastMaxCols[q], hc[CNT_FAST_MAXCOLS + q]);
            if (upper[q] > done[pk][q]) any = true;
            if (c->fastMaxCols[q] > fast_smem_cols(q)) global = true;   // shared column scratch: slices must not overlap
        }
        if (global) serialize = true;                        // ... and no part's reverse sweep may run beside another part's forward sweep
        if (any) {
            const bool second = (k & 1) && !global;
            cudaStream_t st = second ? c->bulk_stream2 : c->bulk_stream;
            CUDA_TRY(c, cudaStreamWaitEvent(st, c->ev_snap[k & 1], 0));
            if (swb_launch_fast8_range_fwd(c, done[pk], upper, st)) return -1;
            if (swb_launch_fast_range_fwd(c, done[pk], upper, st)) return -1;
            if (second) used2 = true;
            for (int q = 0; q < SWB_NFWD; ++q) done[pk][q] = upper[q];
        }
        if (lastOfPart) {
            // the part's sandwich lists are swept once, after its last piece: behind every plain forward slice of the part (they
            // append to the sandwich lists) and the piece's prepare kernel
            int nSw = 0, nFwdPlain = 0;
            for (int q = 0; q < SWB_NBUCKETS; ++q) { swc[pk][q] = hc[CNT_SW_FWD + q]; nSw += swc[pk][q]; }
            for (int q = 0; q < SWB_NFWD; ++q) nFwdPlain += done[pk][q];
            nFastPart[pk] = nSw + nFwdPlain;
            memcpy(c->swCounts, swc[pk], sizeof c->swCounts);
            CUDA_TRY(c, cudaStreamWaitEvent(c->bulk_stream, c->ev_snap[k & 1], 0));
            if (used2) { CUDA_TRY(c, cudaEventRecord(c->ev_bulk_join2, c->bulk_stream2)); CUDA_TRY(c, cudaStreamWaitEvent(c->bulk_stream, c->ev_bulk_join2, 0)); }
            if ((nSw > 0 || nFwdPlain > 0) && !(d.opt & 128)) {
                int ub[SWB_NBUCKETS];
                bucket_upper_bounds(c, done[pk], ub);
                if (swb_launch_sandwich_fwd(c, ub, c->bulk_stream)) return -1;
            }
            CUDA_TRY(c, cudaEventRecord(c->ev_part_fwd[pk], c->bulk_stream));
        }
    }
    if (npieces == 0) {
        CUDA_TRY(c, cudaEventRecord(c->ev[EV_H2D1], s)); CUDA_TRY(c, cudaEventRecord(c->ev[EV_PREP], s));
        for (int k = 0; k < c->nparts; ++k) CUDA_TRY(c, cudaEventRecord(c->ev_part_fwd[k], c->bulk_stream));
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_fwd_end, c->bulk_stream));
    // only the table entries some pair refers to are resident: a later swb_compute on this context must not touch the rest
    d.n_reads = tr.front; d.n_windows = tw.front;
    if (!scan) {                                            // lengths as seen by the pieces (every entry a pair refers to)
        d.max_rlen = tr.maxlen; d.max_wlen = tw.maxlen;
        if (compute_setup_lens(c)) return -1;
    }
    d.seq_encoding = SWB_SEQ_CODES;
    c->have_batch = true;
    TR(c, "pieces_enqueued");
    c->early.active = c->nparts == 1 && !getenv("SWB200_NO_EARLY_D2H");
    c->early.started = false; c->early.results = results; c->early.arena = cigar_arena; c->early.cap = cigar_arena ? cigar_cap : 0;
    // the tails, part after part, each behind its own part's forward sweeps (behind all of them if the column scratch is shared)
    bool redone = false;
    for (int k = 0; k < c->nparts; ++k) {
        if (k > 0 && c->parts[k].p1 <= c->parts[k].p0) continue;
        use_part(c, k);
        memcpy(c->swCounts, swc[k], sizeof c->swCounts);
        const int rc = compute_tail(c, done[k], nFastPart[k], c->ev_part_fwd[serialize ? c->nparts - 1 : k]);
        if (rc < 0) return rc;
        if (rc == 1) { redone = true; break; }              // redone with a larger CIGAR arena
    }
    c->early.active = false;
    if (!redone && compute_finish(c)) return -1;
    if (c->early.started) {
        const int rc = redone ? 1 : finish_early_download(c, results, cigar_arena, cigar_cap, cigar_used);
        c->early.started = false;
        if (rc <= 0) return rc;                              // done (0) or failed (< 0); 1: fall back to the full download
        CUDA_TRY(c, cudaStreamSynchronize(c->copy_stream));      // the early copies must not land after the full ones
    }
    return swb_download(c, results, cigar_arena, cigar_cap, cigar_used);
}

// ------------------------------------------------------------------------------------------------
// one-shot entry: single pass for small batches, two-lane pipeline for large ones
// ------------------------------------------------------------------------------------------------

__global__ void k_rebase_cigar(swb_result* res, int32_t n, long long base)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && res[p].cigar_len > 0) res[p].cigar_off += base;
}

struct ChunkPlan { int32_t p0 = 0, p1 = 0; swb_batch b; ChunkView v; };

static void add_timing(swb_timing& a, const swb_timing& t) {
    a.ms_total += t.ms_total; a.ms_prepare += t.ms_prepare; a.ms_forward += t.ms_forward; a.ms_reverse += t.ms_reverse;
    a.ms_traceback += t.ms_traceback; a.ms_h2d += t.ms_h2d; a.ms_d2h += t.ms_d2h;
    a.cells_forward += t.cells_forward; a.cells_reverse += t.cells_reverse; a.cells_band += t.cells_band;
    a.n_fast += t.n_fast; a.n_exact += t.n_exact; a.n_launches += t.n_launches; a.h2d_bytes += t.h2d_bytes; a.d2h_bytes += t.d2h_bytes;
    a.n_sw_certified += t.n_sw_certified; a.n_sw_rejected += t.n_sw_rejected; a.n_sw_verified += t.n_sw_verified;
    a.ms_band_round0 += t.ms_band_round0; a.ms_band_rest += t.ms_band_rest; a.ms_certify += t.ms_certify; a.band_rounds += t.band_rounds;
}

// Split the pairs into contiguous chunks, each with the slice of the read/window tables its pairs touch.  Callers that
// emit pairs in locus order (all of indelPost's call sites) get disjoint or barely overlapping slices; if the
// slices would re-upload much more than the tables themselves, the batch is not pipelined.
static bool plan_chunks(const swb_batch* b, std::vector<ChunkPlan>& plans)
{
    const int64_t np = b->n_pairs;
    const bool off = getenv("SWB200_NO_PIPELINE") != nullptr;
    if (off || np < 524288 || b->n_reads <= 0 || b->n_windows <= 0 || seq_shift_of(b->seq_encoding)) return false;
    // Chunk sizes: a small first chunk so that the kernels start while most of the batch is still crossing PCIe, then
    // growing ones (per-chunk fixed latencies -- host round trips, tail rounds -- favour few, large chunks).
    // SWB200_CHUNK_PLAN overrides the relative sizes ("1,2,3"), SWB200_CHUNK_PAIRS forces equal chunks of that many pairs.
    std::vector<double> weights;
    if (const char* e = getenv("SWB200_CHUNK_PAIRS")) {
        const int64_t chunkPairs = std::max<int64_t>(16384, atoll(e));
        weights.assign((size_t)std::max<int64_t>(2, (np + chunkPairs - 1) / chunkPairs), 1.0);
    } else if (const char* e2 = getenv("SWB200_CHUNK_PLAN")) {
        for (const char* q = e2; *q; ) { char* end = nullptr; const double w = strtod(q, &end); if (end == q) break; if (w > 0) weights.push_back(w); q = *end ? end + 1 : end; }
    }
    if (weights.size() < 2) weights = {1, 2, 3, 3, 3.5, 3.5};
    const int nchunks = (int)weights.size();
    std::vector<int64_t> bounds(nchunks + 1, 0);
    { double tot = 0, acc = 0; for (double w : weights) tot += w;
      for (int k = 0; k < nchunks; ++k) { acc += weights[k]; bounds[k + 1] = k + 1 == nchunks ? np : (int64_t)((double)np * acc / tot) & ~(int64_t)1; } }
    int64_t total_r = 0, total_w = 0;
    for (int32_t i = 0; i < b->n_reads; ++i) total_r = std::max<int64_t>(total_r, b->read_off[i] + std::max(b->read_len[i], 0));
    for (int32_t i = 0; i < b->n_windows; ++i) total_w = std::max<int64_t>(total_w, b->win_off[i] + std::max(b->win_len[i], 0));
    int64_t sliced = 0;
    plans.assign(nchunks, ChunkPlan());
    for (int k = 0; k < nchunks; ++k) {
        ChunkPlan& pl = plans[k];
        pl.p0 = (int32_t)bounds[k]; pl.p1 = (int32_t)bounds[k + 1];
        int32_t rlo = INT32_MAX, rhi = -1, wlo = INT32_MAX, whi = -1;
        for (int32_t p = pl.p0; p < pl.p1; ++p) {
            const int32_t r = b->pair_read[p], w = b->pair_win[p];
            if (r >= 0 && r < b->n_reads) { rlo = std::min(rlo, r); rhi = std::max(rhi, r); }
            if (w >= 0 && w < b->n_windows) { wlo = std::min(wlo, w); whi = std::max(whi, w); }
        }
        if (rhi < 0) { rlo = 0; rhi = -1; }
        if (whi < 0) { wlo = 0; whi = -1; }
        int64_t rb0 = INT64_MAX, rb1 = 0, wb0 = INT64_MAX, wb1 = 0;
        for (int32_t i = rlo; i <= rhi; ++i) { rb0 = std::min<int64_t>(rb0, b->read_off[i]); rb1 = std::max<int64_t>(rb1, b->read_off[i] + std::max(b->read_len[i], 0)); }
        for (int32_t i = wlo; i <= whi; ++i) { wb0 = std::min<int64_t>(wb0, b->win_off[i]); wb1 = std::max<int64_t>(wb1, b->win_off[i] + std::max(b->win_len[i], 0)); }
        if (rhi < rlo) { rb0 = 0; rb1 = 0; }
        if (whi < wlo) { wb0 = 0; wb1 = 0; }
        if (rb0 < 0 || wb0 < 0) return false;
        sliced += (rb1 - rb0) + (wb1 - wb0);
        pl.b = *b;
        pl.b.n_pairs = pl.p1 - pl.p0;
        pl.b.n_reads = rhi - rlo + 1; pl.b.n_windows = whi - wlo + 1;
        pl.b.reads = b->reads + rb0; pl.b.read_off = b->read_off + rlo; pl.b.read_len = b->read_len + rlo;
        pl.b.windows = b->windows + wb0; pl.b.win_off = b->win_off + wlo; pl.b.win_len = b->win_len + wlo;
        pl.b.pair_read = b->pair_read + pl.p0; pl.b.pair_win = b->pair_win + pl.p0;
        pl.b.ref_beg = b->ref_beg ? b->ref_beg + pl.p0 : nullptr; pl.b.ref_len = b->ref_len ? b->ref_len + pl.p0 : nullptr;
        pl.b.gap_open = b->gap_open + pl.p0; pl.b.gap_ext = b->gap_ext + pl.p0;
        pl.b.mask_len = b->mask_len ? b->mask_len + pl.p0 : nullptr;
        pl.v.sliced = true; pl.v.ridx_base = rlo; pl.v.widx_base = wlo; pl.v.rbyte_base = rb0; pl.v.wbyte_base = wb0;
        pl.v.n_reads_total = b->n_reads; pl.v.n_windows_total = b->n_windows;
    }
    return sliced <= (total_r + total_w) + (total_r + total_w) / 2 + (1 << 20);
}

extern "C" int swb_align_batch(swb_ctx* c, const swb_batch* b, swb_result* results, uint32_t* cigar_arena, int64_t cigar_cap, int64_t* cigar_used) {
    if (!c) return -1;
    std::vector<ChunkPlan> plans;
    const double t_call = now_ms();
    // SWB200_PIPELINE = stream (default: streamed single pass) | lanes (two-lane chunk pipeline) | off
    const char* mode = getenv("SWB200_PIPELINE");
    const bool lanes = (mode && !strcmp(mode, "lanes")) || getenv("SWB200_CHUNK_PAIRS") || getenv("SWB200_CHUNK_PLAN");
    const bool off = (mode && !strcmp(mode, "off")) || getenv("SWB200_NO_PIPELINE");
    if (b && !off && !lanes && b->n_pairs >= 262144 && b->n_reads > 0 && b->n_windows > 0) {
        if (g_trace) c->trace.clear();
        // on ANY failing exit earlier pieces may still have async copies reading the caller's host arrays and kernels in flight:
        // drain every stream before the caller gets its buffers back, and leave no half-built batch behind
        auto drain = [&]() {
            cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->bulk_stream); cudaStreamSynchronize(c->bulk_stream2);
            cudaStreamSynchronize(c->stream2); cudaStreamSynchronize(c->stream3); cudaStreamSynchronize(c->stream4);
            for (int k = 0; k < SWB_NSIDE; ++k) cudaStreamSynchronize(c->side_stream[k]);
            cudaStreamSynchronize(c->stream);
            cudaGetLastError();
        };
        int rc = align_batch_streamed(c, b, results, cigar_arena, cigar_cap, cigar_used, getenv("SWB200_ALWAYS_SCAN") != nullptr);
        if (rc == -3) {
            // a sequence blob larger than the buffers of the previous call: let everything queued so far drain, then size exactly
            drain();
            rc = align_batch_streamed(c, b, results, cigar_arena, cigar_cap, cigar_used, true);
        }
        if (rc != 0) {
            const std::string keep = c->err;
            drain();
            c->err = keep;
            if (rc != -2) { c->have_batch = false; c->computed = false; }
        }
        if (g_trace) { fprintf(stderr, "TRACE streamed:"); for (auto& e : c->trace) fprintf(stderr, " %s@%.2f", e.first, e.second - t_call); fprintf(stderr, " end@%.2f\n", now_ms() - t_call); }
        return rc;
    }
    if (!b || !plan_chunks(b, plans)) {
        int rc = swb_upload(c, b);
        if (rc) return rc;
        rc = swb_compute(c);
        if (rc) return rc;
        return swb_download(c, results, cigar_arena, cigar_cap, cigar_used);
    }
    // ---- pipelined: two lanes (stream + workspace each) driven by two host threads, chunks alternate between them,
    //      so one lane's H2D / D2H and host round trips overlap the other lane's kernels
    if (!c->sibling) {
        c->sibling = swb_create(c->device);
        if (!c->sibling) { c->err = std::string("cannot create the second pipeline lane: ") + g_create_err; return -1; }
    }
    std::atomic<long long> arena_used(0);
    std::atomic<int> failed(0);
    std::string errs[2];
    swb_timing acc[2]; memset(acc, 0, sizeof acc);
    auto worker = [&](int li) {
        swb_ctx* L = li ? c->sibling : c;
        if (cudaSetDevice(L->device) != cudaSuccess) { failed = 1; errs[li] = "cudaSetDevice failed"; return; }
        for (size_t k = (size_t)li; k < plans.size(); k += 2) {
            if (failed.load()) return;
            const ChunkPlan& pl = plans[k];
            TR(L, "chunk_begin");
            if (upload_view(L, &pl.b, pl.v) || swb_compute(L)) { failed = 1; errs[li] = L->err; return; }
            const int32_t n = pl.p1 - pl.p0;
            const long long used = (long long)L->h_bump[1];
            const long long base = arena_used.fetch_add(used);
            cudaStream_t s = L->stream;
            bool ok = true;
            ok = ok && cudaEventRecord(L->ev[EV_D2H0], s) == cudaSuccess;
            if (n > 0 && base != 0) k_rebase_cigar<<<(n + 255) / 256, 256, 0, s>>>(L->d.res, n, base);
            if (n > 0) ok = ok && cudaMemcpyAsync(results + pl.p0, L->d.res, (size_t)n * sizeof(swb_result), cudaMemcpyDeviceToHost, s) == cudaSuccess;
            if (used > 0 && base + used <= cigar_cap && cigar_arena)
                ok = ok && cudaMemcpyAsync(cigar_arena + base, L->d.cigar, (size_t)used * 4, cudaMemcpyDeviceToHost, s) == cudaSuccess;
            ok = ok && cudaEventRecord(L->ev[EV_D2H1], s) == cudaSuccess;
            ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
            if (!ok) { failed = 1; errs[li] = std::string("download: ") + cudaGetErrorString(cudaGetLastError()); return; }
            TR(L, "download_done");
            L->tm.d2h_bytes = (int64_t)n * (int64_t)sizeof(swb_result) + used * 4;
            cudaEventElapsedTime(&L->tm.ms_d2h, L->ev[EV_D2H0], L->ev[EV_D2H1]);
            cudaEventElapsedTime(&L->tm.ms_h2d, L->ev[EV_H2D0], L->ev[EV_H2D1]);
            add_timing(acc[li], L->tm);
        }
    };
    const double t_begin = now_ms();
    if (g_trace) { c->trace.clear(); c->sibling->trace.clear(); fprintf(stderr, "TRACE plan %.2f ms, %d chunks\n", t_begin - t_call, (int)plans.size()); }
    std::thread t1(worker, 1);
    worker(0);
    t1.join();
    if (g_trace) {
        for (int li = 0; li < 2; ++li) {
            swb_ctx* L = li ? c->sibling : c;
            fprintf(stderr, "TRACE lane%d:", li);
            for (auto& e : L->trace) fprintf(stderr, " %s@%.2f", e.first, e.second - t_begin);
            fprintf(stderr, "\n");
        }
    }
    cudaSetDevice(c->device);
    if (failed.load()) { c->err = !errs[0].empty() ? errs[0] : errs[1]; c->computed = false; return -1; }
    memset(&c->tm, 0, sizeof c->tm);
    add_timing(c->tm, acc[0]); add_timing(c->tm, acc[1]);
    c->computed = false; c->have_batch = false;            // the lanes hold chunks, not the batch: swb_download is not valid after this call
    const long long used = arena_used.load();
    if (cigar_used) *cigar_used = used;
    if (used > cigar_cap || (used > 0 && !cigar_arena)) { c->err = "cigar arena too small"; return -2; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CIGAR -> indel records (swb_indels.cuh)
// ------------------------------------------------------------------------------------------------
static int run_indels(swb_ctx* c, bool aos, const uint32_t* d_cigar, const int64_t* d_coff, const int32_t* d_clen, const int32_t* d_rs, const int32_t* d_qs, int32_t n,
                      int64_t* indel_off, int32_t* indel_cnt, int32_t* read_end, swb_indel* indels, int64_t cap, int64_t* used) {
    cudaStream_t s = c->stream;
    const size_t np = (size_t)std::max(n, 0);
    CUDA_TRY(c, c->b_ind_off.ensure(np * 8 + 16));
    CUDA_TRY(c, c->b_ind_cnt.ensure(np * 4 + 16));
    CUDA_TRY(c, c->b_ind_rend.ensure(np * 4 + 16));
    CUDA_TRY(c, c->b_ind_misc.ensure(16));
    const int64_t devCap = std::max<int64_t>(cap, 1);
    CUDA_TRY(c, c->b_ind_recs.ensure((size_t)devCap * sizeof(swb_indel) + 16));
    CUDA_TRY(c, cudaMemsetAsync(c->b_ind_misc.p, 0, 16, s));
    unsigned long long* bump = (unsigned long long*)c->b_ind_misc.p;
    int32_t* overflow = (int32_t*)((char*)c->b_ind_misc.p + 8);
    if (n > 0) {
        const unsigned blocks = (unsigned)((np + 127) / 128);
        if (aos) k_indels<true><<<blocks, 128, 0, s>>>(c->d.res, d_cigar, nullptr, nullptr, nullptr, nullptr, n, (int64_t*)c->b_ind_off.p, (int32_t*)c->b_ind_cnt.p, (int32_t*)c->b_ind_rend.p, (swb_indel*)c->b_ind_recs.p, devCap, bump, overflow);
        else k_indels<false><<<blocks, 128, 0, s>>>(nullptr, d_cigar, d_coff, d_clen, d_rs, d_qs, n, (int64_t*)c->b_ind_off.p, (int32_t*)c->b_ind_cnt.p, (int32_t*)c->b_ind_rend.p, (swb_indel*)c->b_ind_recs.p, devCap, bump, overflow);
        CUDA_TRY(c, cudaGetLastError());
    }
    unsigned long long h_misc[2] = {0, 0};
    CUDA_TRY(c, cudaMemcpyAsync(h_misc, c->b_ind_misc.p, 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    const int64_t need = (int64_t)h_misc[0];
    if (used) *used = need;
    if (need > cap || (need > 0 && !indels)) { c->err = "indel arena too small"; return -2; }
    if (n > 0) {
        CUDA_TRY(c, cudaMemcpyAsync(indel_off, c->b_ind_off.p, np * 8, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaMemcpyAsync(indel_cnt, c->b_ind_cnt.p, np * 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(c, cudaMemcpyAsync(read_end, c->b_ind_rend.p, np * 4, cudaMemcpyDeviceToHost, s));
    }
    if (need > 0) CUDA_TRY(c, cudaMemcpyAsync(indels, c->b_ind_recs.p, (size_t)need * sizeof(swb_indel), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    return 0;
}

extern "C" int swb_indels(swb_ctx* c, int64_t* indel_off, int32_t* indel_cnt, int32_t* read_end, swb_indel* indels, int64_t cap, int64_t* used) {
    if (!c) return -1;
    if (!c->computed) { c->err = "swb_indels: nothing computed on this context (after a two-lane pipelined swb_align_batch the results are not resident: use swb_indels_from_cigars)"; return -1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->d.n_pairs && (!indel_off || !indel_cnt || !read_end)) { c->err = "swb_indels: missing output arrays"; return -1; }
    return run_indels(c, true, c->d.cigar, nullptr, nullptr, nullptr, nullptr, c->d.n_pairs, indel_off, indel_cnt, read_end, indels, cap, used);
}

extern "C" int swb_indels_from_cigars(swb_ctx* c, int32_t n, const uint32_t* cigar_arena, int64_t arena_len, const int64_t* cigar_off, const int32_t* cigar_len,
                                      const int32_t* ref_start, const int32_t* read_start, int64_t* indel_off, int32_t* indel_cnt, int32_t* read_end,
                                      swb_indel* indels, int64_t cap, int64_t* used) {
    if (!c) return -1;
    if (n < 0 || arena_len < 0 || (n > 0 && (!cigar_off || !cigar_len || !ref_start || !read_start || !indel_off || !indel_cnt || !read_end)) || (arena_len > 0 && !cigar_arena)) { c->err = "swb_indels_from_cigars: bad arguments"; return -1; }
    for (int32_t i = 0; i < n; ++i)
        if (cigar_len[i] < 0 || cigar_off[i] < 0 || cigar_off[i] + cigar_len[i] > arena_len) { c->err = "swb_indels_from_cigars: CIGAR outside the arena"; return -1; }
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const size_t np = (size_t)n;
    CUDA_TRY(c, c->b_ind_cig.ensure((size_t)arena_len * 4 + 16));
    CUDA_TRY(c, c->b_ind_coff.ensure(np * 8 + 16));
    CUDA_TRY(c, c->b_ind_clen.ensure(np * 4 + 16));
    CUDA_TRY(c, c->b_ind_rs.ensure(np * 4 + 16));
    CUDA_TRY(c, c->b_ind_qs.ensure(np * 4 + 16));
    if (arena_len) CUDA_TRY(c, cudaMemcpyAsync(c->b_ind_cig.p, cigar_arena, (size_t)arena_len * 4, cudaMemcpyHostToDevice, s));
    if (n) {
        CUDA_TRY(c, cudaMemcpyAsync(c->b_ind_coff.p, cigar_off, np * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(c, cudaMemcpyAsync(c->b_ind_clen.p, cigar_len, np * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(c, cudaMemcpyAsync(c->b_ind_rs.p, ref_start, np * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(c, cudaMemcpyAsync(c->b_ind_qs.p, read_start, np * 4, cudaMemcpyHostToDevice, s));
    }
    return run_indels(c, false, (const uint32_t*)c->b_ind_cig.p, (const int64_t*)c->b_ind_coff.p, (const int32_t*)c->b_ind_clen.p, (const int32_t*)c->b_ind_rs.p,
                      (const int32_t*)c->b_ind_qs.p, n, indel_off, indel_cnt, read_end, indels, cap, used);
}

extern "C" void swb_rebase_cigar_offsets(swb_result* results, int64_t n, int64_t base) {
    if (!results || base == 0) return;
    for (int64_t i = 0; i < n; ++i) if (results[i].cigar_len > 0) results[i].cigar_off += base;
}

extern "C" int swb_slice_table(const int64_t* off, const int32_t* len, int64_t n, int shift, int64_t* out_off, int64_t* extent) {
    int64_t lo = INT64_MAX, hi = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (off[i] < 0 || len[i] < 0) return -1;
        lo = std::min<int64_t>(lo, off[i]); hi = std::max<int64_t>(hi, off[i] + seq_bytes(len[i], shift));
    }
    if (n <= 0) { lo = 0; hi = 0; }
    for (int64_t i = 0; i < n; ++i) out_off[i] = off[i] - lo;
    extent[0] = lo; extent[1] = hi;
    return 0;
}

extern "C" int swb_get_timing(const swb_ctx* c, swb_timing* out) {
    if (!c || !out) return -1;
    *out = c->tm;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// reference-compatible single-pair interface (ssw.h:86-139)
// ------------------------------------------------------------------------------------------------

struct _profile {            // what sswpy.pyx:59-64 declares: read, mat, readLen, n, bias
    const int8_t* read;
    const int8_t* mat;
    int32_t readLen;
    int32_t n;
    uint8_t bias;
    int8_t score_size;
};

static swb_ctx* g_default_ctx = nullptr;
static std::mutex g_align_mu;

static swb_ctx* default_ctx() {
    if (!g_default_ctx) {
        int dev = 0;
        const char* e = getenv("SWB200_DEVICE");
        if (e) dev = atoi(e);
        g_default_ctx = swb_create(dev);
    }
    return g_default_ctx;
}

extern "C" s_profile* ssw_init(const int8_t* read, const int32_t readLen, const int8_t* mat, const int32_t n, const int8_t score_size) {
    s_profile* p = (s_profile*)calloc(1, sizeof(s_profile));
    p->read = read; p->mat = mat; p->readLen = readLen; p->n = n; p->score_size = score_size;     // aliases caller memory like ssw.c:803-804
    p->bias = 0;
    if (score_size == 0 || score_size == 2) {
        int bias = 0;
        for (int i = 0; i < n * n; ++i) if (mat[i] < bias) bias = mat[i];
        p->bias = (uint8_t)std::abs(bias);
    }
    return p;
}

extern "C" void init_destroy(s_profile* p) { free(p); }

extern "C" s_align* ssw_align(const s_profile* prof, const int8_t* ref, int32_t refLen, const uint8_t weight_gapO, const uint8_t weight_gapE,
                              const uint8_t flag, const uint16_t filters, const int32_t filterd, const int32_t maskLen) {
    std::lock_guard<std::mutex> lk(g_align_mu);
    if (!prof) { fprintf(stderr, "Please call the function ssw_init before ssw_align.\n"); return nullptr; }
    swb_ctx* c = default_ctx();
    if (!c) { fprintf(stderr, "libswb200: %s\n", swb_last_error(nullptr)); return nullptr; }
    if (maskLen < 15) fprintf(stderr, "When maskLen < 15, the function ssw_align doesn't return 2nd best alignment information.\n");   // ssw.c:837-839
    if (prof->score_size != 0 && prof->score_size != 1 && prof->score_size != 2) {
        fprintf(stderr, "Please call the function ssw_init before ssw_align.\n");                                                         // ssw.c:856-859
        return nullptr;
    }
    swb_batch b; memset(&b, 0, sizeof b);
    int64_t zero = 0; int32_t rl = prof->readLen, wl = refLen, idx0 = 0, ml = maskLen;
    uint8_t go = weight_gapO, ge = weight_gapE;
    b.n_pairs = 1; b.n_reads = 1; b.n_windows = 1; b.seq_encoding = SWB_SEQ_CODES;
    b.reads = prof->read; b.read_off = &zero; b.read_len = &rl;
    b.windows = ref; b.win_off = &zero; b.win_len = &wl;
    b.pair_read = &idx0; b.pair_win = &idx0; b.gap_open = &go; b.gap_ext = &ge; b.mask_len = &ml;
    b.mat = prof->mat; b.n = prof->n; b.score_size = prof->score_size; b.flag = flag; b.filters = filters; b.filterd = filterd;
    swb_result r; int64_t used = 0;
    std::vector<uint32_t> arena((size_t)rl + (size_t)wl + 8);
    int rc = swb_align_batch(c, &b, &r, arena.data(), (int64_t)arena.size(), &used);
    if (rc) { fprintf(stderr, "libswb200: %s\n", swb_last_error(c)); return nullptr; }
    if (r.status == SWB_ERR_BYTE_ONLY) {
        fprintf(stderr, "Please set 2 to the score_size parameter of the function ssw_init, otherwise the alignment results will be incorrect.\n");   // ssw.c:849
        return nullptr;
    }
    if (r.status != SWB_OK) { fprintf(stderr, "libswb200: invalid input (empty sequence or code outside the matrix)\n"); return nullptr; }
    s_align* a = (s_align*)calloc(1, sizeof(s_align));
    a->score1 = r.score1; a->score2 = r.score2; a->ref_begin1 = r.ref_begin1; a->ref_end1 = r.ref_end1;
    a->read_begin1 = r.read_begin1; a->read_end1 = r.read_end1; a->ref_end2 = r.ref_end2; a->flag = r.flag;
    a->cigar = nullptr; a->cigarLen = 0;
    if (r.cigar_len > 0) {
        a->cigar = (uint32_t*)malloc((size_t)r.cigar_len * 4);
        memcpy(a->cigar, arena.data() + r.cigar_off, (size_t)r.cigar_len * 4);
        a->cigarLen = r.cigar_len;
    }
    return a;
}

extern "C" void align_destroy(s_align* a) { if (a) { free(a->cigar); free(a); } }

// ssw.h:171-190
extern "C" char swb_cigar_int_to_op(uint32_t v) { return (v & 0xfU) > 8 ? 'M' : "MIDNSHP=X"[v & 0xfU]; }
extern "C" uint32_t swb_cigar_int_to_len(uint32_t v) { return v >> 4; }
extern "C" uint32_t swb_to_cigar_int(uint32_t length, char op) {
    uint32_t code = 0;
    switch (op) { case 'M': code = 0; break; case 'I': code = 1; break; case 'D': code = 2; break; case 'N': code = 3; break;
                  case 'S': code = 4; break; case 'H': code = 5; break; case 'P': code = 6; break; case '=': code = 7; break; case 'X': code = 8; break; default: code = 0; }
    return (length << 4) | code;
}
