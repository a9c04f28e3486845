// swb_exact2.cuh — packed variant of the exact striped-schedule emulation (see swb_exact.cuh for the semantics).
//
// Same literal emulation of sw_sse2_byte / sw_sse2_word (ssw.c:197-586), but every thread carries TWO adjacent SSE2
// lanes in the 16-bit halves of one register and updates them with native packed instructions:
//   _mm_adds_epu8(a,b)  -> __viaddmin_u16x2(a, b, 255)          _mm_subs_epu8(a,b) -> __viaddmax_s16x2(a, -b, 0)
//   _mm_max_epu8        -> VIMNMX / VIMNMX3 .S16x2 (values are 0..255, so signed 16-bit order = unsigned 8-bit order)
//   _mm_cmpgt_epi8      -> flip bit 7 of both operands, unsigned 16-bit max, "changed?" (signed-byte order preserved)
//   _mm_slli_si128(v,1) -> __shfl_up of the neighbour's register + one PRMT
// A group is 8 threads (8-bit mode, 16 lanes) or 4 threads (16-bit mode, 8 lanes); the 4 / 8 groups of a warp run in
// lock step like in swb_exact.cuh.  8-bit values live in 16-bit lanes, so nothing saturates spuriously; the 16-bit
// mode requires max(mat)*readLen <= 32000 (otherwise k_exact's scalar lanes are used).
#pragma once
#include "swb_common.cuh"

__device__ __forceinline__ uint32_t x2_prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }
__device__ __forceinline__ uint32_t x2_max_s(uint32_t a, uint32_t b) { uint32_t d; asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t x2_max_u(uint32_t a, uint32_t b) { uint32_t d; asm("max.u16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t x2_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

#define SWB_EXACT_SPARSE_MAX 64     // most jobs the one-job-per-warp schedule of k_exact2 serves

__host__ __device__ inline int exact2_smem_per_group(int mode, int n, int max_rlen) {
    const int W = mode ? 8 : 16;
    const int seg = (max_rlen + W - 1) / W;
    return (n + 3) * seg * (W / 2) * 4;             // profile + H + E + best column, one 32-bit word per lane pair
}

template <int MODE, int DIR>
__global__ void __launch_bounds__(128)
k_exact2(SwbDev d, const int32_t* __restrict__ jobs, const int32_t* __restrict__ njobs_ptr, int segAlloc, int smemPerGroup, int sched)
{
    // sched 0: one group per job, the groups of a warp in lock step (throughput).
    // The groups of a warp advance column by column together, and a column costs every group as much as the slowest one
    // (the lazy-F loop runs 16*segLen steps once scores pass 128): a handful of jobs is served much faster one per WARP.
    // sched 1 / 2 are the two halves of a launch pair that picks by the job count on the device (no host round trip):
    //   1 = as 0 but only if there are more than SWB_EXACT_SPARSE_MAX jobs,  2 = one job per warp, only if there are at most that many.
    constexpr int W = MODE ? 8 : 16;                  // SSE2 lanes per alignment
    constexpr int T = W / 2;                          // threads per alignment
    constexpr unsigned FULL = 0xffffffffu;
    constexpr unsigned GBITS = (1u << T) - 1u;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int njobs = *njobs_ptr;
    const int lane = threadIdx.x & 31;
    const int gt = lane % T;                          // thread inside the group: SSE2 lanes 2*gt and 2*gt+1
    const int gw = lane / T;                          // group inside the warp
    const int groupInBlock = threadIdx.x / T;
    if (sched == 1 && njobs <= SWB_EXACT_SPARSE_MAX) return;
    if (sched == 2 && njobs > SWB_EXACT_SPARSE_MAX) return;
    const int job = sched == 2 ? (int)(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) : (int)(blockIdx.x * (blockDim.x / T) + groupInBlock);
    if ((sched == 2 ? blockIdx.x * (blockDim.x / 32) : blockIdx.x * (blockDim.x / T)) >= (unsigned)njobs) return;
    const bool valid = job < njobs && (sched != 2 || gw == 0);
    const int p = valid ? jobs[job] : -1;

    int rl = 0, cols = 0, go = 0, ge = 0, maskLen = 0, terminate = 0;
    const int8_t* read = nullptr; const int8_t* ref = nullptr;
    const int n = d.n;
    const int bias = d.bias;
    if (valid) {
        read = d.reads + d.p_roff[p];
        ref = d.windows + d.p_woff[p];
        go = d.gap_open[p]; ge = d.gap_ext[p];
        maskLen = d.p_mask[p];
        if (DIR == 0) { rl = d.p_rlen[p]; cols = d.p_wlen[p]; terminate = MODE ? 65535 : 255; }
        else {
            const swb_result& r = d.res[p];
            rl = r.read_end1 + 1; cols = r.ref_end1 + 1; terminate = MODE ? r.score1 : (r.score1 & 255);
        }
    }
    const int segLen = (rl + W - 1) / W;

    uint32_t* gbase = reinterpret_cast<uint32_t*>(smem_raw + (size_t)groupInBlock * smemPerGroup);
    uint32_t* prof = gbase;                           // [nt][j][t]
    uint32_t* Hc = gbase + n * segAlloc * T;          // [j][t]
    uint32_t* Ec = Hc + segAlloc * T;
    uint32_t* Hb = Ec + segAlloc * T;

    // query profile, two striped rows per word (qP_byte ssw.c:163-188 / qP_word ssw.c:386-408)
    for (int idx = gt; idx < n * segLen * T; idx += T) {
        const int nt = idx / (segLen * T);
        const int rem = idx - nt * segLen * T;
        const int j = rem / T, t = rem - j * T;
        int v[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = j + (2 * t + u) * segLen;
            if (r >= rl) v[u] = MODE ? 0 : bias;
            else {
                const int rb = DIR ? read[rl - 1 - r] : read[r];
                v[u] = d.mat[nt * n + rb] + (MODE ? 0 : bias);
                if (MODE == 0) v[u] &= 0xff;          // stored through an int8_t and reloaded unsigned (ssw.c:174, 182)
            }
        }
        prof[(nt * segAlloc + j) * T + t] = x2_pack(v[0], v[1]);
    }
    for (int idx = gt; idx < segLen * T; idx += T) { Hc[idx] = 0; Ec[idx] = 0; Hb[idx] = 0; }
    __syncwarp();

    int maxSeg = segLen, maxCols = cols;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        maxSeg = max(maxSeg, __shfl_xor_sync(FULL, maxSeg, o));
        maxCols = max(maxCols, __shfl_xor_sync(FULL, maxCols, o));
    }

    const uint32_t nGo = x2_pack(-go, -go), nGe = x2_pack(-ge, -ge), nBias = x2_pack(-bias, -bias);
    const uint32_t satP = MODE ? 0x7fff7fffu : 0x00ff00ffu;
    int best = 0;
    int end_ref = MODE ? 0 : -1;
    bool alive = valid && cols > 0;
    bool overflow = false;
    uint16_t* colmax = valid ? d.colmax + (size_t)p * d.colmax_stride : nullptr;
    long long cells = 0;

    for (int c = 0; c < maxCols; ++c) {
        const bool act = alive && c < cols;
        const int i = DIR ? cols - 1 - c : c;
        const int nt = act ? ref[i] : 0;
        const uint32_t* P = prof + nt * segAlloc * T + gt;

        // vH = last stripe of the previous column shifted up by one SSE2 lane (ssw.c:264-265 / 467-468)
        uint32_t vH = (act && segLen > 0) ? Hc[(segLen - 1) * T + gt] : 0u;
        {
            uint32_t up = __shfl_up_sync(FULL, vH, 1, T);
            if (gt == 0) up = 0;
            vH = x2_prmt(up, vH, 0x5432u);            // lo = neighbour's hi lane, hi = own lo lane
        }
        uint32_t vF = 0, vMax = 0;

        for (int j = 0; j < maxSeg; ++j) {            // ssw.c:274-299 / 480-504
            if (act && j < segLen) {
                uint32_t e = Ec[j * T + gt];
                const uint32_t hOld = Hc[j * T + gt];
                uint32_t h;
                if (MODE == 0) {
                    h = __viaddmin_u16x2(vH, P[j * T], satP);         // adds_epu8
                    h = __viaddmax_s16x2(h, nBias, 0u);               // subs_epu8(bias)
                } else {
                    h = __vadd2(vH, P[j * T]);                        // adds_epi16 (no saturation reachable: host checks the range)
                }
                h = __vimax3_s16x2(h, e, vF);
                vMax = x2_max_s(vMax, h);
                Hc[j * T + gt] = h;
                h = __viaddmax_s16x2(h, nGo, 0u);                     // subs_epu(gapO)
                e = __viaddmax_s16x2_relu(e, nGe, h);                 // max(subs_epu(e, gapE), h)
                Ec[j * T + gt] = e;
                vF = __viaddmax_s16x2_relu(vF, nGe, h);
                vH = hOld;
            }
        }
        if (act) cells += (long long)segLen * W;

        // lazy-F loop with the group-wide early exit (ssw.c:302-313 / 507-518)
        bool lazy = act;
        const unsigned GMW = GBITS << (gw * T);
        for (int k = 0; k < W; ++k) {
            if (!__any_sync(FULL, lazy)) break;
            {
                uint32_t up = __shfl_up_sync(FULL, vF, 1, T);
                if (gt == 0) up = 0;
                const uint32_t sh = x2_prmt(up, vF, 0x5432u);
                if (lazy) vF = sh;
            }
            for (int j = 0; j < maxSeg; ++j) {
                bool pred = false;
                const bool on = lazy && j < segLen;
                if (on) {
                    uint32_t h = Hc[j * T + gt];
                    h = x2_max_s(h, vF);
                    vMax = x2_max_s(vMax, h);
                    Hc[j * T + gt] = h;
                    h = __viaddmax_s16x2(h, nGo, 0u);
                    vF = __viaddmax_s16x2(vF, nGe, 0u);
                    if (MODE == 0) {                                   // _mm_cmpgt_epi8: SIGNED byte compare (ssw.c:311)
                        const uint32_t a = vF ^ 0x00800080u, b = h ^ 0x00800080u;
                        pred = x2_max_u(a, b) != b;
                    } else {
                        pred = x2_max_s(vF, h) != h;                   // _mm_cmpgt_epi16 (ssw.c:516)
                    }
                }
                const unsigned b = __ballot_sync(FULL, pred);
                if (on && (b & GMW) == 0u) lazy = false;
                if (j + 1 >= maxSeg || !__any_sync(FULL, lazy && j + 1 < segLen)) break;
            }
        }

        // column maximum, running maximum, best column (ssw.c:316-337 / 521-539)
        int cm = max((int)(int16_t)(vMax & 0xffffu), (int)(int16_t)(vMax >> 16));
#pragma unroll
        for (int o = T / 2; o > 0; o >>= 1) cm = max(cm, __shfl_xor_sync(FULL, cm, o, T));
        if (act) {
            bool stop = false;
            if (cm > best) {
                best = cm;
                if (MODE == 0 && best + bias >= 255) { overflow = true; stop = true; }
                else {
                    end_ref = i;
                    for (int j = 0; j < segLen; ++j) Hb[j * T + gt] = Hc[j * T + gt];
                }
            }
            if (!stop) {
                if (gt == 0) colmax[i] = (uint16_t)cm;
                if (cm == terminate) stop = true;
            }
            if (stop) alive = false;
        }
        if (!__any_sync(FULL, alive)) break;
    }

    if (!valid) return;
    const unsigned GM = GBITS << (gw * T);

    // smallest read index holding the maximum in the best column (ssw.c:341-349 / 543-551)
    int end_read = rl - 1;
    for (int j = 0; j < segLen; ++j) {
        const uint32_t hb = Hb[j * T + gt];
        if ((int)(hb & 0xffffu) == best) { const int r = j + (2 * gt) * segLen; if (r < end_read) end_read = r; }
        if ((int)(hb >> 16) == best) { const int r = j + (2 * gt + 1) * segLen; if (r < end_read) end_read = r; }
    }
#pragma unroll
    for (int o = T / 2; o > 0; o >>= 1) end_read = min(end_read, __shfl_xor_sync(GM, end_read, o, T));

    if (gt == 0) warp_count(d.counters + (DIR ? CNT_CELLS_REV : CNT_CELLS_FWD), (unsigned long long)cells);

    swb_result& r = d.res[p];
    if (DIR == 0) {
        int s2 = 0, r2 = 0;
        if (!overflow) {                               // sub-optimal score outside the mask (ssw.c:366-379 / 568-581)
            __syncwarp(GM);
            const int edgeL = max(end_ref - maskLen, 0);
            const int edgeR = min(end_ref + maskLen, cols) + (MODE ? 0 : 1);
            int bv = 0, bi = 0x7fffffff;
            for (int i = gt; i < cols; i += T) {
                if (i < edgeL || i >= edgeR) {
                    const int v = colmax[i];
                    if (v > bv) { bv = v; bi = i; }
                }
            }
#pragma unroll
            for (int o = T / 2; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(GM, bv, o, T), oi = __shfl_xor_sync(GM, bi, o, T);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (bv > 0) { s2 = bv; r2 = bi; }
        }
        if (gt == 0) {
            atomicAdd(d.counters + CNT_EXACT_JOBS, 1);
            if (MODE == 0 && overflow) {
                if (d.p_state[p] & PST_HAVE_WORD) { /* verification only: the 16-bit result already stored stands */ }
                else if (d.score_size == 2) list_push(d.list[LIST_WORD_FWD], d.counters + CNT_WORD_FWD, p);
                else { r.status = SWB_ERR_BYTE_ONLY; }
            } else {
                if (d.p_state[p] & PST_HAVE_WORD) atomicAdd(d.counters + CNT_VERIFY_BYTE, 1);
                d.p_state[p] = 0; d.t_bw[p] = 0; d.t_best[p] = 0;
                r.ref_begin1 = -1; r.read_begin1 = -1; r.cigar_len = 0; r.cigar_off = 0; r.flag = 0;
                r.score1 = (uint16_t)best; r.ref_end1 = end_ref; r.read_end1 = end_read;
                if (maskLen >= 15) { r.score2 = (uint16_t)s2; r.ref_end2 = r2; } else { r.score2 = 0; r.ref_end2 = -1; }
                d.p_mode[p] = (uint8_t)MODE;
                const bool scoreOnly = d.flag == 0 || (d.flag == 2 && best < (int)d.filters);
                if (!scoreOnly) list_push(d.list[MODE ? LIST_WORD_REV : LIST_BYTE_REV], d.counters + (MODE ? CNT_WORD_REV : CNT_BYTE_REV), p);
            }
        }
    } else {
        if (gt == 0) {
            r.ref_begin1 = end_ref;
            r.read_begin1 = r.read_end1 - end_read;
            if ((int)r.score1 > best) r.flag = 2;
            const int f = d.flag;
            const bool noCigar = (7 & f) == 0 || ((2 & f) != 0 && (int)r.score1 < (int)d.filters) ||
                                 ((4 & f) != 0 && (r.ref_end1 - r.ref_begin1 > d.filterd || r.read_end1 - r.read_begin1 > d.filterd));
            if (!noCigar) push_band(d, p, r); else d.p_state[p] |= PST_BAND_DONE;
        }
    }
}
