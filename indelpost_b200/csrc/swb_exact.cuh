// swb_exact.cuh — exact emulation of ssw.c's striped SIMD schedule on the GPU.
//
// This kernel reproduces sw_sse2_byte (ssw.c:197-384) and sw_sse2_word (ssw.c:410-586) for ANY input,
// including the cases where the striped algorithm is not plain Smith-Waterman (the
// signed-byte lazy-F exit test of ssw.c:311; gap_open <= gap_extension; 8-bit saturation).  It is the
// always-correct path; the DPX kernel in swb_fast.cuh takes the pairs for which the striped result is
// provably plain Gotoh and everything else lands here.
//
// Mapping: one SSE2 lane = one GPU thread.  A group of W threads (W=16 in byte mode, W=8 in word mode)
// owns one alignment; thread l holds the rows l*segLen .. (l+1)*segLen-1 exactly like SIMD lane l of
// the striped layout (ssw.c:169-186).  `_mm_slli_si128` becomes __shfl_up_sync inside the group and
// `_mm_movemask_epi8(cmpgt)` becomes a ballot.  The 2 (byte) / 4 (word) groups of a warp run in lock
// step: every loop bound is the warp maximum and a group that has finished a phase is predicated off,
// so the cost of a warp is the maximum over its groups, not the sum of divergent paths.
//
// Shared memory per group: query profile (n x segLen x W int8), H, E and the best column (W x segLen
// elements each), all laid out [j][lane] so the W lanes of a group touch consecutive addresses.
#pragma once
#include "swb_common.cuh"

template <int MODE> struct ExactTraits;
template <> struct ExactTraits<0> { static constexpr int W = 16; typedef uint8_t elem; static constexpr int SATMAX = 255; };
template <> struct ExactTraits<1> { static constexpr int W = 8;  typedef int16_t elem; static constexpr int SATMAX = 32767; };

__host__ __device__ inline int exact_smem_per_group(int mode, int n, int max_rlen) {
    const int W = mode ? 8 : 16;
    const int seg = (max_rlen + W - 1) / W;
    const int esz = mode ? 2 : 1;
    int bytes = n * seg * W + 3 * seg * W * esz;
    return (bytes + 15) & ~15;
}

// MODE 0: byte, 1: word.  DIR 0: forward pass (ssw.c:843-847), 1: reverse pass (ssw.c:875-886).
template <int MODE, int DIR>
__global__ void __launch_bounds__(128)
k_exact(SwbDev d, const int32_t* __restrict__ jobs, const int32_t* __restrict__ njobs_ptr, int segAlloc, int smemPerGroup)
{
    typedef ExactTraits<MODE> TR;
    typedef typename TR::elem elem;
    constexpr int W = TR::W;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int njobs = *njobs_ptr;
    const int lane = threadIdx.x & 31;
    const int gl = lane % W;                          // SIMD lane inside the group
    const int gw = lane / W;                          // group inside the warp
    const int groupInBlock = threadIdx.x / W;
    const int job = (blockIdx.x * (blockDim.x / W)) + groupInBlock;
    if (blockIdx.x * (blockDim.x / W) >= njobs) return;          // whole block idle
    const bool valid = job < njobs;
    const int p = valid ? jobs[job] : -1;

    // ---- per-alignment parameters --------------------------------------------------------------
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

    unsigned char* gbase = smem_raw + (size_t)groupInBlock * smemPerGroup;
    int8_t* prof = reinterpret_cast<int8_t*>(gbase);                       // [nt][j][lane]
    elem* Hc = reinterpret_cast<elem*>(gbase + n * segAlloc * W);            // [j][lane]
    elem* Ec = Hc + segAlloc * W;
    elem* Hb = Ec + segAlloc * W;

    // ---- query profile (qP_byte ssw.c:163-188 / qP_word ssw.c:386-408) -------------------------
    for (int idx = gl; idx < n * segLen * W; idx += W) {
        const int nt = idx / (segLen * W);
        const int rem = idx - nt * segLen * W;
        const int j = rem / W, l = rem - j * W;
        const int r = j + l * segLen;
        int v;
        if (r >= rl) v = MODE ? 0 : bias;
        else {
            const int rb = DIR ? read[rl - 1 - r] : read[r];               // seq_reverse, ssw.c:774-785
            v = d.mat[nt * n + rb] + (MODE ? 0 : bias);
        }
        prof[(nt * segAlloc + j) * W + l] = (int8_t)v;
    }
    for (int idx = gl; idx < segLen * W; idx += W) { Hc[idx] = 0; Ec[idx] = 0; Hb[idx] = 0; }
    __syncwarp();

    // warp-wide loop bounds
    int maxSeg = segLen, maxCols = cols;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        maxSeg = max(maxSeg, __shfl_xor_sync(FULL, maxSeg, o));
        maxCols = max(maxCols, __shfl_xor_sync(FULL, maxCols, o));
    }

    int best = 0;
    int end_ref = MODE ? 0 : -1;                      // ssw.c:427 vs ssw.c:220
    bool alive = valid && cols > 0;                   // group-uniform
    bool overflow = false;
    uint16_t* colmax = valid ? d.colmax + (size_t)p * d.colmax_stride : nullptr;
    long long cells = 0;

    for (int c = 0; c < maxCols; ++c) {
        const bool act = alive && c < cols;
        const int i = DIR ? cols - 1 - c : c;
        const int nt = act ? ref[i] : 0;
        const int8_t* P = prof + nt * segAlloc * W + gl;

        // vH = last stripe of the previous column shifted by one lane (ssw.c:264-265 / 467-468)
        int vH = (act && segLen > 0) ? (int)Hc[(segLen - 1) * W + gl] : 0;
        vH = __shfl_up_sync(FULL, vH, 1, W);
        if (gl == 0) vH = 0;
        int vF = 0, vMax = 0;

        // ---- main striped sweep (ssw.c:274-299 / 480-504) --------------------------------------
        for (int j = 0; j < maxSeg; ++j) {
            if (act && j < segLen) {
                const int pj = MODE ? (int)P[j * W] : (int)(uint8_t)P[j * W];
                int e = (int)Ec[j * W + gl];
                const int hOld = (int)Hc[j * W + gl];
                int h = min(vH + pj, TR::SATMAX);
                if (MODE == 0) h = max(h - bias, 0);
                h = max(max(h, e), vF);
                vMax = max(vMax, h);
                Hc[j * W + gl] = (elem)h;
                h = max(h - go, 0);
                e = max(max(e - ge, 0), h);
                Ec[j * W + gl] = (elem)e;
                vF = max(max(vF - ge, 0), h);
                vH = hOld;
            }
        }
        if (act) cells += (long long)segLen * W;

        // ---- lazy-F loop with the group-wide early exit (ssw.c:302-313 / 507-518) --------------
        bool lazy = act;
        for (int k = 0; k < W; ++k) {
            if (!__any_sync(FULL, lazy)) break;
            int sh = __shfl_up_sync(FULL, vF, 1, W);
            if (lazy) vF = gl == 0 ? 0 : sh;
            for (int j = 0; j < maxSeg; ++j) {
                bool pred = false;
                const bool on = lazy && j < segLen;
                if (on) {
                    int h = (int)Hc[j * W + gl];
                    h = max(h, vF);
                    vMax = max(vMax, h);
                    Hc[j * W + gl] = (elem)h;
                    h = max(h - go, 0);
                    vF = max(vF - ge, 0);
                    if (MODE == 0) pred = (int)(int8_t)vF > (int)(int8_t)h;    // signed byte compare, ssw.c:311
                    else pred = vF > h;
                }
                const unsigned b = __ballot_sync(FULL, pred);
                if (on && ((b >> (gw * W)) & ((1u << W) - 1u)) == 0u) lazy = false;
                if (j + 1 >= maxSeg || !__any_sync(FULL, lazy && j + 1 < segLen)) break;
            }
        }

        // ---- column maximum, running maximum, best column (ssw.c:316-337 / 521-539) ------------
        int cm = vMax;
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) cm = max(cm, __shfl_xor_sync(FULL, cm, o, W));
        if (act) {
            bool stop = false;
            if (cm > best) {
                best = cm;
                if (MODE == 0 && best + bias >= 255) { overflow = true; stop = true; }
                else {
                    end_ref = i;
                    for (int j = 0; j < segLen; ++j) Hb[j * W + gl] = Hc[j * W + gl];
                }
            }
            if (!stop) {
                if (gl == 0) colmax[i] = (uint16_t)cm;
                if (cm == terminate) stop = true;
            }
            if (stop) alive = false;
        }
        if (!__any_sync(FULL, alive)) break;
    }

    if (!valid) return;
    const unsigned GM = (W == 32) ? FULL : (((1u << W) - 1u) << (gw * W));   // from here on groups may diverge: group-scoped shuffles only

    // ---- smallest read index holding the maximum in the best column (ssw.c:341-349 / 543-551) --
    int end_read = rl - 1;
    for (int j = 0; j < segLen; ++j) {
        if ((int)Hb[j * W + gl] == best) { const int r = j + gl * segLen; if (r < end_read) end_read = r; }
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) end_read = min(end_read, __shfl_xor_sync(GM, end_read, o, W));

    if (gl == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(d.counters + (DIR ? CNT_CELLS_REV : CNT_CELLS_FWD)), (unsigned long long)cells * 1ull);
    }

    swb_result& r = d.res[p];
    if (DIR == 0) {
        // ---- sub-optimal score: first strict maximum outside the mask (ssw.c:366-379 / 568-581) ----
        int s2 = 0, r2 = 0;
        if (!overflow) {
            __syncwarp(GM);
            const int edgeL = max(end_ref - maskLen, 0);
            const int edgeR = min(end_ref + maskLen, cols) + (MODE ? 0 : 1);
            int bv = 0, bi = 0x7fffffff;
            for (int i = gl; i < cols; i += W) {
                if (i < edgeL || i >= edgeR) {
                    const int v = colmax[i];
                    if (v > bv || (v == bv && v > 0 && i < bi)) { bv = v; bi = i; }
                }
            }
#pragma unroll
            for (int o = W / 2; o > 0; o >>= 1) {
                const int ov = __shfl_xor_sync(GM, bv, o, W), oi = __shfl_xor_sync(GM, bi, o, W);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (bv > 0) { s2 = bv; r2 = bi; }
        }
        if (gl == 0) {
            atomicAdd(d.counters + CNT_EXACT_JOBS, 1);
            if (MODE == 0 && overflow) {
                // 8-bit pass overflowed (ssw.c:358, 844-852)
                if (d.p_state[p] & PST_HAVE_WORD) { /* verification only: the 16-bit result already stored stands */ }
                else if (d.score_size == 2) list_push(d.list[LIST_WORD_FWD], d.counters + CNT_WORD_FWD, p);
                else { r.status = SWB_ERR_BYTE_ONLY; }
            } else {
                if (d.p_state[p] & PST_HAVE_WORD) atomicAdd(d.counters + CNT_VERIFY_BYTE, 1);   // the provisional 16-bit result was wrong: redo in 8-bit semantics
                d.p_state[p] = 0; d.t_bw[p] = 0; d.t_best[p] = 0;
                r.ref_begin1 = -1; r.read_begin1 = -1; r.cigar_len = 0; r.cigar_off = 0; r.flag = 0;
                r.score1 = (uint16_t)best; r.ref_end1 = end_ref; r.read_end1 = end_read;
                if (maskLen >= 15) { r.score2 = (uint16_t)s2; r.ref_end2 = r2; }      // ssw.c:864-870
                else { r.score2 = 0; r.ref_end2 = -1; }
                d.p_mode[p] = (uint8_t)MODE;
                const bool scoreOnly = d.flag == 0 || (d.flag == 2 && best < (int)d.filters);   // ssw.c:872
                if (!scoreOnly) list_push(d.list[MODE ? LIST_WORD_REV : LIST_BYTE_REV], d.counters + (MODE ? CNT_WORD_REV : CNT_BYTE_REV), p);
            }
        }
    } else {
        if (gl == 0) {
            r.ref_begin1 = end_ref;                                    // ssw.c:885-886
            r.read_begin1 = r.read_end1 - end_read;
            if ((int)r.score1 > best) r.flag = 2;                      // ssw.c:888-891
            const int f = d.flag;
            const bool noCigar = (7 & f) == 0 || ((2 & f) != 0 && (int)r.score1 < (int)d.filters) ||
                                 ((4 & f) != 0 && (r.ref_end1 - r.ref_begin1 > d.filterd || r.read_end1 - r.read_begin1 > d.filterd));   // ssw.c:894
            if (!noCigar) push_band(d, p, r); else d.p_state[p] |= PST_BAND_DONE;
        }
    }
}
