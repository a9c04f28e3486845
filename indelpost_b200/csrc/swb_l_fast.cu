// swb_l_fast.cu — launches of the DPX sweep (swb_fast.cuh); compiled once per direction and per flavour
// (-DSWB_FAST_DIR=0 / 1, -DSWB_FAST_SW=0 plain Gotoh / 1 sandwich), so the instantiations build in parallel
#include "swb_host.h"
#include "swb_fast.cuh"

#ifndef SWB_FAST_SW
#define SWB_FAST_SW 0
#endif

// fast-path launch geometry: groups of FAST_G threads, 10 bytes of shared memory per column per group
// listBase: first of the SWB_NBUCKETS job lists this launch family reads; verifyX >= 0: overflow-verification launch (SW forward only)
template <int R, int DIR, int SW, int G = FAST_G>
static int launch_fast_one(swb_ctx* c, int listBase, int bucket, int maxCols, int firstPair, int upperBoundPairs, cudaStream_t st, int verifyX = -1, bool forceSmem = false) {
    SwbDev& d = c->d;
    const int colAlloc = (std::max(maxCols, 8) + 7) & ~7;                     // longest window among this bucket's pairs
    // (forceSmem: a launch that may run beside another one of its family must not use the shared global column scratch)
    const bool globalCols = !forceSmem && colAlloc > (G == 8 ? SWB_FAST8_SMEM_COLS : SWB_FAST_SMEM_COLS);
    const size_t per = (size_t)colAlloc * (globalCols ? 2 : (G == 8 ? 8 : 10));
    // long windows: the global column-best scratch is bounded, the bucket is served in slices of the job list
    int slicePairs = upperBoundPairs - firstPair;
    if (globalCols) {
        const size_t budget = (size_t)4 << 30;
        const size_t perPairPair = (size_t)2 * colAlloc * 4;
        const size_t evenBudget = std::max<size_t>(2, 2 * (budget / perPairPair));      // pairs per slice (even: lane pairs stay intact)
        slicePairs = (size_t)slicePairs <= evenBudget ? slicePairs : (int)evenBudget;
        CUDA_TRY(c, c->b_fastcols.ensure((size_t)((slicePairs + 1) / 2) * perPairPair + 16));
        d.fast_cols = (uint32_t*)c->b_fastcols.p;
    } else d.fast_cols = nullptr;
    int groups = 128 / G;
    while (groups > 2 && groups * per > (size_t)c->smem_optin - 1024) groups /= 2;
    const int threads = groups * G;
    static std::atomic<bool> attr_set[SWB_MAX_DEVICES] = {};               // per template instantiation and device
    if (!attr_set[c->device % SWB_MAX_DEVICES]) {
        cudaFuncSetAttribute(k_fast<R, DIR, false, SW, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024);
        cudaFuncSetAttribute(k_fast<R, DIR, true, SW, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024);
        attr_set[c->device % SWB_MAX_DEVICES] = true;
    }
    const int slot = listBase + bucket;
    for (int off = firstPair; off < upperBoundPairs; off += slicePairs) {
        const int n = std::min(slicePairs, upperBoundPairs - off);
        const int ngroups = (n + 1) / 2;
        int blocks = (ngroups + groups - 1) / groups;
        // the sandwich lists are launched against an upper bound of their length (the forward sweep appends to them): a bounded
        // grid that strides over the list instead of tens of thousands of blocks that find nothing to do
        if (SW) blocks = std::min(blocks, c->n_sm * 8);
        if (globalCols) k_fast<R, DIR, true, SW, G><<<blocks, threads, groups * per, st>>>(d, d.list[slot], d.counters + slot, colAlloc, off, slicePairs, verifyX);
        else k_fast<R, DIR, false, SW, G><<<blocks, threads, groups * per, st>>>(d, d.list[slot], d.counters + slot, colAlloc, off, slicePairs, verifyX);
        c->tm.n_launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

template <int DIR, int SW>
static int launch_fast_bucket(swb_ctx* c, int listBase, int b, int rowsBucket, int maxCols, int f, int n, cudaStream_t st, int verifyX = -1, bool forceSmem = false) {
    switch (rowsBucket) {      // R = 2 * (bucket + 1) rows per thread
        case 0: return launch_fast_one<2, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
        case 1: return launch_fast_one<4, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
        case 2: return launch_fast_one<6, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
        case 3: return launch_fast_one<8, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
        case 4: return launch_fast_one<10, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
        case 5: return launch_fast_one<12, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
        case 6: return launch_fast_one<14, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
        case 7: return launch_fast_one<16, DIR, SW>(c, listBase, b, maxCols, f, n, st, verifyX, forceSmem);
    }
    return -1;
}

// one launch per non-empty read-length bucket over the list ranges [first[b], counts[b])
template <int DIR, int SW>
static int launch_fast_range(swb_ctx* c, int listBase, const int* first, const int* counts, cudaStream_t st) {
    for (int b = 0; b < SWB_NBUCKETS; ++b) {
        const int n = counts[b], f = first ? first[b] : 0;
        if (n <= f) continue;
        const int rc = launch_fast_bucket<DIR, SW>(c, listBase, b, b, c->fastMaxCols[b], f, n, st);
        if (rc) return rc;
    }
    return 0;
}

#ifdef SWB_FAST_G8
// forward sweep with 8 threads per lane pair: families SWB_NBUCKETS .. SWB_NFWD-1 (first / counts / c->fastMaxCols are indexed by family)
int swb_launch_fast8_range_fwd(swb_ctx* c, const int* first, const int* counts, cudaStream_t st) {
    for (int b = 0; b < SWB_NF8; ++b) {
        const int fam = SWB_NBUCKETS + b;
        const int n = counts[fam], f = first ? first[fam] : 0;
        if (n <= f) continue;
        int rc = -1;
        switch (b) {      // rows per thread = padded length / 8
            case 0: rc = launch_fast_one<4, 0, 0, 8>(c, LIST_F8_FWD, b, c->fastMaxCols[fam], f, n, st); break;
            case 1: rc = launch_fast_one<7, 0, 0, 8>(c, LIST_F8_FWD, b, c->fastMaxCols[fam], f, n, st); break;
            case 2: rc = launch_fast_one<10, 0, 0, 8>(c, LIST_F8_FWD, b, c->fastMaxCols[fam], f, n, st); break;
            case 3: rc = launch_fast_one<13, 0, 0, 8>(c, LIST_F8_FWD, b, c->fastMaxCols[fam], f, n, st); break;
            case 4: rc = launch_fast_one<16, 0, 0, 8>(c, LIST_F8_FWD, b, c->fastMaxCols[fam], f, n, st); break;
            case 5: rc = launch_fast_one<19, 0, 0, 8>(c, LIST_F8_FWD, b, c->fastMaxCols[fam], f, n, st); break;
        }
        if (rc) return rc;
    }
    return 0;
}
#elif SWB_FAST_SW == 0
#if SWB_FAST_DIR == 0
int swb_launch_fast_range_fwd(swb_ctx* c, const int* first, const int* counts, cudaStream_t st) { return launch_fast_range<0, 0>(c, LIST_FAST_FWD, first, counts, st); }
#else
int swb_launch_fast_range_rev(swb_ctx* c, const int* first, const int* counts, cudaStream_t st) { return launch_fast_range<1, 0>(c, LIST_FAST_REV, first, counts, st); }
#endif
#else
#if SWB_FAST_DIR == 0
int swb_launch_sandwich_fwd(swb_ctx* c, const int* counts, cudaStream_t st) { return launch_fast_range<0, 1>(c, LIST_SW_FWD, nullptr, counts, st); }
// overflow verification of the pairs in list `listSlot` (any read length of the batch: the widest bucket's instantiation, reads are
// end-aligned in its rows); pairs the lower bound cannot settle are appended to list `xSlot`.  Returns 1 if the batch's windows are
// too long for the shared-memory column scratch (the caller then goes to the exact kernel directly).
int swb_launch_sandwich_verify(swb_ctx* c, int listSlot, int xSlot, int upperBound, cudaStream_t st) {
    const SwbDev& d = c->d;
    int maxCols = 0;
    for (int f = 0; f < SWB_NFWD; ++f) maxCols = std::max(maxCols, c->fastMaxCols[f]);
    const int lp16 = (std::min(d.max_rlen, 32 * SWB_NBUCKETS) + 15) & ~15;
    if (maxCols > 1024 || d.max_rlen > 32 * SWB_NBUCKETS || upperBound <= 0) return 1;      // 10 KB of shared memory per lane pair at most
    const int rb = std::max(0, (lp16 + 31) / 32 - 1);
    return launch_fast_bucket<0, 1>(c, listSlot, 0, rb, maxCols, 0, upperBound, st, xSlot, /*forceSmem=*/true);      // two verifications may run side by side
}
#else
int swb_launch_sandwich_rev(swb_ctx* c, const int* counts, cudaStream_t st) { return launch_fast_range<1, 1>(c, LIST_SW_REV, nullptr, counts, st); }
#endif
#endif
