// swb_l_fast.cu — launches of the DPX sweep (swb_fast.cuh); compiled once per direction (-DSWB_FAST_DIR=0 / 1)
#include "swb_host.h"
#include "swb_fast.cuh"

// fast-path launch geometry: groups of FAST_G threads, 10 bytes of shared memory per column per group

template <int R, int DIR>
static int launch_fast_one(swb_ctx* c, int bucket, int firstPair, int upperBoundPairs, cudaStream_t st) {
    SwbDev& d = c->d;
    const int colAlloc = (std::max(c->fastMaxCols[bucket], 8) + 7) & ~7;      // longest window among this bucket's pairs
    const bool globalCols = colAlloc > SWB_FAST_SMEM_COLS;
    const size_t per = (size_t)colAlloc * (globalCols ? 2 : 10);
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
    int groups = 128 / FAST_G;
    while (groups > 2 && groups * per > (size_t)c->smem_optin - 1024) groups /= 2;
    const int threads = groups * FAST_G;
    static std::atomic<bool> attr_set[SWB_MAX_DEVICES] = {};               // per template instantiation and device
    if (!attr_set[c->device % SWB_MAX_DEVICES]) {
        cudaFuncSetAttribute(k_fast<R, DIR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024);
        cudaFuncSetAttribute(k_fast<R, DIR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 1024);
        attr_set[c->device % SWB_MAX_DEVICES] = true;
    }
    const int slot = (DIR ? LIST_FAST_REV : LIST_FAST_FWD) + bucket;
    for (int off = firstPair; off < upperBoundPairs; off += slicePairs) {
        const int n = std::min(slicePairs, upperBoundPairs - off);
        const int ngroups = (n + 1) / 2;
        const int blocks = (ngroups + groups - 1) / groups;
        if (globalCols) k_fast<R, DIR, true><<<blocks, threads, groups * per, st>>>(d, d.list[slot], d.counters + slot, colAlloc, off, slicePairs);
        else k_fast<R, DIR, false><<<blocks, threads, groups * per, st>>>(d, d.list[slot], d.counters + slot, colAlloc, off, slicePairs);
        c->tm.n_launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

// one launch per non-empty read-length bucket (R = 2*(bucket+1) rows per thread) over the list ranges [first[b], counts[b])
template <int DIR>
static int launch_fast_range(swb_ctx* c, const int* first, const int* counts, cudaStream_t st) {
    for (int b = 0; b < SWB_NBUCKETS; ++b) {
        const int n = counts[b], f = first ? first[b] : 0;
        if (n <= f) continue;
        int rc = 0;
        switch (b) {
            case 0: rc = launch_fast_one<2, DIR>(c, b, f, n, st); break;
            case 1: rc = launch_fast_one<4, DIR>(c, b, f, n, st); break;
            case 2: rc = launch_fast_one<6, DIR>(c, b, f, n, st); break;
            case 3: rc = launch_fast_one<8, DIR>(c, b, f, n, st); break;
            case 4: rc = launch_fast_one<10, DIR>(c, b, f, n, st); break;
            case 5: rc = launch_fast_one<12, DIR>(c, b, f, n, st); break;
            case 6: rc = launch_fast_one<14, DIR>(c, b, f, n, st); break;
            case 7: rc = launch_fast_one<16, DIR>(c, b, f, n, st); break;
        }
        if (rc) return rc;
    }
    return 0;
}

#if SWB_FAST_DIR == 0
int swb_launch_fast_range_fwd(swb_ctx* c, const int* first, const int* counts, cudaStream_t st) { return launch_fast_range<0>(c, first, counts, st); }
#else
int swb_launch_fast_range_rev(swb_ctx* c, const int* first, const int* counts, cudaStream_t st) { return launch_fast_range<1>(c, first, counts, st); }
#endif
