// swb_l_revband.cu — launches of the banded reverse pass (swb_revband.cuh)
#include "swb_host.h"
#include "swb_revband.cuh"

// banded reverse pass (swb_revband.cuh): one launch per band class, each on its own stream so that the classes
// (a few thousand warps each) overlap; grids are sized by an upper bound (the classes were filled by the forward
// sweep, no host round trip), blocks beyond the real count exit at once
int swb_launch_rev_band(swb_ctx* c, int upperBoundPairs) {
    if (upperBoundPairs <= 0 || (c->d.opt & 4)) return 0;
    const SwbDev& d = c->d;
    const int T = SWB_REVB_THREADS;
    const int blocks = ((upperBoundPairs + 1) / 2 + T - 1) / T;
    const int rows = std::min(d.max_rlen, 32 * SWB_NBUCKETS);
    CUDA_TRY(c, cudaEventRecord(c->ev_rev_fork, c->stream));
#define SWB_REVB_LAUNCH(cls, WI, WD) { \
        const size_t smem = (size_t)revb_stride_words(rows, WI + WD + 1) * 4 * T; \
        static std::atomic<bool> attr[SWB_MAX_DEVICES] = {}; \
        if (!attr[c->device % SWB_MAX_DEVICES]) { cudaFuncSetAttribute(k_rev_band<WI, WD>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 4096); attr[c->device % SWB_MAX_DEVICES] = true; } \
        CUDA_TRY(c, cudaStreamWaitEvent(c->rev_stream[cls], c->ev_rev_fork, 0)); \
        k_rev_band<WI, WD><<<blocks, T, smem, c->rev_stream[cls]>>>(d, d.list[LIST_REVB + cls], d.counters + LIST_REVB + cls, rows); \
        c->tm.n_launches++; \
        CUDA_TRY(c, cudaEventRecord(c->ev_rev_join[cls], c->rev_stream[cls])); }
    SWB_REVB_CLASSES(SWB_REVB_LAUNCH)
#undef SWB_REVB_LAUNCH
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}