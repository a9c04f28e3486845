// swb_l_bandreg.cu — launches of the register-band traceback kernels (swb_bandreg.cuh); compiled in two halves of the
// half-width range (-DSWB_BR_HALF=0: W 1-12, 1: W 13-24)
#include "swb_host.h"
#include "swb_bandreg.cuh"

// register-band kernels (swb_bandreg.cuh): one launch per exact half-width, spread over the side streams
template <int W>
static int launch_band_reg_one(swb_ctx* c, int listSlot, int njobs, int nextBase, int nextBaseW, int resume, cudaStream_t st) {
    const SwbDev& d = c->d;
    const int rows = std::min(d.max_rlen, SWB_BANDREG_MAXROWS);
    const size_t smem = (size_t)bandreg_stride_words(rows) * 4 * SWB_BANDREG_THREADS;
    static std::atomic<bool> attr[SWB_MAX_DEVICES] = {};
    if (!attr[c->device % SWB_MAX_DEVICES]) { cudaFuncSetAttribute(k_band_reg<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin - 4096); attr[c->device % SWB_MAX_DEVICES] = true; }
    k_band_reg<W><<<(njobs + SWB_BANDREG_THREADS - 1) / SWB_BANDREG_THREADS, SWB_BANDREG_THREADS, smem, st>>>(d, d.list[listSlot], njobs, nextBase, nextBaseW, resume, rows);
    c->tm.n_launches++;
    return 0;
}
#define SWB_BR_CASE(W) case W: return launch_band_reg_one<W>(c, listSlot, njobs, nextBase, nextBaseW, resume, st);
#if SWB_BR_HALF == 0
int swb_launch_band_reg_lo(swb_ctx* c, int w, int listSlot, int njobs, int nextBase, int nextBaseW, int resume, cudaStream_t st) {
    switch (w) {
        SWB_BR_CASE(1) SWB_BR_CASE(2) SWB_BR_CASE(3) SWB_BR_CASE(4) SWB_BR_CASE(5) SWB_BR_CASE(6) SWB_BR_CASE(7) SWB_BR_CASE(8)
        SWB_BR_CASE(9) SWB_BR_CASE(10) SWB_BR_CASE(11) SWB_BR_CASE(12)
    }
    return -1;
}
#else
int swb_launch_band_reg_hi(swb_ctx* c, int w, int listSlot, int njobs, int nextBase, int nextBaseW, int resume, cudaStream_t st) {
    switch (w) {
        SWB_BR_CASE(13) SWB_BR_CASE(14) SWB_BR_CASE(15) SWB_BR_CASE(16) SWB_BR_CASE(17) SWB_BR_CASE(18) SWB_BR_CASE(19) SWB_BR_CASE(20)
        SWB_BR_CASE(21) SWB_BR_CASE(22) SWB_BR_CASE(23) SWB_BR_CASE(24)
    }
    return -1;
}
#endif
