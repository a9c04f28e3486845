// swb_l_band.cu — launches of the literal banded_sw kernel (swb_band.cuh) and the warp-per-alignment band kernel (swb_bandwarp.cuh)
#include "swb_host.h"
#include "swb_band.cuh"
#include "swb_bandwarp.cuh"

int swb_launch_band(swb_ctx* c, int which, int blocks, int listBase, int firstClass, int lastClass, int nextBase, cudaStream_t st) {
    const SwbDev& d = c->d;
    if (blocks <= 0) return 0;
    switch (which) {
        case 0: k_band<0, SWB_BAND_THREADS><<<blocks, SWB_BAND_THREADS, 0, st>>>(d, listBase, firstClass, lastClass, nextBase); break;
        case 1: k_band<SWB_BAND_LOCAL_BW, SWB_BAND_THREADS><<<blocks, SWB_BAND_THREADS, band_smem_bytes(SWB_BAND_LOCAL_BW, SWB_BAND_THREADS), st>>>(d, listBase, firstClass, lastClass, nextBase); break;
        case 2: k_band<SWB_BAND_MID_BW, SWB_BAND_MID_THREADS><<<blocks, SWB_BAND_MID_THREADS, band_smem_bytes(SWB_BAND_MID_BW, SWB_BAND_MID_THREADS), st>>>(d, listBase, firstClass, lastClass, nextBase); break;
        case 3: k_band<SWB_BAND_WIDE_BW, SWB_BAND_WIDE_THREADS><<<blocks, SWB_BAND_WIDE_THREADS, band_smem_bytes(SWB_BAND_WIDE_BW, SWB_BAND_WIDE_THREADS), st>>>(d, listBase, firstClass, lastClass, nextBase); break;
        case 4: k_band<SWB_BAND_HUGE_BW, SWB_BAND_HUGE_THREADS><<<blocks, SWB_BAND_HUGE_THREADS, band_smem_bytes(SWB_BAND_HUGE_BW, SWB_BAND_HUGE_THREADS), st>>>(d, listBase, firstClass, lastClass, nextBase); break;
        default: return -1;
    }
    c->tm.n_launches++;
    return 0;
}

// wide-band jobs: split by schedule on the device (no host round trip: both kernels are launched against the list's length as an
// upper bound and read their real job counts), eight lanes per alignment for half-widths up to SWB_BANDQ_MAXW, a whole warp otherwise
int swb_launch_band_warp(swb_ctx* c, int listSlot, int njobs, int nextBase, cudaStream_t st, cudaStream_t st2, cudaEvent_t evSplit) {
    const SwbDev& d = c->d;
    if (njobs <= 0) return 0;
    CUDA_TRY(c, cudaMemsetAsync(d.counters + LIST_BANDQ_TMP, 0, 8, st));      // LIST_BANDQ_TMP, LIST_BANDW_TMP
    k_bandwarp_split<<<(njobs + 127) / 128, 128, 0, st>>>(d, d.list[listSlot], njobs, LIST_BANDQ_TMP, LIST_BANDW_TMP);
    CUDA_TRY(c, cudaEventRecord(evSplit, st));
    CUDA_TRY(c, cudaStreamWaitEvent(st2, evSplit, 0));
    constexpr int JQ = SWB_BANDWARP_WARPS * 4;
    // the two schedules side by side (st: eight lanes per alignment, st2: a warp per alignment)
    k_band_warp<8><<<(njobs + JQ - 1) / JQ, 32 * SWB_BANDWARP_WARPS, 0, st>>>(d, d.list[LIST_BANDQ_TMP], njobs, d.counters + LIST_BANDQ_TMP, nextBase);
    k_band_warp<32><<<(njobs + SWB_BANDWARP_WARPS - 1) / SWB_BANDWARP_WARPS, 32 * SWB_BANDWARP_WARPS, 0, st2>>>(d, d.list[LIST_BANDW_TMP], njobs, d.counters + LIST_BANDW_TMP, nextBase);
    c->tm.n_launches += 3;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

cudaError_t swb_band_set_attrs() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_band<SWB_BAND_LOCAL_BW, SWB_BAND_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, band_smem_bytes(SWB_BAND_LOCAL_BW, SWB_BAND_THREADS))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_band<SWB_BAND_MID_BW, SWB_BAND_MID_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, band_smem_bytes(SWB_BAND_MID_BW, SWB_BAND_MID_THREADS))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_band<SWB_BAND_WIDE_BW, SWB_BAND_WIDE_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, band_smem_bytes(SWB_BAND_WIDE_BW, SWB_BAND_WIDE_THREADS))) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_band<SWB_BAND_HUGE_BW, SWB_BAND_HUGE_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, band_smem_bytes(SWB_BAND_HUGE_BW, SWB_BAND_HUGE_THREADS));
}
