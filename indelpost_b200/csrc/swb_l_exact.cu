// swb_l_exact.cu — launches of the exact striped emulation kernels (swb_exact.cuh, swb_exact2.cuh)
#include "swb_host.h"
#include "swb_exact.cuh"
#include "swb_exact2.cuh"

// fewJobsLikely: the list is expected to hold a handful of jobs (overflow verification) but its length is only known on the
// device: launch the dense and the one-job-per-warp schedule of k_exact2 side by side, the kernels pick by the real count
template <int MODE, int DIR>
static int launch_exact(swb_ctx* c, int listSlot, int upperBound, cudaStream_t st = nullptr, bool fewJobsLikely = false) {
    if (upperBound <= 0) return 0;
    if (!st) st = c->stream;
    const SwbDev& d = c->d;
    const int W = MODE ? 8 : 16;
    const int segAlloc = (d.max_rlen + W - 1) / W;
    const int per = exact_smem_per_group(MODE, d.n, d.max_rlen);
    int groups = 128 / W;                                   // groups per block at 128 threads
    while (groups > 32 / W && (size_t)groups * per > (size_t)c->smem_optin) groups /= 2;
    if ((size_t)groups * per > (size_t)c->smem_optin) { c->err = "read too long for the exact kernel's shared-memory profile"; return -1; }
    // packed variant (two SSE2 lanes per thread) whenever its 16-bit lanes cannot saturate; SWB200_OPT bit1 forces the scalar-lane kernel
    const bool packed = !(d.opt & 2) && (MODE == 0 || (long long)d.max_score * d.max_rlen <= 32000);
    if (packed) {
        const int T2 = W / 2;
        const int per2 = exact2_smem_per_group(MODE, d.n, d.max_rlen);
        int g2 = 128 / T2;
        while (g2 > 32 / T2 && (size_t)g2 * per2 > (size_t)c->smem_optin) g2 /= 2;
        if ((size_t)g2 * per2 <= (size_t)c->smem_optin) {
            static std::atomic<bool> attr2[SWB_MAX_DEVICES] = {};          // function attributes are per device: one flag per device, not one per process
            if (!attr2[c->device % SWB_MAX_DEVICES]) { cudaFuncSetAttribute(k_exact2<MODE, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, c->smem_optin); attr2[c->device % SWB_MAX_DEVICES] = true; }
            k_exact2<MODE, DIR><<<(upperBound + g2 - 1) / g2, g2 * T2, (size_t)g2 * per2, st>>>(d, d.list[listSlot], d.counters + listSlot, segAlloc, per2, fewJobsLikely ? 1 : 0);
            c->tm.n_launches++;
            if (fewJobsLikely) {
                const int warps = g2 * T2 / 32;
                k_exact2<MODE, DIR><<<(SWB_EXACT_SPARSE_MAX + warps - 1) / warps, g2 * T2, (size_t)g2 * per2, st>>>(d, d.list[listSlot], d.counters + listSlot, segAlloc, per2, 2);
                c->tm.n_launches++;
            }
            CUDA_TRY(c, cudaGetLastError());
            return stage_check(c, MODE ? (DIR ? "exact2 word rev" : "exact2 word fwd") : (DIR ? "exact2 byte rev" : "exact2 byte fwd"));
        }
    }
    const int threads = groups * W;
    const int blocks = (upperBound + groups - 1) / groups;
    k_exact<MODE, DIR><<<blocks, threads, (size_t)groups * per, st>>>(d, d.list[listSlot], d.counters + listSlot, segAlloc, per);
    c->tm.n_launches++;
    CUDA_TRY(c, cudaGetLastError());
    return stage_check(c, MODE ? (DIR ? "exact word rev" : "exact word fwd") : (DIR ? "exact byte rev" : "exact byte fwd"));
}

int swb_launch_exact(swb_ctx* c, int mode, int dir, int listSlot, int upperBound, cudaStream_t st, bool fewJobsLikely) {
    if (mode == 0) return dir == 0 ? launch_exact<0, 0>(c, listSlot, upperBound, st, fewJobsLikely) : launch_exact<0, 1>(c, listSlot, upperBound, st, fewJobsLikely);
    return dir == 0 ? launch_exact<1, 0>(c, listSlot, upperBound, st, fewJobsLikely) : launch_exact<1, 1>(c, listSlot, upperBound, st, fewJobsLikely);
}

// opt in to large dynamic shared memory for the scalar-lane kernels (k_exact2 does it at its first launch per device)
cudaError_t swb_exact_set_attrs(int smem_optin) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_exact<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_exact<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_exact<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_exact<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
}
