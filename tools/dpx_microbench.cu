// DPX / integer-pipe micro-benchmark for the roofline denominator.
// Measures sustained warp-instructions / clk / SM for the instruction kinds the
// forward Smith-Waterman sweep is built from.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_microbench dpx_microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { uint32_t d; asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 32768;
constexpr int CHAINS = 8;

template <int KIND>
__global__ void __launch_bounds__(256) bench(uint32_t* out, unsigned long long* cyc, uint32_t seed) {
    uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    uint32_t v[CHAINS], w[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { v[c] = seed * (threadIdx.x + 1) + c; w[c] = seed ^ (c * 77u + threadIdx.x); }
    uint32_t k1 = seed | 0x00010001u, k2 = seed * 3u, sel = (seed & 0x7777u) | 0x8080u;
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (KIND == 0) v[c] = __viaddmax_s16x2(v[c], k1, w[c]);
            if (KIND == 1) v[c] = __vimax3_s16x2_relu(v[c], k1, w[c]);
            if (KIND == 2) v[c] = __vimax_s16x2_relu(v[c], w[c]);
            if (KIND == 3) { asm volatile("add.s16x2 %0, %1, %2;" : "=r"(v[c]) : "r"(v[c]), "r"(k1)); }
            if (KIND == 4) v[c] = __byte_perm(v[c], w[c], sel);
            if (KIND == 5) v[c] = __viaddmax_s32(v[c], k1, w[c]);
            if (KIND == 6) v[c] = max((int)v[c] + (int)k1, (int)w[c]);   // plain s32 add+max (IADD3 + IMNMX or VIADDMNMX)
            if (KIND == 7) {                                             // the 7-instruction cell-pair body
                uint32_t s = __byte_perm(k1, k2, w[c] & 0x7777u);
                uint32_t h = __viaddmax_s16x2(v[c], s, w[c]);
                h = __vimax_s16x2_relu(h, k2);
                k2 = vmax2(k2, h);
                uint32_t hg; asm volatile("add.s16x2 %0, %1, %2;" : "=r"(hg) : "r"(h), "r"(k1));
                w[c] = __viaddmax_s16x2(w[c], sel, hg);
                v[c] = __viaddmax_s16x2(v[c] ^ h, sel, hg);
            }
            if (KIND == 8) {                                             // DPX + an FMA-pipe IMAD per DPX (dual-pipe probe)
                v[c] = __viaddmax_s16x2(v[c], k1, w[c]);
                w[c] = w[c] * k1 + k2;
            }
            if (KIND == 10) { v[c] = (v[c] & w[c]) ^ (w[c] >> 1) ; w[c] = (w[c] | v[c]) ^ k1; }         // 2 LOP3-ish + shift
            if (KIND == 11) { v[c] = v[c] + w[c]; w[c] = w[c] + v[c]; }                              // IADD chain
            if (KIND == 14) { v[c] = __viaddmax_s16x2(v[c], k1, w[c]); w[c] = __vimax_s16x2_relu(w[c], v[c]); }   // half + full
            if (KIND == 15) { v[c] = __viaddmax_s16x2(v[c], k1, w[c]); w[c] = w[c] + v[c]; }                      // DPX + IADD
            if (KIND == 16) { v[c] = __viaddmax_s16x2(v[c], k1, w[c]); w[c] = vmax2(w[c], v[c]); w[c] = __vimax_s16x2_relu(w[c], k2); }  // half + 2 full
            if (KIND == 17) { v[c] = vmax2(v[c], w[c]); w[c] = w[c] + v[c]; }                                      // vimnmx + iadd
            if (KIND == 18) { v[c] = max((int)v[c], (int)w[c]); w[c] = w[c] ^ (v[c] + k1); }                        // 32-bit IMNMX + alu
            if (KIND == 19) { bool ph, pl; v[c] = __vibmax_s16x2(v[c], w[c], &ph, &pl); if (ph) w[c] += k1; if (pl) w[c] ^= k2; }
            if (KIND == 20) { v[c] = __viaddmax_s16x2(v[c], k1, w[c]); w[c] = (w[c] & 0xffffu) * k1 + v[c]; }      // DPX + IMAD (dependent)
            if (KIND == 21) { v[c] = __viaddmax_s16x2_relu(v[c], k1, w[c]); }
            if (KIND == 12) { v[c] = __viaddmax_s16x2(v[c], k1, w[c]); if (c == 0) w[0] = __shfl_up_sync(0xffffffffu, w[0], 1); }
            if (KIND == 13) { v[c] = __viaddmax_s16x2(v[c], k1, w[c]); w[c] = __byte_perm(w[c], k2, sel); }
            if (KIND == 9) {                                             // DPX + LDS mix (1 LDS per 6 DPX)
                uint32_t s = sm[(v[c] + c) & 1023];
                v[c] = __viaddmax_s16x2(v[c], s, w[c]);
                v[c] = __vimax_s16x2_relu(v[c], k2);
                w[c] = __viaddmax_s16x2(w[c], sel, v[c]);
                v[c] = __viaddmax_s16x2(v[c], k1, w[c]);
                w[c] = vmax2(w[c], v[c]);
                v[c] = __viaddmax_s16x2(v[c], sel, w[c]);
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = k2;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= v[c] ^ w[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) { cyc[3 * blockIdx.x] = smid; cyc[3 * blockIdx.x + 1] = t0; cyc[3 * blockIdx.x + 2] = t1; }
}

struct Case { const char* name; int kind; int instr_per_chain; };

template <int KIND>
int run(const char* name, int instr_per_chain, int blocks_per_sm, uint32_t* out, unsigned long long* cyc, int nsm) {
    int blocks = nsm * blocks_per_sm;
    bench<KIND><<<blocks, 256>>>(out, cyc, 12345u);   // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<KIND><<<blocks, 256>>>(out, cyc, 12345u);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long* h = new unsigned long long[3 * blocks];
    CK(cudaMemcpy(h, cyc, 3 * blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long lo[256], hi[256]; int cnt[256];
    for (int i = 0; i < 256; ++i) { lo[i] = ~0ull; hi[i] = 0; cnt[i] = 0; }
    for (int i = 0; i < blocks; ++i) { int s = (int)h[3*i]; if (h[3*i+1] < lo[s]) lo[s] = h[3*i+1]; if (h[3*i+2] > hi[s]) hi[s] = h[3*i+2]; cnt[s]++; }
    double avg = 0, rate = 0; int ns = 0;
    for (int s = 0; s < 256; ++s) if (cnt[s]) { double span = (double)(hi[s] - lo[s]); avg += span; rate += (double)cnt[s] * 8 * ITERS * CHAINS * instr_per_chain / span; ns++; }
    avg /= ns; rate /= ns; delete[] h;
    double warp_instr_per_sm = (double)blocks_per_sm * 8 /*warps*/ * ITERS * CHAINS * instr_per_chain;
    double lanes = warp_instr_per_sm * 32 * nsm;
    printf("{\"case\":\"%s\",\"blocks_per_sm\":%d,\"ms\":%.4f,\"avg_cycles\":%.0f,\"warp_instr_per_clk_per_sm\":%.3f,\"lane_ops_per_clk_per_sm\":%.2f,\"Tlane_ops_per_s\":%.3f,\"implied_mhz\":%.0f}\n",
           name, blocks_per_sm, ms, avg, rate, rate * 32, lanes / (ms * 1e-3) / 1e12, avg / (ms * 1e-3) / 1e6);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", p.name, nsm, p.clockRate);
    uint32_t* out; unsigned long long* cyc;
    CK(cudaMalloc(&out, (size_t)nsm * 8 * 256 * 4)); CK(cudaMalloc(&cyc, (size_t)nsm * 16 * 3 * 8));
    for (int bps : {4}) {
        run<0>("viaddmax_s16x2", 1, bps, out, cyc, nsm);
        run<1>("vimax3_s16x2_relu", 1, bps, out, cyc, nsm);
        run<2>("vimax_s16x2_relu", 1, bps, out, cyc, nsm);
        run<3>("add_s16x2", 1, bps, out, cyc, nsm);
        run<4>("prmt", 1, bps, out, cyc, nsm);
        run<5>("viaddmax_s32", 1, bps, out, cyc, nsm);
        run<6>("add_max_s32_plain", 1, bps, out, cyc, nsm);
        run<7>("cellpair_body_7instr", 7, bps, out, cyc, nsm);
        run<8>("viaddmax_s16x2+imad", 2, bps, out, cyc, nsm);
        run<9>("6dpx+1lds", 7, bps, out, cyc, nsm);
        run<10>("lop3x2+shf", 3, bps, out, cyc, nsm);
        run<11>("iadd_x2", 2, bps, out, cyc, nsm);
        run<14>("viaddmax+vimax_relu", 2, bps, out, cyc, nsm);
        run<15>("viaddmax+iadd", 2, bps, out, cyc, nsm);
        run<16>("viaddmax+2x_vimnmx2", 3, bps, out, cyc, nsm);
        run<17>("vimnmx2+iadd", 2, bps, out, cyc, nsm);
        run<18>("imnmx32+lop+iadd", 3, bps, out, cyc, nsm);
        run<19>("vibmax+2pred", 3, bps, out, cyc, nsm);
        run<20>("viaddmax+imad_dep", 2, bps, out, cyc, nsm);
        run<21>("viaddmax_s16x2_relu", 1, bps, out, cyc, nsm);
        run<12>("8dpx+1shfl(count dpx)", 1, bps, out, cyc, nsm);
        run<13>("dpx+prmt", 2, bps, out, cyc, nsm);
    }
    return 0;
}
