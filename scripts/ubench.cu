// ubench.cu -- instruction-pipe and shared-memory micro-benchmarks for sm_100a (development tool, not part of the
// library): which pipes the filter's instructions use and what they sustain per SM, alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench scripts/ubench.cu   (here);  ./scripts/ubench (GPU box)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int U = 8; // independent chains

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_pipe(uint32_t *out, uint32_t seed, uint32_t k1, uint32_t k2)
{
    extern __shared__ uint32_t smem_raw[];
    // the table at the first 64 KB-aligned shared address of a 128 KB allocation (what the kernel's PRMT address needs)
    const uint32_t raw_sa = (uint32_t)__cvta_generic_to_shared(smem_raw);
    uint32_t *lut = smem_raw + (((0x10000u - (raw_sa & 0xffffu)) & 0xffffu) >> 2);
    for (uint32_t i = threadIdx.x; i < 256 * 64; i += blockDim.x) lut[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lutlane = (uint32_t)__cvta_generic_to_shared(lut) + lane * 4;
    uint32_t a[U], b[U];
#pragma unroll
    for (int j = 0; j < U; j++) { a[j] = seed + threadIdx.x * 7 + j; b[j] = seed * 3 + j; }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int j = 0; j < U; j++) {
            if (MODE == 0) a[j] = __byte_perm(a[j], k1, 0x7604);                 // PRMT (ALU)
            if (MODE == 1) a[j] = __dp4a(a[j], k1, k2);                          // IDP.4A
            if (MODE == 2) a[j] = a[j] * k1 + k2;                                // IMAD
            if (MODE == 3) a[j] = (a[j] & k1) ^ k2;                              // LOP3
            if (MODE == 4) { a[j] = __byte_perm(a[j], k1, 0x7604); b[j] = b[j] * k1 + k2; }  // PRMT + IMAD
            if (MODE == 5) { a[j] = (a[j] & k1) ^ k2; b[j] = b[j] * k1 + k2; }   // LOP3 + IMAD
            if (MODE == 6) { a[j] = __byte_perm(a[j], k1, 0x7604); b[j] = __dp4a(b[j], k1, k2); } // PRMT + IDP4A
            if (MODE == 7) { // bank-private LDS lookup chain: address from the previous value (latency-bound per chain)
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(lutlane + ((a[j] & 0xffu) << 7)));
                a[j] = v;
            }
            if (MODE == 8) a[j] = __shfl_sync(0xffffffffu, a[j], (lane + 1) & 31) + 1;  // SHFL
            if (MODE == 9) { // LDS lookup + SHFL mixed
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(lutlane + ((a[j] & 0xffu) << 7)));
                a[j] = v;
                b[j] = __shfl_sync(0xffffffffu, b[j], (lane + 1) & 31) + 1;
            }
            if (MODE == 10) { // the filter step as the kernel has it: PRMT address, LDS, IMAD, LOP3
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(__byte_perm(b[j], lutlane, 0x7604)));
                a[j] = (a[j] * k1 + 255u) & v;
                b[j] += a[j];
            }
            if (MODE == 11) { // the same with an IDP.4A address
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(__dp4a(b[j] & 0x7f7f7f7fu, 0x00000080u, lutlane)));
                a[j] = (a[j] * k1 + 255u) & v;
                b[j] += a[j];
            }
            if (MODE == 12) a[j] = __funnelshift_l(a[j], k1, 8) ^ k2;           // SHF + LOP3
            if (MODE == 13) a[j] = __popc(a[j]) + k1;                            // POPC
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < U; j++) s += a[j] + b[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int instr_per_step, int threads)
{
    uint32_t *out;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaFuncSetAttribute(k_pipe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_pipe<MODE><<<148, threads, 128 * 1024>>>(out, 1, 0x01000193u, 0x9e3779b9u);
    cudaEventRecord(e0);
    k_pipe<MODE><<<148, threads, 128 * 1024>>>(out, 1, 0x01000193u, 0x9e3779b9u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double steps = (double)ITER * U * (threads / 32);
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s threads %4d: %.3f ms, %.3f warp-steps/cycle/SM = %.2f cycles per step (x%d listed instr; nominal clock %d MHz) err=%s\n",
           name, threads, ms, steps / cycles, cycles / steps, instr_per_step, clk / 1000, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main()
{
    for (int threads : {1024, 512}) {
        run<0>("PRMT", 1, threads);
        run<1>("IDP.4A", 1, threads);
        run<2>("IMAD", 1, threads);
        run<3>("LOP3", 1, threads);
        run<4>("PRMT+IMAD", 2, threads);
        run<5>("LOP3+IMAD", 2, threads);
        run<6>("PRMT+IDP.4A", 2, threads);
        run<7>("LDS lookup (+LOP/SHF addr)", 1, threads);
        run<8>("SHFL (+IADD)", 1, threads);
        run<9>("LDS lookup + SHFL", 2, threads);
        run<10>("filter step PRMT addr (5+)", 5, threads);
        run<11>("filter step IDP4A addr (6)", 6, threads);
        run<12>("SHF+LOP3", 2, threads);
        run<13>("POPC+IADD", 2, threads);
    }
    return 0;
}
