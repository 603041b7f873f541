// pipes.cu -- issue-throughput micro-benchmark for the instruction mixes the bit-plane kernels are made of (B200, sm_100a).
// Prints warp-instructions per clock per SM for each mix (8 independent chains per thread, 32 warps per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu && tools/ubench/pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
typedef uint32_t u32;
#define ITERS 4096
#define CH 8

template <int MODE>
__global__ void __launch_bounds__(256) k(u32 *out, u32 a0, u32 two, u32 m1)
{
    u32 x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { x[i] = a0 + threadIdx.x * 7 + i; y[i] = a0 * 3 + i + blockIdx.x; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (MODE == 0) { x[i] = (x[i] & y[i]) ^ m1; }                                       // LOP3
            if (MODE == 1) { x[i] = __funnelshift_l(y[i], x[i], 1); }                           // SHF
            if (MODE == 2) { x[i] = x[i] * two + y[i]; }                                        // IMAD
            if (MODE == 3) { unsigned long long p = (unsigned long long)x[i] * two; x[i] = (u32)p ^ (u32)(p >> 32); }  // IMAD.WIDE + LOP
            if (MODE == 4) { x[i] = __byte_perm(x[i], y[i], 0x5432); }                          // PRMT
            if (MODE == 5) { __half2 h = *reinterpret_cast<__half2 *>(&x[i]); __half2 g = *reinterpret_cast<__half2 *>(&y[i]);
                             h = __hfma2(h, g, g); x[i] = *reinterpret_cast<u32 *>(&h); }       // HFMA2
            if (MODE == 6) { x[i] = (x[i] & y[i]) ^ m1; y[i] = y[i] * two + m1; }               // LOP3 + IMAD (two pipes)
            if (MODE == 7) { x[i] = (x[i] & y[i]) ^ m1; y[i] = __funnelshift_l(x[i], y[i], 1); }// LOP3 + SHF (one pipe)
            if (MODE == 8) { x[i] = __ballot_sync(0xffffffffu, x[i] > y[i]) + y[i]; }           // VOTE (+IADD)
            if (MODE == 9) { x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1) + y[i]; }             // SHFL (+IADD)
            if (MODE == 10) { x[i] = __popc(x[i]) + y[i]; }                                      // POPC (+IADD)
            if (MODE == 11) { x[i] = (u32)__ffs(x[i]) + y[i]; }                                  // FLO (+IADD)
            if (MODE == 12) { x[i] = __match_any_sync(0xffffffffu, x[i] & 15u) + y[i]; }         // MATCH
            if (MODE == 13) { x[i] = x[i] + y[i] + m1; }                                         // IADD3
            if (MODE == 14) { __half2 h = *reinterpret_cast<__half2 *>(&x[i]); __half2 g = *reinterpret_cast<__half2 *>(&y[i]);
                              x[i] = __hgt2_mask(h, g) ^ y[i]; }                                 // HSET2 + LOP
            if (MODE == 15) { x[i] = (x[i] << 1) | (y[i] >> 31); }                               // SHL/SHR/LOP as written
            if (MODE == 16) { x[i] = x[i] * two; y[i] = (y[i] & x[i]) | m1; }                    // shift-by-IMAD + LOP3
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= x[i] ^ y[i];
    if (s == 0x12345678u) out[0] = s;
}

template <int MODE> void run(const char *name, double instr_per_iter_chain)
{
    u32 *d; cudaMalloc(&d, 4);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    dim3 g(sms * 4), b(256);
    k<MODE><<<g, b>>>(d, 1, 2, 0x55); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<MODE><<<g, b>>>(d, 1, 2, 0x55); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    double winstr = (double)sms * 4 * 8 * ITERS * CH * instr_per_iter_chain;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (nominal clock %d MHz; counts %g source ops per chain-iteration)\n", name, ms,
           winstr / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, instr_per_iter_chain);
    cudaFree(d);
}

int main()
{
    run<0>("LOP3", 1); run<1>("SHF funnel", 1); run<2>("IMAD", 1); run<3>("IMAD.WIDE+LOP", 2); run<4>("PRMT", 1);
    run<5>("HFMA2", 1); run<6>("LOP3+IMAD", 2); run<7>("LOP3+SHF", 2); run<8>("VOTE+IADD", 2); run<9>("SHFL+IADD", 2);
    run<10>("POPC+IADD", 2); run<11>("FLO+IADD", 2); run<12>("MATCH+IADD", 2); run<13>("IADD3", 1); run<14>("HSET2+LOP", 2);
    run<15>("SHL|SHR (as written)", 1); run<16>("IMAD(x2)+LOP3", 2);
    return 0;
}
