// wpattern.cu -- HBM write throughput of the mask-byte store patterns (16 planes of 4096 x 4096 bytes = 268 MB):
//   A  morph-like: a warp owns 30 word columns (960 B per row) and walks down a strip of TR rows (rows 4096 B apart); units ordered
//      columns fastest, then planes, then strips (fk_morph_lab's order)
//   B  a warp owns a full 4096-byte row segment and walks down TR rows (one contiguous run per warp and strip)
//   C  sequential: consecutive warps write consecutive 1 KB chunks (grid-stride)
// each with `delay` dependent ALU iterations between rows (emulates the compute of the real kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/wpattern tools/ubench/wpattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(void *p, uint32_t v)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t spin(uint32_t v, int n)
{
    for (int i = 0; i < n; i++) v = v * 1664525u + 1013904223u;
    return v;
}

template <int TR>
__global__ void __launch_bounds__(128) kA(uint8_t *out, int h, int w, int K, int delay, int order)
{
    const int wcols = (w / 32 + 29) / 30, strips = h / TR;
    const long long unit = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (unit >= (long long)wcols * strips * K) return;
    const int lane = threadIdx.x & 31;
    const int wx = unit % wcols;
    const long long u2 = unit / wcols;
    const int p = order ? (int)(u2 % K) : (int)(u2 / strips), strip = order ? (int)(u2 / K) : (int)(u2 % strips);
    const int c = wx * 30 - 1 + lane;
    const bool owned = lane >= 1 && lane <= 30 && c >= 0 && c < w / 32;
    uint8_t *row = out + (size_t)p * h * w + (size_t)strip * TR * w + 32 * c;
    uint32_t v = lane;
    for (int r = 0; r < TR; r++) {
        v = spin(v, delay);
        if (owned) st256(row, v | 1u);
        row += w;
    }
}

template <int TR>
__global__ void __launch_bounds__(128) kB(uint8_t *out, int h, int w, int K, int delay)
{
    const int strips = h / TR;
    const long long unit = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (unit >= (long long)strips * K) return;
    const int lane = threadIdx.x & 31;
    const int p = (int)(unit % K), strip = (int)(unit / K);
    uint8_t *row = out + (size_t)p * h * w + (size_t)strip * TR * w;
    uint32_t v = lane;
    for (int r = 0; r < TR; r++) {
        v = spin(v, delay);
        for (int q = 0; q < w / 1024; q++) st256(row + 1024 * q + 32 * lane, v | 1u);
        row += w;
    }
}

__global__ void __launch_bounds__(1024) kC(uint8_t *out, size_t n, int delay)
{
    const size_t warps = (size_t)gridDim.x * blockDim.x / 32, wi = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32;
    const int lane = threadIdx.x & 31;
    uint32_t v = lane;
    for (size_t o = wi * 1024; o < n; o += warps * 1024) {
        v = spin(v, delay);
        st256(out + o + 32 * lane, v | 1u);
    }
}

int main()
{
    const int h = 4096, w = 4096, K = 16;
    const size_t n = (size_t)K * h * w;
    uint8_t *d, *flush;
    cudaMalloc(&d, n); cudaMalloc(&flush, 384u << 20);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, auto launch) {
        float best = 1e9f, sum = 0.f;
        for (int it = 0; it < 7; it++) {
            cudaMemset(flush, it, 384u << 20);
            cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%-44s best %.1f us  mean %.1f us  %.0f GB/s (best)  err=%s\n", name, best * 1e3, sum / 5 * 1e3, n / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int delay : {0, 100, 300}) {
        char nm[96];
        const int wcols = 5;
        snprintf(nm, 96, "A TR=64 planes-fastest delay=%d", delay);
        run(nm, [&] { kA<64><<<(wcols * (h / 64) * K + 3) / 4, 128>>>(d, h, w, K, delay, 1); });
        snprintf(nm, 96, "A TR=64 strips-fastest delay=%d", delay);
        run(nm, [&] { kA<64><<<(wcols * (h / 64) * K + 3) / 4, 128>>>(d, h, w, K, delay, 0); });
        snprintf(nm, 96, "A TR=16 strips-fastest delay=%d", delay);
        run(nm, [&] { kA<16><<<(wcols * (h / 16) * K + 3) / 4, 128>>>(d, h, w, K, delay, 0); });
        snprintf(nm, 96, "B TR=64 full rows delay=%d", delay);
        run(nm, [&] { kB<64><<<((h / 64) * K + 3) / 4, 128>>>(d, h, w, K, delay); });
        snprintf(nm, 96, "B TR=16 full rows delay=%d", delay);
        run(nm, [&] { kB<16><<<((h / 16) * K + 3) / 4, 128>>>(d, h, w, K, delay); });
        snprintf(nm, 96, "C sequential 1 KB per warp delay=%d", delay);
        run(nm, [&] { kC<<<148, 1024>>>(d, n, delay); });
        snprintf(nm, 96, "C sequential, 148 x 8 CTAs of 256 delay=%d", delay);
        run(nm, [&] { kC<<<148 * 8, 256>>>(d, n, delay); });
    }
    run("cudaMemsetAsync", [&] { cudaMemsetAsync(d, 1, n); });
    return 0;
}
