// hostbw.cu -- host <-> device copy throughput of ONE process on ONE GPU for three kinds of pinned host memory, to be run on
// several GPUs at once (tools/ubench/hostbw_all.sh): does the box's aggregate limit depend on how the host buffer is pinned?
//   mode 0: cudaHostAlloc(default)   1: cudaHostAlloc(write-combined) for the upload source   2: 2 MB-aligned malloc +
//   madvise(MADV_HUGEPAGE) + cudaHostRegister (transparent huge pages: fewer IOMMU translations)
//   nvcc -O3 -o tools/ubench/hostbw tools/ubench/hostbw.cu && tools/ubench/hostbw DEVICE MODE SECONDS
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cuda_runtime.h>
#include <sys/mman.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

static void *alloc_host(size_t n, int mode, bool upload_src)
{
    void *p = nullptr;
    if (mode == 2) {
        if (posix_memalign(&p, 2 << 20, n) != 0) return nullptr;
        madvise(p, n, MADV_HUGEPAGE);
        memset(p, 1, n);
        if (cudaHostRegister(p, n, cudaHostRegisterDefault) != cudaSuccess) return nullptr;
        return p;
    }
    unsigned flags = (mode == 1 && upload_src) ? cudaHostAllocWriteCombined : cudaHostAllocDefault;
    if (cudaHostAlloc(&p, n, flags) != cudaSuccess) return nullptr;
    memset(p, 1, n);
    return p;
}

int main(int argc, char **argv)
{
    const int dev = argc > 1 ? atoi(argv[1]) : 0, mode = argc > 2 ? atoi(argv[2]) : 0;
    const double secs = argc > 3 ? atof(argv[3]) : 2.0;
    CK(cudaSetDevice(dev));
    const size_t nin = 48u << 20, nout = 64u << 20;
    void *hin = alloc_host(nin, mode, true), *hout = alloc_host(nout, mode, false);
    if (!hin || !hout) { printf("host allocation failed (mode %d)\n", mode); return 1; }
    void *din, *dout;
    CK(cudaMalloc(&din, nin)); CK(cudaMalloc(&dout, nout));
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    for (int i = 0; i < 3; i++) { CK(cudaMemcpyAsync(din, hin, nin, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(hout, dout, nout, cudaMemcpyDeviceToHost, s2)); }
    CK(cudaDeviceSynchronize());
    auto t0 = std::chrono::steady_clock::now();
    long iters = 0;
    double el = 0;
    do {
        for (int i = 0; i < 4; i++) { CK(cudaMemcpyAsync(din, hin, nin, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(hout, dout, nout, cudaMemcpyDeviceToHost, s2)); }
        CK(cudaDeviceSynchronize());
        iters += 4;
        el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } while (el < secs);
    printf("dev %d mode %d: %.1f GB/s both directions (%.1f up + %.1f down)\n", dev, mode, iters * (double)(nin + nout) / el / 1e9, iters * (double)nin / el / 1e9,
           iters * (double)nout / el / 1e9);
    return 0;
}
