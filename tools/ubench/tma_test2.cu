// tma_test2.cu -- the CUDA programming guide's bulk-tensor sample (libcu++ wrappers), to tell a kernel bug from an environment limit
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
#define SMEM_W 64
#define SMEM_H 32
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out)
{
    __shared__ alignas(128) int smem_buffer[SMEM_H][SMEM_W];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SMEM_H * SMEM_W; i += blockDim.x) out[i] = smem_buffer[i / SMEM_W][i % SMEM_W];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main()
{
    const int GW = 1024, GH = 1024;
    std::vector<int> h((size_t)GW * GH);
    for (size_t i = 0; i < h.size(); i++) h[i] = (int)i;
    int *d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm{};
    cuuint64_t size[2] = {GW, GH};
    cuuint64_t stride[1] = {GW * sizeof(int)};
    cuuint32_t box[2] = {SMEM_W, SMEM_H};
    cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    int *o; cudaMalloc(&o, SMEM_W * SMEM_H * 4);
    kernel<<<1, 128>>>(tm, 64, 32, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<int> got(SMEM_W * SMEM_H);
        cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < SMEM_H; i++) for (int j = 0; j < SMEM_W; j++) bad += got[i * SMEM_W + j] != h[(size_t)(32 + i) * GW + 64 + j];
        printf("mismatches %d\n", bad);
    }
    return 0;
}
