// tma_test.cu -- minimal cp.async.bulk.tensor.2d load (u32 elements) of a box from a byte image, as resize_tma.cu issues it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_test tools/ubench/tma_test.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap tm, const CUtensorMap *gtm, const uint8_t *gsrc, uint32_t *out, int bw, int br, int c0, int r0, int variant)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        if (variant == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (variant == 10) {                         // barrier only
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        } else if (variant == 11) {                  // 1-D bulk copy of the first box row
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(bw * 4)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem)), "l"(gsrc), "r"((uint32_t)(bw * 4)), "r"(bar) : "memory");
        } else if (variant == 20) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(br * bw * 4)) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                         ::"r"(smem_u32(smem)), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(bar), "r"(c0), "r"(r0), "l"(0x1000000000000000ull) : "memory");
        } else {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(br * bw * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem)), "l"(variant >= 2 ? reinterpret_cast<unsigned long long>(gtm) : reinterpret_cast<unsigned long long>(&tm)), "r"(c0 * (variant == 3 ? 4 : 1)), "r"(r0), "r"(bar) : "memory");
        }
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
    for (int i = threadIdx.x; i < bw * br; i += blockDim.x) out[i] = reinterpret_cast<uint32_t *>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const int bw_arg = argc > 1 ? atoi(argv[1]) : 64, var_arg = argc > 2 ? atoi(argv[2]) : 0;
    const int dt_arg = argc > 3 ? atoi(argv[3]) : 0, l2_arg = argc > 4 ? atoi(argv[4]) : 1, rank_arg = argc > 5 ? atoi(argv[5]) : 2;
    const int sh = 1080, sw = 1920;
    const size_t pitch = 3 * sw;
    std::vector<uint8_t> h(pitch * sh);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 24);
    uint8_t *d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    for (int bw : {bw_arg}) {
        const int br = 33;
        CUtensorMap tm;
        const CUtensorMapDataType dts[3] = {CU_TENSOR_MAP_DATA_TYPE_UINT32, CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_DATA_TYPE_FLOAT32};
        const int esz = dt_arg == 1 ? 1 : 4;
        const cuuint64_t gdim[3] = {(cuuint64_t)(3 * sw / esz), (cuuint64_t)sh, 1};
        const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * sh};
        const cuuint32_t box[3] = {(cuuint32_t)(bw * 4 / esz), (cuuint32_t)br, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tm, dts[dt_arg], rank_arg, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, l2_arg ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("bw=%d encode result %d\n", bw, (int)r);
        uint32_t *o; cudaMalloc(&o, (size_t)bw * br * 4);
        for (int variant = var_arg; variant <= var_arg; variant++) {
            const int c0 = argc > 6 ? atoi(argv[6]) : 37, r0 = 100;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
            CUtensorMap *gtm; cudaMalloc(&gtm, sizeof(tm)); cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
            if (variant == 21) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(1); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = bw * br * 4 + 256;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                e = cudaLaunchKernelEx(&cfg, k, tm, (const CUtensorMap *)gtm, (const uint8_t *)d, o, bw, br, c0, r0, 0);
                printf("  launchEx: %s\n", cudaGetErrorString(e));
            } else
            k<<<1, 256, bw * br * 4 + 256>>>(tm, gtm, d, o, bw, br, c0, r0, variant);
            e = cudaDeviceSynchronize();
            printf("  variant %d: %s\n", variant, cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            std::vector<uint32_t> got((size_t)bw * br);
            cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (int r2 = 0; r2 < br; r2++)
                for (int c = 0; c < bw; c++) {
                    uint32_t want = 0;
                    if (c0 + c < 3 * sw / 4) memcpy(&want, &h[(size_t)(r0 + r2) * pitch + 4 * (c0 + c)], 4);
                    bad += got[(size_t)r2 * bw + c] != want;
                }
            printf("  mismatches %zu\n", bad);
        }
    }
    return 0;
}
