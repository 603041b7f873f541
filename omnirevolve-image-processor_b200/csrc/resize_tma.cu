// resize_tma.cu -- stage 01, fractional INTER_AREA (01_resize.py:15-20 with the default max_dimension: 4096 -> 2000 etc.; arithmetic
// SURVEY A.1 (iii)): the separable form OpenCV itself uses, with the source rectangle of a destination tile staged by TMA.
//
// A CTA produces RT_TX x RT_TY destination pixels:
//   1. ONE bulk-tensor copy (cp.async.bulk.tensor.2d, completion on an mbarrier) brings the source rectangle the tile touches into
//      shared memory -- no LDG / STS instructions, rows and columns past the image edge arrive as zeros and are never fetched.
//      The image is described to the TMA unit as a 2-D tensor of 32-bit words (3 * sw / 4 words per row), so a box of up to 1 KB
//      per row is one copy.  (Measured on B200: the byte offset of the box's first column must be a multiple of 16 -- an odd
//      inner coordinate raises "illegal instruction", tools/ubench/tma_test.cu -- so the box starts at the 16-byte boundary below
//      the tile's first source byte.)
//   2. a thread owns a destination column of the tile.  Horizontal sums: for a source row r the float32 sum  sum_k p[k] * alpha[k]
//      in OpenCV's order (k ascending, products rounded, the first term is the bare product).  The taps are 3 * nt consecutive
//      source BYTES -- fetched as aligned words, re-aligned with funnel shifts, and turned into float32 by one PRMT (byte ->
//      mantissa of 2^23 + b) and one FFMA per byte: fl((2^23 + b) * a - 2^23 * a) = fl(b * a) exactly (2^23 * a is a power-of-two
//      multiple of a, hence exact), i.e. the conversion and the rounded product of the reference in one instruction.  The sum of
//      a source row stays in registers when the next destination row uses the row again (the fractional rows).
//   3. vertical sums: out = rint(sum_j beta[j] * H[r_j]) with the first term the bare product, as in the generic kernel;
//      bytes are collected in shared memory and leave as 16-byte row segments.
// Results are bit-identical to fk_resize_frac / k_resize_frac (tests/test_gpu_parity.py::test_resize_area runs all families).
#include "fast_kernels.cuh"

#include <cuda.h>
#include <math.h>

#define RT_TX 128                         // destination columns per CTA = threads per CTA (a thread owns a column of the tile)
#define RT_TY 16                          // destination rows per CTA
#define RT_THREADS RT_TX
#define RT_MAXTAPS 8

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

template <int MT>                         // taps per destination pixel and axis held in registers (>= floor(scale) + 2)
__global__ void __launch_bounds__(RT_THREADS) fk_resize_tma(const __grid_constant__ CUtensorMap tm, u8 *__restrict__ dst, int dh, int dw,
                                                            size_t dpitch, const ResizeTabDev t, int bw /* box: words per row */,
                                                            int br /* box: rows */, int out_vec)
{
    constexpr int NB = 3 * MT;                       // tap bytes per destination column
    constexpr int NA = (NB + 3) / 4;                 // words that hold them once they start at byte 0
    constexpr int NWL = ((NB + 2) >> 2) + 1;         // staged words that can hold them at any byte phase
    extern __shared__ u8 smem_raw[];
    u8 *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);     // the bulk-tensor copy wants a 128-byte aligned target
    u8 *s_src = smem;                                                        // [br][4 * bw] (+ 32 bytes: taps a lane does not have)
    u8 *s_out = smem + (((size_t)br * bw * 4 + 32 + 127) & ~(size_t)127);    // [RT_TY][3 * RT_TX]
    int *s_yr = reinterpret_cast<int *>(s_out + RT_TY * 3 * RT_TX);          // [RT_TY][MT]: staged row of tap j
    float *s_yb = reinterpret_cast<float *>(s_yr + RT_TY * MT);              // [RT_TY][MT]: its weight
    int *s_yn = reinterpret_cast<int *>(s_yb + RT_TY * MT);                  // [RT_TY]: taps of the row
    __shared__ __align__(8) unsigned long long s_bar;

    const int x0 = blockIdx.x * RT_TX, y0 = blockIdx.y * RT_TY;
    const int x1 = min(dw, x0 + RT_TX), y1 = min(dh, y0 + RT_TY);
    const int xs0 = t.xsi[t.xofs[x0]];
    const int ys0 = t.ysi[t.yofs[y0]];
    const int c0 = ((3 * xs0) >> 2) & ~3;                                    // first staged word of a row: the box must start on a 16-byte
                                                                             // boundary of the row (a misaligned inner coordinate faults)
    const u32 bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((u32)(br * bw * 4)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(s_src)), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(c0), "r"(ys0), "r"(bar)
                     : "memory");
    }
    // ---- while the copy is in flight: this thread's horizontal taps, the tile's vertical taps ----
    const int lx = threadIdx.x;
    const int x = x0 + lx;
    const bool xin = x < x1;
    int nt = 0, a0 = 0;
    float wgt[MT], wneg[MT];
#pragma unroll
    for (int q = 0; q < MT; q++) wgt[q] = wneg[q] = 0.f;
    if (xin) {
        const int xb = t.xofs[x];
        nt = t.xofs[x + 1] - xb;
        a0 = 3 * t.xsi[xb] - 4 * c0;                                         // byte offset of the first tap inside a staged row
#pragma unroll
        for (int q = 0; q < MT; q++)
            if (q < nt) { wgt[q] = t.xal[xb + q]; wneg[q] = __fmul_rn(-8388608.f, wgt[q]); }
    }
    for (int i = threadIdx.x; i < RT_TY; i += RT_THREADS) {
        const int y = y0 + i;
        int n = 0;
        if (y < y1) {
            const int jb = t.yofs[y];
            n = t.yofs[y + 1] - jb;
            for (int j = 0; j < n && j < MT; j++) { s_yr[i * MT + j] = t.ysi[jb + j] - ys0; s_yb[i * MT + j] = t.yal[jb + j]; }
        }
        s_yn[i] = n;
    }
    __syncthreads();                                     // the barrier is initialised, the vertical taps are staged
    // ---- wait for the rectangle ----
    {
        u32 done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    // ---- this thread's column: per destination row the vertical sum of horizontal sums; the horizontal sum of a source row is
    // kept when the next destination row shares the row (the fractional rows do) ----
    if (xin) {
        const int sh8 = (a0 & 3) * 8;
        const u32 *col = reinterpret_cast<const u32 *>(s_src) + (a0 >> 2);
        int cached = -1;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
        auto hsum = [&](const int r) {
            const u32 *row = col + (size_t)r * bw;
            u32 wv[NWL];
#pragma unroll
            for (int i = 0; i < NWL; i++) wv[i] = row[i];
            u32 al[NA];                                                      // the tap bytes, starting at byte 0 of al[0]
#pragma unroll
            for (int i = 0; i < NA; i++) al[i] = __funnelshift_r(wv[i], i + 1 < NWL ? wv[i + 1] : 0u, sh8);
            float hs[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < MT; q++) {
                // a tap this column does not have carries the weights 0 / -0: its product is +0 and leaves the (non-negative) sum as
                // it is, so no lane needs a test here
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int b = 3 * q + c;
                    // 0x4B0000bb = 2^23 + byte
                    const float m = __uint_as_float(__byte_perm(al[b >> 2], 0x4B000000u, 0x7650 + (b & 3)));
                    const float prod = __fmaf_rn(m, wgt[q], wneg[q]);       // = fl(byte * alpha)
                    hs[c] = q == 0 ? prod : __fadd_rn(hs[c], prod);
                }
            }
            h0 = hs[0]; h1 = hs[1]; h2 = hs[2];
        };
        const int nrow = y1 - y0;
        for (int i = 0; i < nrow; i++) {
            const int n = s_yn[i];
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            auto tap = [&](const int j, const int r, const float beta) {
                if (r != cached) { hsum(r); cached = r; }
                if (j == 0) { s0 = __fmul_rn(beta, h0); s1 = __fmul_rn(beta, h1); s2 = __fmul_rn(beta, h2); }
                else {
                    s0 = __fadd_rn(s0, __fmul_rn(beta, h0)); s1 = __fadd_rn(s1, __fmul_rn(beta, h1)); s2 = __fadd_rn(s2, __fmul_rn(beta, h2));
                }
            };
            if (n <= MT) {
                for (int j = 0; j < n; j++) tap(j, s_yr[i * MT + j], s_yb[i * MT + j]);
            } else {                                                         // more vertical taps than the staged table holds
                const int jb = t.yofs[y0 + i];
                for (int j = 0; j < n; j++) tap(j, t.ysi[jb + j] - ys0, t.yal[jb + j]);
            }
            u8 *o = s_out + i * (3 * RT_TX) + 3 * lx;
            o[0] = (u8)min(255, max(0, __float2int_rn(s0)));
            o[1] = (u8)min(255, max(0, __float2int_rn(s1)));
            o[2] = (u8)min(255, max(0, __float2int_rn(s2)));
        }
    }
    __syncthreads();
    // ---- the tile's rows leave as 16-byte segments ----
    const int rowbytes = 3 * (x1 - x0), rows = y1 - y0;
    if (out_vec && (rowbytes & 15) == 0) {
        const int nv = rowbytes >> 4;
        for (int i = threadIdx.x; i < rows * nv; i += RT_THREADS) {
            const int ry = i / nv, v = i - ry * nv;
            *reinterpret_cast<uint4 *>(dst + (size_t)(y0 + ry) * dpitch + (size_t)3 * x0 + 16 * v) =
                *reinterpret_cast<const uint4 *>(s_out + ry * (3 * RT_TX) + 16 * v);
        }
    } else {
        for (int i = threadIdx.x; i < rows * rowbytes; i += RT_THREADS) {
            const int ry = i / rowbytes, b = i - ry * rowbytes;
            dst[(size_t)(y0 + ry) * dpitch + (size_t)3 * x0 + b] = s_out[ry * (3 * RT_TX) + b];
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// cudaErrorNotSupported: the geometry is outside this kernel (rows not word-addressable for the TMA unit, too many taps, box too
// large) -- the caller takes fk_resize_frac / the generic kernel.
cudaError_t tma_resize_frac(const u8 *src, int sh, int sw, size_t spitch, u8 *dst, int dh, int dw, size_t dpitch, const ResizeTabDev *tab,
                            cudaStream_t st)
{
    const double scx = (double)sw / dw, scy = (double)sh / dh;
    if (((uintptr_t)src & 15) || (spitch & 15) || ((3 * (size_t)sw) & 3)) return cudaErrorNotSupported;
    const int need = (int)floor(scx) + 2;
    if ((int)ceil(scx) + 1 > RT_MAXTAPS || (int)ceil(scy) + 1 > RT_MAXTAPS) return cudaErrorNotSupported;
    const int span_px = (int)(scx * RT_TX) + 3;
    const int bw = ((3 * span_px + 15 + 3) / 4 + 3) & ~3;                      // words per row of the box (a multiple of 16 bytes), with
                                                                               // room for the 16-byte alignment of its first column
    const int br = (int)(scy * RT_TY) + 3;
    if (bw > 256 || br > 256) return cudaErrorNotSupported;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)(3 * (size_t)sw / 4), (cuuint64_t)sh};
    const cuuint64_t gstr[1] = {(cuuint64_t)spitch};
    const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)br};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<u8 *>(src), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorNotSupported;
    const int mt = need <= 3 ? 3 : need <= 4 ? 4 : need <= 6 ? 6 : RT_MAXTAPS;
    const size_t smem = (((size_t)br * bw * 4 + 32 + 127) & ~(size_t)127) + (size_t)RT_TY * 3 * RT_TX + (size_t)RT_TY * mt * 8 +
                        RT_TY * sizeof(int) + 128;
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const int out_vec = (((uintptr_t)dst | dpitch) & 15) == 0 && ((3 * RT_TX) & 15) == 0;
    dim3 grid((dw + RT_TX - 1) / RT_TX, (dh + RT_TY - 1) / RT_TY);
#define RT_LAUNCH(MT)                                                                                                             \
    do {                                                                                                                          \
        cudaError_t e = cudaFuncSetAttribute(fk_resize_tma<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);         \
        if (e != cudaSuccess) return e;                                                                                           \
        fk_resize_tma<MT><<<grid, RT_THREADS, smem, st>>>(tm, dst, dh, dw, dpitch, *tab, bw, br, out_vec);                        \
    } while (0)
    if (mt == 3) RT_LAUNCH(3);
    else if (mt == 4) RT_LAUNCH(4);
    else if (mt == 6) RT_LAUNCH(6);
    else RT_LAUNCH(RT_MAXTAPS);
#undef RT_LAUNCH
    return cudaGetLastError();
}
