// edges3.cu -- stage 03 after the morphology, edge_kernel_size 3, on bit-planes:
//   GaussianBlur(3x3, sigma 0) -> Sobel -> L1 magnitude -> non-maximum suppression -> double threshold
//   (03_edge_detect.py:33-34; arithmetic SURVEY.md A.2 / A.5).  Output: candidate / strong BIT-planes for the
//   hysteresis kernel, and the edge BYTE planes for the strong set (the final answer except for promoted weak pixels).
//
// One lane owns one 32-pixel word column and streams down a strip of rows; a warp is 32 adjacent word
// columns (1024 pixels).  Everything dense is SIMD-in-register:
//   bits --6-bit LUT--> horizontal (1,2,1) sums, 4 px per 32-bit word (bytes)
//        --rolling rows--> v = (1,2,1)x(1,2,1) bit sum in [0,16]     B = 16 v - (v > 8)   (exact 8.8 blur)
//        --unpack--> 2 px per word (16-bit lanes): vertical Sobel parts, dx, dy, |dx|+|dy|  (VIADD/VIMNMX.16x2)
//        --sign bits--> 32-bit "m > low" mask per lane-row.
// Only the ~7 % of pixels with m > low go through the 32-bit direction test + neighbour compare; their
// m / dx / dy come from a per-warp shared-memory row ring.
//
// Borders.  Blur uses REFLECT_101 on the mask, Sobel uses REPLICATE on the blurred image, m = 0 outside.  Both
// are obtained by patching only the BITS that enter the pipeline: with bit(-1) := bit(1) (REFLECT_101) and
// additionally bit(-2) := bit(0), the blur of position -1 is  b(-2) + 2 b(-1) + b(0) = 2 b(0) + 2 b(1)  = the
// blur of position 0, i.e. B(-1) = B(0) (REPLICATE) falls out of the normal arithmetic; same at the far side
// and for rows.
#include "fast_kernels.cuh"

#include <cuda_fp16.h>

#define E3V_WARPS 2
#define E3V_LSTRIDE 40                    // u16 per lane region (80 bytes: conflict-free uint4 stores)

// Per-warp shared memory.  Every lane stores ITS OWN window of a row (no cross-lane exchange is needed to
// build the rows); the compacted candidate list lets any lane process any lane's candidates.
struct __align__(16) E3WarpSmem {
    u16 m[3][32 * E3V_LSTRIDE];           // ring of magnitude rows (fp16 bit patterns); lane region index i <-> window pixel e = i + 2
    u16 dx[2][1024];                      // dx, dy of the lanes' own pixels, two rows alternating (fp16 bit patterns):
    u16 dy[2][1024];                      // index = 32 * lane + pixel
    u16 list[1024];                       // candidates of one row: (owner lane << 5) | pixel
    u32 cw[32], sw[32];                   // result words of the row being resolved
};

__device__ __forceinline__ u32 e3_range_mask(int start_px, int w)
{
    int lo = max(0, -start_px), hi = min(32, w - start_px);
    if (hi <= lo) return 0u;
    u32 m = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
    return m & ~((1u << lo) - 1u);
}

__device__ __forceinline__ __half2 as_h2(u32 x) { return *reinterpret_cast<__half2 *>(&x); }
__device__ __forceinline__ u32 as_u32(__half2 x) { return *reinterpret_cast<u32 *>(&x); }

__global__ void __launch_bounds__(E3V_WARPS * 32) fk_edges3_simd(const u32 *__restrict__ m2, int ws, size_t plane, int h, int w, int low,
                                                                 int high, int tr, int wcols, u32 *__restrict__ sbits,
                                                                 u32 *__restrict__ cbits, u8 *__restrict__ edges,
                                                                 size_t estride, size_t epitch, int aligned16, int *__restrict__ wl_count,
                                                                 u32 *__restrict__ worklist, int wl_cap)
{
    __shared__ E3WarpSmem sm[E3V_WARPS];
    __shared__ u32 s_lut6[64];
    if (threadIdx.x < 64) {
        u32 x = threadIdx.x, v = 0;
        for (int i = 0; i < 4; i++) v |= (((x >> i) & 1u) + 2u * ((x >> (i + 1)) & 1u) + ((x >> (i + 2)) & 1u)) << (8 * i);
        s_lut6[x] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    E3WarpSmem &S = sm[wid];
    const int ww = (w + 31) >> 5;
    const int gw = blockIdx.x * E3V_WARPS + wid;
    const int wx = gw % wcols, strip = gw / wcols;
    const int k = blockIdx.y;
    const int y0 = strip * tr, y1 = min(h, y0 + tr);
    if (y0 >= h) return;
    const int c = wx * 32 + lane;
    const bool active = c < ww;
    const u32 *src = m2 + (size_t)k * plane;
    const u32 pv = active ? e3_range_mask(32 * c, w) : 0u;
    const int eW = 4 + w - 32 * c;                       // window index of pixel w (the first one right of the image)
    // the window needs bits up to e = 38 and B up to e = 37, so pixel w matters for eW in [5, 38]: besides the lane
    // that owns pixel w-1 this is the previous lane when w % 32 is 1 or 2
    const bool has_rb = eW >= 5 && eW <= 38;
    const bool is_lb = c == 0;
    const int lowc = min(low, 2041), highc = min(high, 2041);
    const __half2 low_h = __floats2half2_rn((float)lowc, (float)lowc);
    const int high_bits = (int)__half_as_ushort(__float2half_rn((float)highc));
    const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f), k2 = __floats2half2_rn(2.f, 2.f);
    const int lreg = lane * E3V_LSTRIDE;                 // this lane's region (u16 units)

    // one pipeline step: consumes bit row t; (hB, hA) = horizontal sums of rows t-1, t-2, hC receives row t;
    // (UB, UA) = blurred rows t-2, t-3 as half2, UC receives row t-1.
    u32 cand_prev = 0u;
    u32 pf_l = 0u, pf_o = 0u, pf_r = 0u;                 // words of the next bit row (software prefetch)
    auto load_row = [&](const int t) {
        // rows -1 and h mirror rows 1 and h-2 (REFLECT_101); rows -2 and h+1 repeat rows 0 and h-1 (see header)
        int rt = -1;
        if (t >= 0 && t < h) rt = t;
        else if (t == -1) rt = min(1, h - 1);
        else if (t == -2) rt = 0;
        else if (t == h) rt = max(h - 2, 0);
        else if (t == h + 1) rt = h - 1;
        pf_l = pf_o = pf_r = 0u;
        if (rt >= 0) {
            const u32 *row = src + (size_t)rt * ws;
            if (c > 0 && c - 1 < ww) pf_l = __ldg(row + c - 1);
            if (active) pf_o = __ldg(row + c);
            if (c + 1 < ww) pf_r = __ldg(row + c + 1);
        }
    };
    auto step = [&](const int t, u32 (&hA)[10], u32 (&hB)[10], u32 (&hC)[10], u32 (&UA)[20], u32 (&UB)[20], u32 (&UC)[20]) {
        // ---- (1) bit row t (prefetched one step ahead), 40-pixel window: index e <-> pixel 32c - 4 + e ---------------
        u32 lo = (pf_l >> 28) | (pf_o << 4), hi = (pf_o >> 28) | (pf_r << 4);
        if (is_lb) {                                   // pixels -1, -2 := pixels 1, 0
            if (w > 1) lo = (lo & ~0xCu) | ((lo >> 2) & 0xCu);
            else lo = (lo & ~0xCu) | ((lo >> 2) & 4u) | ((lo >> 1) & 8u);
        }
        if (has_rb) {                                  // pixels w, w+1 := pixels w-2, w-1
            unsigned long long win = ((unsigned long long)hi << 32) | lo;
            unsigned long long two = w > 1 ? (win >> (eW - 2)) & 3ull : ((win >> (eW - 1)) & 1ull) * 3ull;
            win = (win & ~(3ull << eW)) | (two << eW);
            lo = (u32)win; hi = (u32)(win >> 32);
        }
        load_row(t + 1);
        // ---- (2) horizontal (1,2,1) sums via the 6-bit LUT ---------------------------------------------------------
        hC[0] = s_lut6[(lo << 1) & 63u];
#pragma unroll
        for (int q = 1; q <= 6; q++) hC[q] = s_lut6[(lo >> (4 * q - 1)) & 63u];
        hC[7] = s_lut6[__funnelshift_r(lo, hi, 27) & 63u];
        hC[8] = s_lut6[__funnelshift_r(lo, hi, 31) & 63u];
        hC[9] = s_lut6[(hi >> 3) & 63u];
        // ---- (3) blurred row t-1: bytes -> half2 ---------------------------------------------------------------------
        {
            u32 B[10];
#pragma unroll
            for (int q = 0; q < 10; q++) {
                u32 v = hA[q] + hC[q] + (hB[q] << 1);                      // bytes in [0,16]
                u32 cc = ((v + 0x07070707u) >> 4) & 0x01010101u;          // v > 8
                B[q] = v * 16u - cc;
            }
#pragma unroll
            for (int q = 0; q < 10; q++) {                                 // 0x6400 | b is the half 1024 + b
                UC[2 * q] = as_u32(__hsub2(as_h2(__byte_perm(B[q], 0x64646464u, 0x4140)), k1024));
                UC[2 * q + 1] = as_u32(__hsub2(as_h2(__byte_perm(B[q], 0x64646464u, 0x4342)), k1024));
            }
        }
        // ---- (4) row r = t-2: Sobel, magnitude, candidate mask; rows into shared memory ----------------------------
        const int r = t - 2;
        u32 cand = 0u;
        if (r >= y0 - 1 && r <= y1) {
            const int slot = (r + 3) % 3;
            u16 *mreg = S.m[slot] + lreg;
            if (r >= 0 && r < h) {
                u32 V[20], D[20];
#pragma unroll
                for (int j = 0; j < 20; j++) {
                    __half2 ua = as_h2(UA[j]), ub = as_h2(UB[j]), uc = as_h2(UC[j]);
                    V[j] = as_u32(__hfma2(ub, k2, __hadd2(ua, uc)));      // B(y-1) + 2 B(y) + B(y+1)   (<= 1020, exact)
                    D[j] = as_u32(__hsub2(uc, ua));                       // B(y+1) - B(y-1)
                }
                u32 mb[18], dxb[18], dyb[18];
                u32 sv_prev = __byte_perm(V[0], V[1], 0x5432), sd_prev = __byte_perm(D[0], D[1], 0x5432);
                u32 acc = 0u;
#pragma unroll
                for (int j = 1; j <= 18; j++) {
                    u32 sv = __byte_perm(V[j], V[j + 1], 0x5432), sd = __byte_perm(D[j], D[j + 1], 0x5432);
                    __half2 dx = __hsub2(as_h2(sv), as_h2(sv_prev));                                     // |dx| <= 1020
                    __half2 dy = __hfma2(as_h2(D[j]), k2, __hadd2(as_h2(sd_prev), as_h2(sd)));           // |dy| <= 1020
                    __half2 m = __hadd2(__habs2(dx), __habs2(dy));                                       // <= 2040: exact in fp16
                    dxb[j - 1] = as_u32(dx); dyb[j - 1] = as_u32(dy); mb[j - 1] = as_u32(m);
                    if (j >= 2 && j <= 17) acc = (acc >> 1) | (__hgt2_mask(m, low_h) & 0x80008000u);
                    sv_prev = sv; sd_prev = sd;
                }
                // acc: bit i = even pixel 2i, bit 16+i = odd pixel 2i+1  ->  interleave
                u32 ev = acc & 0xffffu, od = acc >> 16;
                ev = (ev | (ev << 8)) & 0x00ff00ffu; ev = (ev | (ev << 4)) & 0x0f0f0f0fu;
                ev = (ev | (ev << 2)) & 0x33333333u; ev = (ev | (ev << 1)) & 0x55555555u;
                od = (od | (od << 8)) & 0x00ff00ffu; od = (od | (od << 4)) & 0x0f0f0f0fu;
                od = (od | (od << 2)) & 0x33333333u; od = (od | (od << 1)) & 0x55555555u;
                cand = (ev | (od << 1)) & pv;
                uint4 *mv = reinterpret_cast<uint4 *>(mreg);
#pragma unroll
                for (int jj = 0; jj < 4; jj++) mv[jj] = make_uint4(mb[4 * jj], mb[4 * jj + 1], mb[4 * jj + 2], mb[4 * jj + 3]);
                *reinterpret_cast<uint2 *>(mreg + 32) = make_uint2(mb[16], mb[17]);
                uint4 *dxv = reinterpret_cast<uint4 *>(S.dx[r & 1] + 32 * lane), *dyv = reinterpret_cast<uint4 *>(S.dy[r & 1] + 32 * lane);
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    dxv[jj] = make_uint4(dxb[1 + 4 * jj], dxb[2 + 4 * jj], dxb[3 + 4 * jj], dxb[4 + 4 * jj]);
                    dyv[jj] = make_uint4(dyb[1 + 4 * jj], dyb[2 + 4 * jj], dyb[3 + 4 * jj], dyb[4 + 4 * jj]);
                }
                // the magnitude is 0 outside the image (only the pixels next to image pixels matter)
                if (is_lb) mreg[1] = 0;                                   // pixel -1 (window e = 3)
                if (has_rb && eW <= 37) mreg[eW - 2] = 0;                 // pixel w
            } else {
                uint4 *mv = reinterpret_cast<uint4 *>(mreg);
#pragma unroll
                for (int jj = 0; jj < 4; jj++) mv[jj] = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint2 *>(mreg + 32) = make_uint2(0u, 0u);
            }
        }
        // ---- (5) NMS + thresholds for row t-3: candidates compacted over the warp, one per lane and round -----------
        const int rn = t - 3;
        const bool do_nms = rn >= y0 && rn < y1;
        int total = 0;
        if (do_nms) {
            u32 mk = cand_prev;
            const int cnt = __popc(mk);
            int x = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            total = __shfl_sync(0xffffffffu, x, 31);
            int pos = x - cnt;
            while (mk) {
                const int e = __ffs(mk) - 1;
                mk &= mk - 1u;
                S.list[pos++] = (u16)((lane << 5) | e);
            }
            S.cw[lane] = 0u; S.sw[lane] = 0u;
        }
        __syncwarp();
        if (do_nms) {
            const u16 *mu = S.m[(rn + 2) % 3], *mc = S.m[rn % 3], *md = S.m[(rn + 1) % 3];
            const u16 *dxr = S.dx[rn & 1], *dyr = S.dy[rn & 1];
            for (int i = lane; i < total; i += 32) {
                const int item = S.list[i];
                const int o = item >> 5, e = item & 31;
                const int bi = o * E3V_LSTRIDE + e;
                const int m0 = mc[bi + 2];
                const int dx = __half2int_rn(__ushort_as_half(dxr[item])), dy = __half2int_rn(__ushort_as_half(dyr[item]));
                const int ax = abs(dx), ay = abs(dy) << 15, tg22 = ax * 13573;
                const bool hz = ay < tg22, vt = ay > tg22 + (ax << 16);
                const int s = (dx ^ dy) < 0 ? -1 : 1;
                const int off = hz ? 1 : (vt ? 0 : s);                    // a = (row above|same)[x - off], b = (row below|same)[x + off]
                const u16 *ra = hz ? mc : mu, *rb = hz ? mc : md;
                const int a = ra[bi + 2 - off], b = rb[bi + 2 + off];
                const bool ok = m0 > a && (m0 > b || ((hz || vt) && m0 == b));
                if (ok) {
                    atomicOr(&S.cw[o], 1u << e);
                    if (m0 > high_bits) atomicOr(&S.sw[o], 1u << e);
                }
            }
        }
        __syncwarp();
        if (do_nms && active) {
            size_t o = (size_t)k * plane + (size_t)rn * ws + c;
            const u32 sw = S.sw[lane];
            const u32 cwv = S.cw[lane];
            cbits[o] = cwv; sbits[o] = sw;
            if (cwv & ~sw) {                                   // a weak candidate: the hysteresis kernel must look at this word
                int i = atomicAdd(wl_count, 1);
                if (i < wl_cap) worklist[i] = (u32)o;
            }
            // edge bytes for E = S; the hysteresis kernel patches the (rare) promoted weak pixels afterwards
            // (arithmetic bit->byte expansion: a 2 KB table would cost the 6th resident CTA per SM)
            store_word_bytes(edges + (size_t)k * estride + (size_t)rn * epitch, 32 * c, w, sw, aligned16 != 0);
        }
        cand_prev = cand;
    };

    u32 hA[10], hB[10], hC[10], UA[20], UB[20], UC[20];
#pragma unroll
    for (int q = 0; q < 10; q++) hA[q] = hB[q] = hC[q] = 0u;
#pragma unroll
    for (int j = 0; j < 20; j++) UA[j] = UB[j] = UC[j] = 0u;
    const int t_end = y1 + 2;
    load_row(y0 - 3);
    for (int t = y0 - 3; t <= t_end; t += 3) {
        step(t, hA, hB, hC, UA, UB, UC);
        if (t + 1 > t_end) break;
        step(t + 1, hB, hC, hA, UB, UC, UA);
        if (t + 2 > t_end) break;
        step(t + 2, hC, hA, hB, UC, UA, UB);
    }
}

cudaError_t launch_edges3_simd(const u32 *m2, int ws, size_t plane, int h, int w, int K, int low, int high, int sm_count,
                               u32 *sbits, u32 *cbits, u8 *edges, size_t estride, size_t epitch, int aligned16, int *wl_count, u32 *worklist, int wl_cap,
                               cudaStream_t st)
{
    const int ww = (w + 31) >> 5, wcols = (ww + 31) / 32;
    // strip height: enough warps for ~2 waves of the machine (12 resident warps per SM), at most 64 rows
    long long target = (long long)(sm_count > 0 ? sm_count : 148) * 12 * 2;
    long long per_row_strips = (long long)wcols * K;
    int strips = (int)((target + per_row_strips - 1) / per_row_strips);
    if (strips < 1) strips = 1;
    int tr = (h + strips - 1) / strips;
    if (tr < 16) tr = 16;
    if (tr > 64) tr = 64;
    strips = (h + tr - 1) / tr;
    long long warps = (long long)strips * wcols;
    dim3 grid((unsigned)((warps + E3V_WARPS - 1) / E3V_WARPS), K);
    fk_edges3_simd<<<grid, E3V_WARPS * 32, 0, st>>>(m2, ws, plane, h, w, low, high, tr, wcols, sbits, cbits, edges, estride, epitch, aligned16, wl_count, worklist, wl_cap);
    return cudaGetLastError();
}
