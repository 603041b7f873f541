// edges3.cu -- stage 03 after the morphology, edge_kernel_size 3, on bit-planes:
//   GaussianBlur(3x3, sigma 0) -> Sobel -> L1 magnitude -> non-maximum suppression -> double threshold
//   (03_edge_detect.py:33-34; arithmetic SURVEY.md A.2 / A.5).  Output: candidate / strong BIT-planes for the
//   hysteresis kernel, and the edge BYTE planes for the strong set (the final answer except for promoted weak pixels).
//
// One lane owns one 32-pixel word column and streams down a strip of rows; a warp is 32 adjacent word
// columns (1024 pixels).  Everything dense is SIMD-in-register:
//   bits --6-bit LUT--> horizontal (1,2,1) sums, 4 px per 32-bit word (bytes)
//        --rolling rows--> v = (1,2,1)x(1,2,1) bit sum in [0,16]     B = 16 v - (v > 8)   (exact 8.8 blur)
//        --unpack--> 2 px per word (fp16x2 lanes, exact: all values are integers <= 2040): vertical Sobel parts, dx, dy,
//                    |dx|+|dy|  (HADD2 / HFMA2)
//        --sign bits--> 32-bit "m > low" mask per lane-row.
// Only the ~7 % of pixels with m > low go through the direction test (float32, exact) + neighbour compare; their
// m / dx / dy come from a per-warp shared-memory row ring; candidates are compacted over the warp.
//
// Two variants of one kernel: DENSE walks every word of every plane; SPARSE (the default) walks only the runs of tiles
// that can hold an edge pixel -- the lists come from the morphology kernel (fast_kernels.cu, MorphRuns) or from
// fk_edge_runs below, the zeros of all other tiles from the zero fill that rides on the assignment kernel.
//
// Borders.  Blur uses REFLECT_101 on the mask, Sobel uses REPLICATE on the blurred image, m = 0 outside.  Both
// are obtained by patching only the BITS that enter the pipeline: with bit(-1) := bit(1) (REFLECT_101) and
// additionally bit(-2) := bit(0), the blur of position -1 is  b(-2) + 2 b(-1) + b(0) = 2 b(0) + 2 b(1)  = the
// blur of position 0, i.e. B(-1) = B(0) (REPLICATE) falls out of the normal arithmetic; same at the far side
// and for rows.
#include "fast_kernels.cuh"

#include <cuda_fp16.h>

#define E3V_WARPS 2
#ifndef E3_NMS_UNROLL
#define E3_NMS_UNROLL 2                   // candidates per lane and NMS round (independent dependency chains)
#endif
#define E3V_LSTRIDE 40                    // u16 per lane region (80 bytes: conflict-free uint4 stores)

// Per-warp shared memory.  Every lane stores ITS OWN window of a row (no cross-lane exchange is needed to
// build the rows); the compacted candidate list lets any lane process any lane's candidates.
struct __align__(128) E3WarpSmem {
    u16 m[3][32 * E3V_LSTRIDE];           // ring of magnitude rows (fp16 bit patterns); lane region index i <-> window pixel e = i + 2
    u16 dx[2][1024];                      // dx, dy of the lanes' own pixels, two rows alternating (fp16 bit patterns): index =
    u16 dy[2][1024];                      // 32 * lane + pixel, the 16-byte chunk of the pixel XOR-swizzled with bits 1, 2 of the lane
                                          // (E3_DXY_SWZ): a quarter warp's 16-byte stores then cover all 32 banks
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

// Work items.  DENSE: a warp = 32 adjacent word columns x one strip of `tr` rows of plane blockIdx.y.
// SPARSE: a lane = one RUN from the run lists of fk_edge_runs (below): word column c of plane k, rows
// [8 j0, 8 j0 + 8 nt) -- only runs whose 5x5 neighbourhoods are not uniform can hold an edge pixel; the 32 lanes of a
// warp take 32 runs of the same length nt (lists are bucketed by nt) and walk them in lock step.  Everything in the
// row pipeline is lane-private (own 40-pixel window, own shared-memory regions), so the lanes of a warp may sit
// anywhere in the image; only the candidate list of the NMS step is shared, to balance the per-pixel work.
struct E3Args {
    const u32 *m2; int ws; size_t plane; int h, w, low, high;
    int tr, wcols;                                    // dense: strip height, warp columns
    u32 *sbits, *cbits; u8 *edges; size_t estride, epitch; int aligned16;
    int *wl_count; u32 *worklist; int wl_cap;         // words with weak candidates, for the hysteresis kernel
    const int *run_counts; int *run_next; const u32 *run_items; unsigned run_off[ET_MAXT];   // sparse: run lists
    const u8 *blur; size_t bstride, bpitch;           // BYTES: the blurred planes (edge_kernel_size 5 / 7, fk_blur_bits), rows 4-byte aligned
};

// BYTES = false: blur 3 from the bit-plane m2 (steps 1-3 below).  BYTES = true: the blurred image comes as u8 planes (any blur
// size; Sobel's BORDER_REPLICATE = clamped coordinates), steps 1-3 are a 40-byte row load.
template <bool SPARSE, bool BYTES>
__global__ void __launch_bounds__(E3V_WARPS * 32, 6) fk_edges3_simd(const __grid_constant__ E3Args A)
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
    const u32 *__restrict__ m2 = A.m2;
    u32 *__restrict__ sbits = A.sbits, *__restrict__ cbits = A.cbits;
    u8 *__restrict__ edges = A.edges;
    const int ws = A.ws, h = A.h, w = A.w;
    const size_t plane = A.plane, estride = A.estride, epitch = A.epitch;
    const int aligned16 = A.aligned16, wl_cap = A.wl_cap;
    int *__restrict__ wl_count = A.wl_count;
    u32 *__restrict__ worklist = A.worklist;
    const int ww = (w + 31) >> 5;
    const int lowc = min(A.low, 2041), highc = min(A.high, 2041);
    const __half2 low_h = __floats2half2_rn((float)lowc, (float)lowc);
    const int high_bits = (int)__half_as_ushort(__float2half_rn((float)highc));
    const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f), k2 = __floats2half2_rn(2.f, 2.f);
    const int lreg = lane * E3V_LSTRIDE;                 // this lane's region (u16 units)
#ifndef E3_NO_DXY_SWZ
    const int dxy_off = 64 * lane, dxy_swz = (lane << 3) & 0x30;     // dx / dy rows: byte offset of the lane's region, chunk swizzle
#else
    const int dxy_off = 64 * lane, dxy_swz = 0;
#endif

    // rows [y0, y1) of word column c of plane k (this lane's item); `nrows` is the warp-uniform row count of the item
    // (y1 <= y0 + nrows); lanes without an item run along with `active` false.
    auto process = [&](const int k, const int c, const int y0, const int y1, const int nrows, const bool active) {
    const u32 *src = m2 + (size_t)k * plane;
    const u32 pv = active ? e3_range_mask(32 * c, w) : 0u;
    const int eW = 4 + w - 32 * c;                       // window index of pixel w (the first one right of the image)
    // the window needs bits up to e = 38 and B up to e = 37, so pixel w matters for eW in [5, 38]: besides the lane
    // that owns pixel w-1 this is the previous lane when w % 32 is 1 or 2
    const bool has_rb = eW >= 5 && eW <= 38;
    const bool is_lb = c == 0;

    // one pipeline step: consumes bit row t = y0 + tt; (hB, hA) = horizontal sums of rows t-1, t-2, hC receives row t;
    // (UB, UA) = blurred rows t-2, t-3 as half2, UC receives row t-1.  Ring slots are indexed by the RELATIVE row
    // (the same for every lane of the warp).
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
        if (rt >= 0 && (active || !SPARSE)) {
            const u32 *row = src + (size_t)rt * ws;
            if (c > 0 && c - 1 < ww) pf_l = __ldg(row + c - 1);
            if (active) pf_o = __ldg(row + c);
            if (c + 1 < ww) pf_r = __ldg(row + c + 1);
        }
    };
    u32 pb[BYTES ? 10 : 1];                              // BYTES: the 40 blurred pixels of the next row (window e <-> pixel 32c - 4 + e)
    auto load_brow = [&](const int t1) {
        if (!BYTES) return;
#pragma unroll
        for (int q = 0; q < (BYTES ? 10 : 1); q++) pb[q] = 0u;
        if (!active) return;
        const u8 *row = A.blur + (size_t)k * A.bstride + (size_t)min(max(t1, 0), h - 1) * A.bpitch;
        if (32 * c - 4 >= 0 && 32 * c + 36 <= w) {
            const u32 *p = reinterpret_cast<const u32 *>(row + 32 * c - 4);
#pragma unroll
            for (int q = 0; q < (BYTES ? 10 : 1); q++) pb[q] = __ldg(p + q);
        } else {
#pragma unroll
            for (int q = 0; q < (BYTES ? 10 : 1); q++) {
                u32 v = 0u;
#pragma unroll
                for (int i = 0; i < 4; i++) v |= (u32)row[min(max(32 * c - 4 + 4 * q + i, 0), w - 1)] << (8 * i);
                pb[q] = v;
            }
        }
    };
    auto step = [&](const int tt, u32 (&hA)[10], u32 (&hB)[10], u32 (&hC)[10], u32 (&UA)[20], u32 (&UB)[20], u32 (&UC)[20]) {
        const int t = y0 + tt;
        // ---- (0) candidate list of row t-3 (its candidate mask is known since the previous step): compacted over the warp.
        // Branch-free and first in the step, so that the shuffle chain of the prefix sum and the list stores overlap with the
        // dense arithmetic of (1)-(4) instead of standing alone in front of the NMS.
        const int rn = t - 3, rnr = tt - 3;
        const bool do_nms = rnr >= 0 && rnr < nrows;                        // warp-uniform
        const bool mine = active && rn < y1;                                // this lane has a row to resolve
        int total;
        {
            const u32 mk = (do_nms && mine) ? cand_prev : 0u;
            const int cnt = __popc(mk);
            int x = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            total = __shfl_sync(0xffffffffu, x, 31);
            int pos = x - cnt;
            const u32 tag = (u32)lane << 5;
#pragma unroll
            for (int e = 0; e < 32; e++)                                    // flat and predicated: no find-first-set chain
                if ((mk >> e) & 1u) S.list[pos++] = (u16)(tag | e);
            S.cw[lane] = 0u; S.sw[lane] = 0u;
        }
        if (BYTES) {
            // ---- (1-3) blurred row t-1 (prefetched one step ahead): bytes -> half2 -----------------------------------------
#pragma unroll
            for (int q = 0; q < 10; q++) {
                const u32 bq = pb[BYTES ? q : 0];
                UC[2 * q] = as_u32(__hsub2(as_h2(__byte_perm(bq, 0x64646464u, 0x4140)), k1024));
                UC[2 * q + 1] = as_u32(__hsub2(as_h2(__byte_perm(bq, 0x64646464u, 0x4342)), k1024));
            }
            load_brow(t);
        } else {
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
        }
        // ---- (4) row r = t-2: Sobel, magnitude, candidate mask; rows into shared memory ----------------------------
        const int r = t - 2, rr = tt - 2;                                   // absolute / relative
        u32 cand = 0u;
        if (rr >= -1 && rr <= nrows) {
            const int slot = (rr + 6) % 3;
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
                u8 *dxv = reinterpret_cast<u8 *>(S.dx[rr & 1]) + dxy_off, *dyv = reinterpret_cast<u8 *>(S.dy[rr & 1]) + dxy_off;
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {
                    *reinterpret_cast<uint4 *>(dxv + ((16 * jj) ^ dxy_swz)) = make_uint4(dxb[1 + 4 * jj], dxb[2 + 4 * jj], dxb[3 + 4 * jj], dxb[4 + 4 * jj]);
                    *reinterpret_cast<uint4 *>(dyv + ((16 * jj) ^ dxy_swz)) = make_uint4(dyb[1 + 4 * jj], dyb[2 + 4 * jj], dyb[3 + 4 * jj], dyb[4 + 4 * jj]);
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
        // ---- (5) NMS + thresholds for row t-3: the candidates (compacted over the warp in (0)) are resolved one per lane and round --
        __syncwarp();
        if (do_nms) {
            // the three magnitude rows live in one array: neighbour = centre row + a (uniform) row offset + a column offset
            const u16 *mc = S.m[(rnr + 6) % 3];
            const int dU = (((rnr + 5) % 3) - ((rnr + 6) % 3)) * (32 * E3V_LSTRIDE), dD = (((rnr + 7) % 3) - ((rnr + 6) % 3)) * (32 * E3V_LSTRIDE);
            const u16 *dxr = S.dx[rnr & 1], *dyr = S.dy[rnr & 1];
            // E3_NMS_UNROLL candidates per lane and round (independent dependency chains); a lane whose later indices fall off
            // the list repeats the last entry -- the result bits are OR-ed in, so a repeat is harmless
            auto nms_one = [&](const int item, int &o, u32 &bit, bool &ok, bool &strong) {
                o = item >> 5;
                const int e = item & 31;
                const u16 *pc = mc + (o * E3V_LSTRIDE + e + 2);
                const int m0 = pc[0];
                // cv2.Canny's direction test  |dy| * 2^15 < |dx| * TG22  /  > |dx| * (TG22 + 2^16), TG22 = 13573, in float32:
                // |dx|, |dy| <= 1020 are integers, so |dx| * 13573 < 2^24 and (|dy| - 2|dx|) * 2^15 are exact
#ifndef E3_NO_DXY_SWZ
                const int si = item ^ ((item >> 3) & 0x18);                 // the writer's chunk swizzle (owner lane bits 1, 2)
#else
                const int si = item;
#endif
                const u32 dxb = dxr[si], dyb = dyr[si];
                const float ax = fabsf(__half2float(__ushort_as_half((u16)dxb))), ay = fabsf(__half2float(__ushort_as_half((u16)dyb)));
                const float tg = __fmul_rn(ax, 13573.f);
                const bool hz = __fmul_rn(ay, 32768.f) < tg, vt = __fmul_rn(__fsub_rn(ay, __fadd_rn(ax, ax)), 32768.f) > tg;
                const int s = ((dxb ^ dyb) & 0x8000u) ? -1 : 1;             // fp16 differences are never -0
                const int t = vt ? 0 : s;                                   // a = (row above | same)[x - off], b = (row below | same)[x + off]
                const int a = pc[hz ? -1 : dU - t], b = pc[hz ? 1 : dD + t];
                ok = m0 > a && (m0 > b || ((hz || vt) && m0 == b));
                strong = m0 > high_bits;
                bit = 1u << e;
            };
            for (int i = lane; i < total; i += 32 * E3_NMS_UNROLL) {
                int item[E3_NMS_UNROLL], o[E3_NMS_UNROLL];
                u32 bt[E3_NMS_UNROLL];
                bool ok[E3_NMS_UNROLL], st[E3_NMS_UNROLL];
#pragma unroll
                for (int q = 0; q < E3_NMS_UNROLL; q++) item[q] = S.list[min(i + 32 * q, total - 1)];
#pragma unroll
                for (int q = 0; q < E3_NMS_UNROLL; q++) nms_one(item[q], o[q], bt[q], ok[q], st[q]);
#pragma unroll
                for (int q = 0; q < E3_NMS_UNROLL; q++)
                    if (ok[q]) {
                        atomicOr(&S.cw[o[q]], bt[q]);
                        if (st[q]) atomicOr(&S.sw[o[q]], bt[q]);
                    }
            }
        }
        __syncwarp();
        if (do_nms && mine) {
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
            if (edges) store_word_bytes(edges + (size_t)k * estride + (size_t)rn * epitch, 32 * c, w, sw, aligned16);     // NULL: bit-planes only
        }
        cand_prev = cand;
    };

    u32 hA[10], hB[10], hC[10], UA[20], UB[20], UC[20];
#pragma unroll
    for (int q = 0; q < 10; q++) hA[q] = hB[q] = hC[q] = 0u;
#pragma unroll
    for (int j = 0; j < 20; j++) UA[j] = UB[j] = UC[j] = 0u;
    const int tt_end = nrows + 2;
    if (BYTES) load_brow(y0 - 4);
    else load_row(y0 - 3);
#ifndef E3_ROTATE_THREE
    // one instance of the step (a third of the code: the three rotated copies do not fit the instruction caches); the rolling
    // rows move by register copies instead
    for (int tt = -3; tt <= tt_end; tt++) {
        step(tt, hA, hB, hC, UA, UB, UC);
#pragma unroll
        for (int q = 0; q < 10; q++) { hA[q] = hB[q]; hB[q] = hC[q]; }
#pragma unroll
        for (int j = 0; j < 20; j++) { UA[j] = UB[j]; UB[j] = UC[j]; }
    }
#else
    for (int tt = -3; tt <= tt_end; tt += 3) {
        step(tt, hA, hB, hC, UA, UB, UC);
        if (tt + 1 > tt_end) break;
        step(tt + 1, hB, hC, hA, UB, UC, UA);
        if (tt + 2 > tt_end) break;
        step(tt + 2, hC, hA, hB, UC, UA, UB);
    }
#endif
    };  // process

    if (!SPARSE) {
        const int gw = blockIdx.x * E3V_WARPS + wid;
        const int wx = gw % A.wcols, strip = gw / A.wcols;
        const int y0 = strip * A.tr;
        if (y0 >= h) return;
        const int c = wx * 32 + lane;
        process(blockIdx.y, c, y0, min(h, y0 + A.tr), A.tr, c < ww);
    } else {
        // run lists, longest runs first; a warp item = 32 consecutive entries of ONE list
        int cnt[ET_MAXT], wcum[ET_MAXT + 1];
        wcum[0] = 0;
#pragma unroll
        for (int b = 0; b < ET_MAXT; b++) {
            cnt[b] = A.run_counts[ET_MAXT - 1 - b];                         // b = 0 <-> nt = ET_MAXT
            wcum[b + 1] = wcum[b] + ((cnt[b] + 31) >> 5);
        }
        for (;;) {
            int wi = 0;
            if (lane == 0) wi = atomicAdd(A.run_next, 1);
            wi = __shfl_sync(0xffffffffu, wi, 0);
            if (wi >= wcum[ET_MAXT]) break;
            int b = 0;
#pragma unroll
            for (int q = 1; q < ET_MAXT; q++) if (wi >= wcum[q]) b = q;
            int cb = cnt[0], wb = wcum[0];
#pragma unroll
            for (int q = 1; q < ET_MAXT; q++) if (b == q) { cb = cnt[q]; wb = wcum[q]; }
            const int nt = ET_MAXT - b;
            const int idx = (wi - wb) * 32 + lane;
            const bool valid = idx < cb;
            const u32 item = valid ? __ldg(A.run_items + A.run_off[nt - 1] + idx) : 0u;
            const int k = (int)(item >> 27), j0 = (int)((item >> 13) & 0x3fffu), c = (int)(item & 0x1fffu);
            const int y0 = j0 * ET_R, nrows = nt * ET_R;
            __syncwarp();
            process(k, c, y0, min(h, y0 + nrows), nrows, valid);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Run lists for the sparse edge kernel.  A TILE is one word column x ET_R rows of one plane.  It is DEAD when the
// tile grown by 2 pixels on every side (clipped to the image) is all 0 or all 1: every 5x5 neighbourhood of its
// pixels is then uniform, so the blurred image is constant there (REFLECT_101 / REPLICATE only mirror pixels of the
// same neighbourhood), dx = dy = 0 and no pixel can exceed `low` >= 0.  Dead tiles get their zeros (candidate /
// strong words, edge bytes) written right here.  Vertically adjacent live tiles are merged into runs of at most
// ET_MAXT tiles (the 6 extra pipeline rows are paid per run); runs are listed per length so that a warp of the edge
// kernel works on 32 equally long runs.
// One lane walks one word column down a strip of ER_STRIP_TILES tiles; a warp = 32 adjacent word columns.
// ------------------------------------------------------------------------------------------------
#define ER_STRIP_TILES 4
#define ER_WARPS 4
#define ER_MAXITEMS ER_STRIP_TILES          // maxt = 1: every live tile is an item

__global__ void __launch_bounds__(ER_WARPS * 32) fk_edge_runs(const u32 *__restrict__ m2, int ws, size_t plane, int h, int w, int wcols,
                                                               u32 *__restrict__ sbits, u32 *__restrict__ cbits, u8 *__restrict__ edges,
                                                               size_t estride, size_t epitch, int aligned16, int *__restrict__ run_counts,
                                                               u32 *__restrict__ run_items, const __grid_constant__ E3RunOff off, int maxt)
{
    __shared__ u32 s_item[ER_WARPS][ER_MAXITEMS][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * ER_WARPS + wid;
    const int wx = gw % wcols, strip = gw / wcols;
    const int k = blockIdx.y;
    const int ww = (w + 31) >> 5;
    const int tiles_y = (h + ET_R - 1) / ET_R;
    const int j_lo = strip * ER_STRIP_TILES, j_hi = min(tiles_y, j_lo + ER_STRIP_TILES);
    if (j_lo >= tiles_y) return;
    const int c = wx * 32 + lane;
    const bool active = c < ww;
    const u32 *src = m2 + (size_t)k * plane;
    // validity of the 36 window pixels 32c-2 .. 32c+33
    const u32 pv = active ? e3_range_mask(32 * c, w) : 0u;
    const u32 lv = (active && c > 0) ? 0xC0000000u : 0u;                    // pixels 32c-2, 32c-1 = bits 30, 31 of word c-1
    const u32 rv = active ? (e3_range_mask(32 * c + 32, w) & 3u) : 0u;      // pixels 32c+32, 32c+33
    // "has a 1" / "has a 0" over the 36-pixel windows of the ET_R rows of tile row j: all rows (A), the first two (F),
    // the last two (L).  The rows are fetched as one batch of independent loads.  Tile j grown by 2 rows is L(j-1) | A(j) | F(j+1).
    struct Fl { u32 a1, a0, f1, f0, l1, l0; };
    auto row_group = [&](const int j) {
        Fl f = {0u, 0u, 0u, 0u, 0u, 0u};
        if (!active || j < 0 || j >= tiles_y) return f;
        u32 o[ET_R], l[ET_R], r[ET_R];
        bool in[ET_R];
#pragma unroll
        for (int i = 0; i < ET_R; i++) {
            const int y = j * ET_R + i;
            in[i] = y < h;
            const u32 *row = src + (size_t)(in[i] ? y : 0) * ws;
            o[i] = __ldg(row + c);
            l[i] = lv ? __ldg(row + c - 1) : 0u;
            r[i] = rv ? __ldg(row + c + 1) : 0u;
        }
        const int last = min(h, j * ET_R + ET_R) - j * ET_R - 1;             // index of the tile's last row inside the image
#pragma unroll
        for (int i = 0; i < ET_R; i++) {
            if (!in[i]) continue;
            const u32 h1 = (o[i] & pv) | (l[i] & lv) | (r[i] & rv), h0 = (~o[i] & pv) | (~l[i] & lv) | (~r[i] & rv);
            f.a1 |= h1; f.a0 |= h0;
            if (i < 2) { f.f1 |= h1; f.f0 |= h0; }
            if (i >= last - 1) { f.l1 |= h1; f.l0 |= h0; }
        }
        return f;
    };
    int n_items = 0, run = 0, run_j0 = 0;
    u32 lens = 0u;                                                          // 3 bits per emitted item: its run length nt
    auto emit = [&]() { s_item[wid][n_items++][lane] = ((u32)k << 27) | ((u32)run_j0 << 13) | (u32)c; };
    Fl prev = row_group(j_lo - 1), cur = row_group(j_lo);
#pragma unroll
    for (int jj = 0; jj < ER_STRIP_TILES; jj++) {
        const int j = j_lo + jj;
        if (j >= j_hi) break;
        const Fl next = row_group(j + 1);
        const u32 has1 = prev.l1 | cur.a1 | next.f1, has0 = prev.l0 | cur.a0 | next.f0;
        const bool live = has1 != 0u && has0 != 0u;
        if (live) {
            if (run == 0) run_j0 = j;
            if (++run == maxt) { lens |= (u32)run << (3 * n_items); emit(); run = 0; }
        } else {
            if (run) { lens |= (u32)run << (3 * n_items); emit(); run = 0; }
            if (active) {
                const int y1 = min(h, j * ET_R + ET_R);
                for (int y = j * ET_R; y < y1; y++) {
                    const size_t o = (size_t)k * plane + (size_t)y * ws + c;
                    cbits[o] = 0u; sbits[o] = 0u;
                    if (edges) store_word_bytes(edges + (size_t)k * estride + (size_t)y * epitch, 32 * c, w, 0u, aligned16);
                }
            }
        }
        prev = cur; cur = next;
    }
    if (run) { lens |= (u32)run << (3 * n_items); emit(); run = 0; }
    // flush: per run length one warp-aggregated reservation
#pragma unroll
    for (int nt = 1; nt <= ET_MAXT; nt++) {
        int mine = 0;
        for (int i = 0; i < n_items; i++) mine += ((lens >> (3 * i)) & 7u) == (u32)nt;
        int x = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        const int total = __shfl_sync(0xffffffffu, x, 31);
        if (total == 0) continue;
        int base = 0;
        if (lane == 31) base = atomicAdd(run_counts + nt - 1, total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + x - mine;
        for (int i = 0; i < n_items; i++)
            if (((lens >> (3 * i)) & 7u) == (u32)nt) run_items[off.v[nt - 1] + pos++] = s_item[wid][i][lane];
    }
}

size_t edges3_run_words(int h, int w, int K, unsigned off[ET_MAXT])
{
    const size_t tiles = (size_t)K * ((h + ET_R - 1) / ET_R) * ((w + 31) / 32);
    size_t at = 0;
    for (int nt = 1; nt <= ET_MAXT; nt++) {
        off[nt - 1] = (unsigned)at;
        at += tiles / nt + 64;                      // a run of nt tiles uses nt tiles; strip seams only shorten runs
    }
    return at;
}

bool edges3_sparse_ok(int h, int w, int K)
{
    unsigned off[ET_MAXT];
    return K <= 32 && (w + 31) / 32 <= 0x2000 && (h + ET_R - 1) / ET_R <= 0x4000 && edges3_run_words(h, w, K, off) < 0xffffffffull;
}

static void e3_dense_grid(int h, int ww, int K, int sm_count, int *tr_out, int *wcols_out, dim3 *grid)
{
    const int wcols = (ww + 31) / 32;
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
    *tr_out = tr; *wcols_out = wcols;
    *grid = dim3((unsigned)((warps + E3V_WARPS - 1) / E3V_WARPS), K);
}

cudaError_t launch_edges3_simd(const u32 *m2, int ws, size_t plane, int h, int w, int K, int low, int high, int sm_count,
                               u32 *sbits, u32 *cbits, u8 *edges, size_t estride, size_t epitch, int aligned16, int *wl_count, u32 *worklist, int wl_cap,
                               cudaStream_t st, const u8 *blur, size_t bstride, size_t bpitch)
{
    E3Args A{};
    A.m2 = m2; A.ws = ws; A.plane = plane; A.h = h; A.w = w; A.low = low; A.high = high;
    A.sbits = sbits; A.cbits = cbits; A.edges = edges; A.estride = estride; A.epitch = epitch; A.aligned16 = aligned16;
    A.wl_count = wl_count; A.worklist = worklist; A.wl_cap = wl_cap;
    dim3 grid;
    e3_dense_grid(h, (w + 31) >> 5, K, sm_count, &A.tr, &A.wcols, &grid);
    if (blur) {
        A.blur = blur; A.bstride = bstride; A.bpitch = bpitch;
        fk_edges3_simd<false, true><<<grid, E3V_WARPS * 32, 0, st>>>(A);
    } else
        fk_edges3_simd<false, false><<<grid, E3V_WARPS * 32, 0, st>>>(A);
    return cudaGetLastError();
}

// Longest run, in tiles: long runs pay the 6 extra pipeline rows less often, short runs give more warp items.  On
// pipeline-like masks about a third of the tiles is live and a run holds (1, 1.6, 2, 2.4) tiles on average for
// maxt = 1..4; take the longest runs that still give every resident warp of the edge kernel an item.
int edges3_pick_maxt(int h, int w, int K, int resident_warps)
{
    const double live = 0.35 * (double)K * ((h + ET_R - 1) / ET_R) * ((w + 31) / 32);
    const double avg[ET_MAXT] = {1.0, 1.6, 2.0, 2.4};
    // at least ~1.5 warp items per resident warp: with fewer the persistent warps end far apart (4096^2, K=8: runs of <= 3 tiles
    // instead of 4 take the kernel from 127 to 122 us; K=16 has enough items either way)
    int maxt = ET_MAXT;
    while (maxt > 1 && live / avg[maxt - 1] / 32.0 < 1.5 * (double)resident_warps) maxt--;
    return maxt;
}

// run lists (fk_edge_runs; zero-fills the dead tiles).  run_counts: ET_MAXT ints + 1 int "next warp item", zeroed by the caller.
cudaError_t launch_edge_runs(const u32 *m2, int ws, size_t plane, int h, int w, int K, u32 *sbits, u32 *cbits, u8 *edges,
                             size_t estride, size_t epitch, int aligned16, int *run_counts, u32 *run_items, int resident_warps, cudaStream_t st)
{
    const int maxt = edges3_pick_maxt(h, w, K, resident_warps);
    const int ww = (w + 31) >> 5, wcols = (ww + 31) / 32;
    const int tiles_y = (h + ET_R - 1) / ET_R, strips = (tiles_y + ER_STRIP_TILES - 1) / ER_STRIP_TILES;
    E3RunOff off;
    edges3_run_words(h, w, K, off.v);
    dim3 grid((unsigned)(((long long)strips * wcols + ER_WARPS - 1) / ER_WARPS), K);
    fk_edge_runs<<<grid, ER_WARPS * 32, 0, st>>>(m2, ws, plane, h, w, wcols, sbits, cbits, edges, estride, epitch, aligned16, run_counts,
                                                 run_items, off, maxt);
    return cudaGetLastError();
}

cudaError_t launch_edges3_sparse(const u32 *m2, int ws, size_t plane, int h, int w, int K, int low, int high, int grid_blocks,
                                 u32 *sbits, u32 *cbits, u8 *edges, size_t estride, size_t epitch, int aligned16, int *wl_count,
                                 u32 *worklist, int wl_cap, const int *run_counts, int *run_next, const u32 *run_items, cudaStream_t st,
                                 const u8 *blur, size_t bstride, size_t bpitch)
{
    E3Args A{};
    A.m2 = m2; A.ws = ws; A.plane = plane; A.h = h; A.w = w; A.low = low; A.high = high;
    A.sbits = sbits; A.cbits = cbits; A.edges = edges; A.estride = estride; A.epitch = epitch; A.aligned16 = aligned16;
    A.wl_count = wl_count; A.worklist = worklist; A.wl_cap = wl_cap;
    A.run_counts = run_counts; A.run_next = run_next; A.run_items = run_items;
    edges3_run_words(h, w, K, A.run_off);
    if (blur) {
        A.blur = blur; A.bstride = bstride; A.bpitch = bpitch;
        fk_edges3_simd<true, true><<<grid_blocks, E3V_WARPS * 32, 0, st>>>(A);
    } else
        fk_edges3_simd<true, false><<<grid_blocks, E3V_WARPS * 32, 0, st>>>(A);
    return cudaGetLastError();
}

int edges3_sparse_blocks_per_sm()
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk_edges3_simd<true, false>, E3V_WARPS * 32, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return per_sm;
}


// ------------------------------------------------------------------------------------------------
// cv2.GaussianBlur(mask, (k, k), 0) for k = 5, 7 on a {0,255} mask given as a bit-plane (03_edge_detect.py:33):
//   out = (255 * sum_y w_y sum_x w_x b(y, x) + 32768) >> 16,  BORDER_REFLECT_101,  w = 8.8 fixed point weights with sum 256
//   k = 5: w = 16 * (1,4,6,4,1)        -> S5 = sum of the small weights in [0,256],   out = (255 S5 + 128) >> 8 = S5 - (S5 > 128)
//   k = 7: w = 4 * (2,7,14,18,14,7,2)  -> S7 in [0,4096],                             out = (255 S7 + 2048) >> 12
// A thread produces 4 adjacent pixels: per source row one table lookup (4 + k - 1 window bits -> the 4 horizontal sums as bytes),
// the vertical sum in two 16-bit lanes per word.  Output: u8 planes (rows 4-byte aligned) for the BYTES variant of the edge kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int e3_reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

#define BB_ROWS 64                        // output rows per strip (+ KS - 1 rows of ramp)

template <int KS>
__host__ __device__ constexpr u32 blur_w(int i)               // the small integer weights: 16 * (1,4,6,4,1) / 4 * (2,7,14,18,14,7,2) = cv2's
{
    return KS == 5 ? (i == 0 || i == 4 ? 1u : i == 2 ? 6u : 4u) : (i == 0 || i == 6 ? 2u : i == 1 || i == 5 ? 7u : i == 3 ? 18u : 14u);
}

// A lane owns 16 adjacent pixels (4 groups of 4) and walks down a strip of rows; a warp covers 512 adjacent pixels.  Per source row:
// three word loads -> a 32-bit window (pixels 16q - 8 .. 16q + 23), per group one table lookup (4 + KS - 1 window bits -> the 4
// horizontal sums as bytes) and the vertical filter in transposed form (KS - 1 running sums in two 16-bit lanes per word, one
// multiply-add each per row: FMA pipe).  REFLECT_101: rows by index, columns by patching the window of the lanes at the borders.
template <int KS>
__global__ void __launch_bounds__(128) fk_blur_bits(const u32 *__restrict__ m2, int ws, size_t plane, int K, int h, int w,
                                                    u8 *__restrict__ out, size_t ostride, size_t opitch, int wcols)
{
    constexpr int R = KS / 2, NB = 4 + 2 * R;
    __shared__ u32 s_lut[1 << NB];
    for (int i = threadIdx.x; i < (1 << NB); i += blockDim.x) {
        u32 v = 0u;
        for (int px = 0; px < 4; px++) {
            u32 sum = 0u;
#pragma unroll
            for (int d = 0; d < KS; d++) sum += blur_w<KS>(d) * (((u32)i >> (px + d)) & 1u);
            v |= sum << (8 * px);
        }
        s_lut[i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wx = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (wx >= wcols) return;
    const int q = wx * 32 + lane;                              // 16-pixel column
    const int k = blockIdx.z;
    const int y0 = blockIdx.y * BB_ROWS, y1 = min(h, y0 + BB_ROWS);
    const int x16 = 16 * q;
    if (x16 >= w) return;
    const int ww = (w + 31) >> 5, wi = q >> 1;
    const bool edge = x16 - R < 0 || x16 + 16 + R > w;         // the window leaves the image: mirrored columns
    const bool tiny = w <= 2 * R || h <= R;                    // reflections may repeat: bit-by-bit path
    const u32 *pl = m2 + (size_t)k * plane;
    u8 *op = out + (size_t)k * ostride + x16;
    u32 ae[4][KS - 1], ao[4][KS - 1];                          // running sums, even / odd pixels of each group
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int i = 0; i < KS - 1; i++) ae[g][i] = ao[g][i] = 0u;
    for (int t = y0 - R; t < y1 + R; t++) {
        const int tr = tiny ? e3_reflect101(t, h) : (t < 0 ? -t : t >= h ? 2 * (h - 1) - t : t);
        const u32 *row = pl + (size_t)tr * ws;
        u32 win;                                               // bit i = pixel 16q - 8 + i
        if (!(edge && tiny)) {
            const u32 o = __ldg(row + wi);
            if (q & 1) win = (o >> 8) | ((wi + 1 < ww ? __ldg(row + wi + 1) : 0u) << 24);
            else win = ((wi > 0 ? __ldg(row + wi - 1) : 0u) >> 24) | (o << 8);
            if (edge) {                                        // pixel -j := pixel j, pixel w-1+j := pixel w-1-j  (j = 1..R)
                if (x16 == 0) {
#pragma unroll
                    for (int j = 1; j <= R; j++) win = (win & ~(1u << (8 - j))) | (((win >> (8 + j)) & 1u) << (8 - j));
                }
                const int il = w - 1 - x16 + 8;                // window bit of pixel w-1 (>= 8: this lane holds image pixels)
                if (x16 + 16 + R > w) {
#pragma unroll
                    for (int j = 1; j <= R; j++)
                        if (il + j < 32) win = (win & ~(1u << (il + j))) | (((win >> (il - j)) & 1u) << (il + j));
                }
            }
        } else {
            win = 0u;
            for (int i = 8 - R; i < 24 + R; i++) {
                const int x = e3_reflect101(x16 - 8 + i, w);
                win |= ((__ldg(row + (x >> 5)) >> (x & 31)) & 1u) << i;
            }
        }
        u32 ob[4];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const u32 hs = s_lut[(win >> (8 + 4 * g - R)) & ((1u << NB) - 1u)];
            const u32 he = hs & 0x00FF00FFu, ho = (hs >> 8) & 0x00FF00FFu;
            // transposed FIR: a[i] holds the partial sum that still lacks its last KS - 1 - i rows
            const u32 se = ae[g][KS - 2] + blur_w<KS>(KS - 1) * he, so = ao[g][KS - 2] + blur_w<KS>(KS - 1) * ho;
#pragma unroll
            for (int i = KS - 2; i >= 1; i--) {
                ae[g][i] = ae[g][i - 1] + blur_w<KS>(i) * he;
                ao[g][i] = ao[g][i - 1] + blur_w<KS>(i) * ho;
            }
            ae[g][0] = blur_w<KS>(0) * he; ao[g][0] = blur_w<KS>(0) * ho;
            const u32 S0 = se & 0xFFFFu, S2 = se >> 16, S1 = so & 0xFFFFu, S3 = so >> 16;
            u32 b0, b1, b2, b3;
            if (KS == 5) {
                b0 = S0 - (S0 > 128u ? 1u : 0u); b1 = S1 - (S1 > 128u ? 1u : 0u); b2 = S2 - (S2 > 128u ? 1u : 0u); b3 = S3 - (S3 > 128u ? 1u : 0u);
            } else {
                b0 = (255u * S0 + 2048u) >> 12; b1 = (255u * S1 + 2048u) >> 12; b2 = (255u * S2 + 2048u) >> 12; b3 = (255u * S3 + 2048u) >> 12;
            }
            ob[g] = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
        }
        const int y = t - R;
        if (y >= y0 && y < y1) *reinterpret_cast<uint4 *>(op + (size_t)y * opitch) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
    }
}

cudaError_t launch_blur_bits(int ksize, const u32 *m2, int ws, size_t plane, int K, int h, int w, u8 *out, size_t ostride, size_t opitch,
                             int blocks, cudaStream_t st)
{
    (void)blocks;
    const int cols16 = (w + 15) / 16, wcols = (cols16 + 31) / 32;
    dim3 grid((wcols + 3) / 4, (h + BB_ROWS - 1) / BB_ROWS, K);
    if (ksize == 5) fk_blur_bits<5><<<grid, 128, 0, st>>>(m2, ws, plane, K, h, w, out, ostride, opitch, wcols);
    else if (ksize == 7) fk_blur_bits<7><<<grid, 128, 0, st>>>(m2, ws, plane, K, h, w, out, ostride, opitch, wcols);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
