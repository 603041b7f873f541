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
};

template <bool SPARSE>
__global__ void __launch_bounds__(E3V_WARPS * 32) fk_edges3_simd(const __grid_constant__ E3Args A)
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
                uint4 *dxv = reinterpret_cast<uint4 *>(S.dx[rr & 1] + 32 * lane), *dyv = reinterpret_cast<uint4 *>(S.dy[rr & 1] + 32 * lane);
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
                const u32 dxb = dxr[item], dyb = dyr[item];
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
            if (edges) store_word_bytes(edges + (size_t)k * estride + (size_t)rn * epitch, 32 * c, w, sw, aligned16 != 0);     // NULL: bit-planes only
        }
        cand_prev = cand;
    };

    u32 hA[10], hB[10], hC[10], UA[20], UB[20], UC[20];
#pragma unroll
    for (int q = 0; q < 10; q++) hA[q] = hB[q] = hC[q] = 0u;
#pragma unroll
    for (int j = 0; j < 20; j++) UA[j] = UB[j] = UC[j] = 0u;
    const int tt_end = nrows + 2;
    load_row(y0 - 3);
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
                    if (edges) store_word_bytes(edges + (size_t)k * estride + (size_t)y * epitch, 32 * c, w, 0u, aligned16 != 0);
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
                               cudaStream_t st)
{
    E3Args A{};
    A.m2 = m2; A.ws = ws; A.plane = plane; A.h = h; A.w = w; A.low = low; A.high = high;
    A.sbits = sbits; A.cbits = cbits; A.edges = edges; A.estride = estride; A.epitch = epitch; A.aligned16 = aligned16;
    A.wl_count = wl_count; A.worklist = worklist; A.wl_cap = wl_cap;
    dim3 grid;
    e3_dense_grid(h, (w + 31) >> 5, K, sm_count, &A.tr, &A.wcols, &grid);
    fk_edges3_simd<false><<<grid, E3V_WARPS * 32, 0, st>>>(A);
    return cudaGetLastError();
}

// Longest run, in tiles: long runs pay the 6 extra pipeline rows less often, short runs give more warp items.  On
// pipeline-like masks about a third of the tiles is live and a run holds (1, 1.6, 2, 2.4) tiles on average for
// maxt = 1..4; take the longest runs that still give every resident warp of the edge kernel an item.
int edges3_pick_maxt(int h, int w, int K, int resident_warps)
{
    const double live = 0.35 * (double)K * ((h + ET_R - 1) / ET_R) * ((w + 31) / 32);
    const double avg[ET_MAXT] = {1.0, 1.6, 2.0, 2.4};
    int maxt = ET_MAXT;
    while (maxt > 1 && live / avg[maxt - 1] / 32.0 < (double)resident_warps) maxt--;
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
                                 u32 *worklist, int wl_cap, const int *run_counts, int *run_next, const u32 *run_items, cudaStream_t st)
{
    E3Args A{};
    A.m2 = m2; A.ws = ws; A.plane = plane; A.h = h; A.w = w; A.low = low; A.high = high;
    A.sbits = sbits; A.cbits = cbits; A.edges = edges; A.estride = estride; A.epitch = epitch; A.aligned16 = aligned16;
    A.wl_count = wl_count; A.worklist = worklist; A.wl_cap = wl_cap;
    A.run_counts = run_counts; A.run_next = run_next; A.run_items = run_items;
    edges3_run_words(h, w, K, A.run_off);
    fk_edges3_simd<true><<<grid_blocks, E3V_WARPS * 32, 0, st>>>(A);
    return cudaGetLastError();
}

int edges3_sparse_blocks_per_sm()
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk_edges3_simd<true>, E3V_WARPS * 32, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    return per_sm;
}
