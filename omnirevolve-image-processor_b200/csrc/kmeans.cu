// kmeans.cu -- OPT-IN replacement for the cv2.kmeans call of stage 02 (02_color_extract.py:39-50).
//
// The reference estimates the K Lab centres with cv2.kmeans (kmeans++ seeding, <= 40 Lloyd iterations, eps 0.5, best of 3
// attempts) on a seeded 200 000-pixel subsample: 0.26-0.29 s on the host per image, hundreds of times the GPU time of everything
// that follows.  cv2.kmeans draws from OpenCV's global RNG, so no other implementation can return its centres bit for bit; this
// one keeps the structure (same sample -- the caller passes the reference's own indices --, kmeans++ with 3 trials per centre,
// Lloyd to the same stopping rule, best of `attempts` by compactness) and is DETERMINISTIC: the sample's 8-bit Lab values are
// accumulated as integers, the compactness in fixed point, so no result depends on the order of the atomics.
// Contract (tests/test_gpu_kmeans.py): the compactness of the returned centres is within 2 % of cv2.kmeans' on the same sample;
// labels and masks then follow from the usual exact assignment to THESE centres.  Off by default: the drop-in stage script uses
// it only when OMNI_B200_KMEANS=gpu.
#include "fast_device.cuh"

#define KM_MAX_K OMNI_MAX_K
#define KM_THREADS 256
#define KM_SEED_THREADS 1024
#define KM_TRIALS 3                       // candidates per new centre in the seeding (OpenCV's generateCentersPP uses 3, too)
#define KM_FIX 1024.0f                    // compactness in 1 / 1024 units (u64)

struct KmState {                          // device block of one k-means run
    float ctr[KM_MAX_K * 3];              // current centres
    float best[KM_MAX_K * 3];             // centres of the best attempt so far
    unsigned long long best_comp;         // its compactness (fixed point); ~0: none yet
    unsigned int sum[KM_MAX_K * 4];       // per centre: sum L, sum a, sum b, count (integers: order-independent)
    unsigned long long comp;              // compactness of the last assignment (fixed point)
    int converged;
};

__device__ __forceinline__ unsigned long long km_mix(unsigned long long x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double km_uniform(unsigned long long seed, unsigned a, unsigned b, unsigned c)
{
    const unsigned long long r = km_mix(seed ^ km_mix(((unsigned long long)a << 40) ^ ((unsigned long long)b << 20) ^ c));
    return (double)(r >> 11) * (1.0 / 9007199254740992.0);
}

// sample[i] = 8-bit Lab of pixel idx[i] (idx == NULL: pixel i), packed L | a << 8 | b << 16
__global__ void __launch_bounds__(KM_THREADS) fk_km_gather(const u8 *__restrict__ px, int w, size_t pitch, const int *__restrict__ idx, int n,
                                                           const u16 *__restrict__ labtab, u32 *__restrict__ sample)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = idx ? idx[i] : i;
    const int y = p / w, x = p - y * w;
    const u8 *q = px + (size_t)y * pitch + 3 * (size_t)x;
    int L, a, b;
    lab_noclamp(labtab, labtab + 256, q[0], q[1], q[2], L, a, b);
    sample[i] = (u32)L | ((u32)a << 8) | ((u32)b << 16);
}

__device__ __forceinline__ float km_d2(u32 s, float c0, float c1, float c2)
{
    const float d0 = (float)(s & 255u) - c0, d1 = (float)((s >> 8) & 255u) - c1, d2 = (float)((s >> 16) & 255u) - c2;
    return d0 * d0 + d1 * d1 + d2 * d2;
}

// kmeans++ seeding, one CTA: the first centre uniformly at random, every further one the best of KM_TRIALS candidates drawn with
// probability proportional to the squared distance to the nearest centre so far.  dist: n floats of scratch.
__global__ void __launch_bounds__(KM_SEED_THREADS) fk_km_seed(const u32 *__restrict__ sample, int n, int K, unsigned long long seed, int attempt,
                                                               float *__restrict__ dist, KmState *__restrict__ S)
{
    __shared__ double s_thr[KM_SEED_THREADS];          // per-thread sums of dist[] (the coarse level of the cumulative distance)
    __shared__ double s_red[KM_SEED_THREADS];
    __shared__ int s_pick[KM_TRIALS];
    __shared__ double s_pot[KM_TRIALS];
    const int tid = threadIdx.x;
    const int per = (n + KM_SEED_THREADS - 1) / KM_SEED_THREADS, lo = min(n, tid * per), hi = min(n, lo + per);
    auto block_sum = [&](double v) {                   // fixed-shape tree: the same bits on every run
        s_red[tid] = v;
        __syncthreads();
        for (int d = KM_SEED_THREADS / 2; d > 0; d >>= 1) {
            if (tid < d) s_red[tid] += s_red[tid + d];
            __syncthreads();
        }
        const double t = s_red[0];
        __syncthreads();
        return t;
    };
    if (tid == 0) {
        const int first = min(n - 1, (int)(km_uniform(seed, attempt, 0, 0) * n));
        const u32 s = sample[first];
        S->ctr[0] = (float)(s & 255u); S->ctr[1] = (float)((s >> 8) & 255u); S->ctr[2] = (float)((s >> 16) & 255u);
    }
    __syncthreads();
    double mine = 0.0;
    {
        const float c0 = S->ctr[0], c1 = S->ctr[1], c2 = S->ctr[2];
        for (int i = lo; i < hi; i++) { const float d = km_d2(sample[i], c0, c1, c2); dist[i] = d; mine += d; }
    }
    s_thr[tid] = mine;
    double total = block_sum(mine);
    for (int k = 1; k < K; k++) {
        // ---- draw the candidates: position r in the cumulative distance (thread sums first, then inside that thread's range) ----
        if (tid < KM_TRIALS) {
            const double r = km_uniform(seed, attempt, k, tid + 1) * total;
            double acc = 0.0;
            int th = KM_SEED_THREADS - 1;
            for (int i = 0; i < KM_SEED_THREADS; i++) { if (acc + s_thr[i] > r) { th = i; break; } acc += s_thr[i]; }
            const int a = min(n, th * per), b = min(n, a + per);
            int pick = max(a, b - 1);
            for (int i = a; i < b; i++) { acc += dist[i]; if (acc > r) { pick = i; break; } }
            s_pick[tid] = min(max(pick, 0), n - 1);
        }
        __syncthreads();
        // ---- potential of each candidate ----
        for (int t = 0; t < KM_TRIALS; t++) {
            const u32 s = sample[s_pick[t]];
            const float c0 = (float)(s & 255u), c1 = (float)((s >> 8) & 255u), c2 = (float)((s >> 16) & 255u);
            double pot = 0.0;
            for (int i = lo; i < hi; i++) pot += fminf(dist[i], km_d2(sample[i], c0, c1, c2));
            const double p = block_sum(pot);
            if (tid == 0) s_pot[t] = p;
        }
        __syncthreads();
        int bt = 0;
        for (int t = 1; t < KM_TRIALS; t++) if (s_pot[t] < s_pot[bt]) bt = t;
        const u32 s = sample[s_pick[bt]];
        const float c0 = (float)(s & 255u), c1 = (float)((s >> 8) & 255u), c2 = (float)((s >> 16) & 255u);
        if (tid == 0) { S->ctr[3 * k] = c0; S->ctr[3 * k + 1] = c1; S->ctr[3 * k + 2] = c2; }
        mine = 0.0;
        for (int i = lo; i < hi; i++) { const float d = fminf(dist[i], km_d2(sample[i], c0, c1, c2)); dist[i] = d; mine += d; }
        __syncthreads();
        s_thr[tid] = mine;
        total = block_sum(mine);
    }
    if (tid == 0) {
        S->converged = 0;
        S->comp = 0ull;
        for (int i = 0; i < KM_MAX_K * 4; i++) S->sum[i] = 0u;
    }
}

// one Lloyd assignment: nearest centre (first minimum) of every sample, integer sums per centre, fixed-point compactness
__global__ void __launch_bounds__(KM_THREADS) fk_km_assign(const u32 *__restrict__ sample, int n, int K, KmState *__restrict__ S, int force)
{
    __shared__ float s_c[KM_MAX_K * 3];
    __shared__ unsigned int s_sum[KM_MAX_K * 4];
    __shared__ unsigned long long s_comp;
    if (!force && S->converged) return;                              // uniform: the run has stopped, later launches are no-ops
    for (int i = threadIdx.x; i < K * 3; i += blockDim.x) s_c[i] = S->ctr[i];
    for (int i = threadIdx.x; i < K * 4; i += blockDim.x) s_sum[i] = 0u;
    if (threadIdx.x == 0) s_comp = 0ull;
    __syncthreads();
    unsigned long long comp = 0ull;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 s = sample[i];
        float bd = 3.0e38f;
        int best = 0;
        for (int k = 0; k < K; k++) {
            const float d = km_d2(s, s_c[3 * k], s_c[3 * k + 1], s_c[3 * k + 2]);
            if (d < bd) { bd = d; best = k; }
        }
        atomicAdd(&s_sum[4 * best], s & 255u);
        atomicAdd(&s_sum[4 * best + 1], (s >> 8) & 255u);
        atomicAdd(&s_sum[4 * best + 2], (s >> 16) & 255u);
        atomicAdd(&s_sum[4 * best + 3], 1u);
        comp += (unsigned long long)(bd * KM_FIX + 0.5f);
    }
    atomicAdd(&s_comp, comp);
    __syncthreads();
    for (int i = threadIdx.x; i < K * 4; i += blockDim.x)
        if (s_sum[i]) atomicAdd(&S->sum[i], s_sum[i]);
    if (threadIdx.x == 0 && s_comp) atomicAdd(&S->comp, s_comp);
}

// centres <- means; stop when no centre moved by more than eps (cv2's rule) -- an empty cluster keeps its centre
__global__ void fk_km_update(int K, float eps2, KmState *__restrict__ S)
{
    if (threadIdx.x != 0 || S->converged) return;
    float shift = 0.f;
    for (int k = 0; k < K; k++) {
        const unsigned cnt = S->sum[4 * k + 3];
        if (cnt) {
            float d2 = 0.f;
            for (int d = 0; d < 3; d++) {
                const float c = (float)((double)S->sum[4 * k + d] / (double)cnt);
                const float dd = c - S->ctr[3 * k + d];
                d2 += dd * dd;
                S->ctr[3 * k + d] = c;
            }
            shift = fmaxf(shift, d2);
        }
    }
    for (int i = 0; i < KM_MAX_K * 4; i++) S->sum[i] = 0u;
    S->comp = 0ull;
    if (shift <= eps2) S->converged = 1;
}

// after the final (forced) assignment of an attempt: keep the attempt if it is the most compact one so far
__global__ void fk_km_keep_best(int K, KmState *__restrict__ S)
{
    if (threadIdx.x != 0) return;
    if (S->comp < S->best_comp) {
        S->best_comp = S->comp;
        for (int i = 0; i < K * 3; i++) S->best[i] = S->ctr[i];
    }
}

// Body of omni_kmeans_lab (capi.cu checks the arguments).  h_idx: n sample indices (host) or NULL (every pixel, n = h * w).
int kmeans_lab(omni_ctx *ctx, const u8 *d_bgr, int h, int w, size_t pitch, const int *h_idx, int n, int K, int attempts, int max_iter,
               float eps, unsigned long long seed, float *h_centers, double *h_compactness, cudaStream_t st)
{
    (void)h;
    OMNI_CUDA(fast_tables());
    const u16 *labtab = fast_lab_table();
    if (!labtab) { omni_set_error("Lab tables not available"); return OMNI_ERR_CUDA; }
    // slot 6 (edge run lists; idle outside the edge pass): [KmState | sample n u32 | dist n f32 | idx n i32]
    const size_t arr = ((size_t)n * 4 + 255) & ~(size_t)255;
    const size_t o_sample = (sizeof(KmState) + 255) & ~(size_t)255, o_dist = o_sample + arr, o_idx = o_dist + arr;
    int rc = omni_ws_reserve(ctx, 6, o_idx + arr);
    if (rc != OMNI_OK) return rc;
    u8 *base = (u8 *)ctx->ws[6];
    KmState *S = (KmState *)base;
    u32 *sample = (u32 *)(base + o_sample);
    float *dist = (float *)(base + o_dist);
    int *d_idx = nullptr;
    if (h_idx) {
        d_idx = (int *)(base + o_idx);
        OMNI_CUDA(cudaMemcpyAsync(d_idx, h_idx, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    OMNI_CUDA(cudaMemsetAsync(S, 0xff, sizeof(KmState), st));        // best_comp = ~0
    {
        KScope ks(ctx, "kmeans_gather", st);
        fk_km_gather<<<(n + KM_THREADS - 1) / KM_THREADS, KM_THREADS, 0, st>>>(d_bgr, w, pitch, d_idx, n, labtab, sample);
        OMNI_CUDA(cudaGetLastError());
    }
    const int blocks = std::max(1, std::min(persist_blocks(ctx, 4), (n + KM_THREADS - 1) / KM_THREADS));
    for (int a = 0; a < attempts; a++) {
        {
            KScope ks(ctx, "kmeans_seed", st);
            fk_km_seed<<<1, KM_SEED_THREADS, 0, st>>>(sample, n, K, seed, a, dist, S);
            OMNI_CUDA(cudaGetLastError());
        }
        for (int it = 0; it < max_iter; it++) {
            KScope ks(ctx, "kmeans_lloyd", st);
            fk_km_assign<<<blocks, KM_THREADS, 0, st>>>(sample, n, K, S, 0);
            fk_km_update<<<1, 32, 0, st>>>(K, eps * eps, S);
            OMNI_CUDA(cudaGetLastError());
        }
        KScope ks(ctx, "kmeans_lloyd", st);
        fk_km_assign<<<blocks, KM_THREADS, 0, st>>>(sample, n, K, S, 1);          // compactness of the final centres
        fk_km_keep_best<<<1, 32, 0, st>>>(K, S);
        OMNI_CUDA(cudaGetLastError());
        if (a + 1 < attempts) {                                                    // sums / compactness of the forced pass
            OMNI_CUDA(cudaMemsetAsync(S->sum, 0, sizeof(S->sum), st));
            OMNI_CUDA(cudaMemsetAsync(&S->comp, 0, sizeof(S->comp), st));
        }
    }
    KmState hs;
    OMNI_CUDA(cudaMemcpyAsync(&hs, S, sizeof(KmState), cudaMemcpyDeviceToHost, st));
    OMNI_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < 3 * K; i++) h_centers[i] = hs.best[i];
    if (h_compactness) *h_compactness = (double)hs.best_comp / (double)KM_FIX;
    return OMNI_OK;
}
