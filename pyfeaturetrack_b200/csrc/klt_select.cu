// Good-feature selection on the GPU (sm_100a).
//
// Replaces goodFeaturesUtils.ScanImageForGoodFeatures (goodFeaturesUtils.pyx:35-73), the candidate sort
// (selectGoodFeatures.py:234-236) and _enforceMinimumDistance (selectGoodFeatures.py:45-135).
//
// The reference's eigenvalues carry the rounding of three float32 summed-area tables built by strictly
// sequential additions (np.cumsum along rows, then along columns); selection ORDER depends on that rounding
// (SURVEY 7.3), so the tables are rebuilt here with the same chains: one thread per row, then one thread per
// column.  A parallel prefix scan would be faster per element but produces different float32 sums.
#include "klt_common.cuh"

// ---- summed-area tables -------------------------------------------------------------------------------
// rows: s[y][x] = s[y][x-1] + p[y][x], p = exact fp32 product (np.power(g,2.) / g*g, pyx:49-51)
__global__ void __launch_bounds__(32)
sat_rows_kernel(const float *__restrict__ gx, const float *__restrict__ gy, size_t pitch, int W, int H,
                float *__restrict__ sxx, float *__restrict__ sxy, float *__restrict__ syy) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= H) return;
    const float *a = gx + (size_t)y * pitch, *b = gy + (size_t)y * pitch;
    float *oxx = sxx + (size_t)y * W, *oxy = sxy + (size_t)y * W, *oyy = syy + (size_t)y * W;
    float axx = 0.f, axy = 0.f, ayy = 0.f;     // 0 + p == p exactly, so starting from 0 equals cumsum's first copy
    int x = 0;
    for (; x + 8 <= W; x += 8) {
        float va[8], vb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { va[i] = a[x + i]; vb[i] = b[x + i]; }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            axx = __fadd_rn(axx, __fmul_rn(va[i], va[i]));
            axy = __fadd_rn(axy, __fmul_rn(va[i], vb[i]));
            ayy = __fadd_rn(ayy, __fmul_rn(vb[i], vb[i]));
            oxx[x + i] = axx; oxy[x + i] = axy; oyy[x + i] = ayy;
        }
    }
    for (; x < W; x++) {
        const float va = a[x], vb = b[x];
        axx = __fadd_rn(axx, __fmul_rn(va, va));
        axy = __fadd_rn(axy, __fmul_rn(va, vb));
        ayy = __fadd_rn(ayy, __fmul_rn(vb, vb));
        oxx[x] = axx; oxy[x] = axy; oyy[x] = ayy;
    }
}
// columns: s[y][x] = s[y-1][x] + s[y][x]; blockIdx.y selects the table
__global__ void __launch_bounds__(128)
sat_cols_kernel(float *__restrict__ s0, float *__restrict__ s1, float *__restrict__ s2, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    float *s = (blockIdx.y == 0 ? s0 : blockIdx.y == 1 ? s1 : s2) + x;
    float acc = s[0];
    int y = 1;
    for (; y + 8 <= H; y += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = s[(size_t)(y + i) * W];
#pragma unroll
        for (int i = 0; i < 8; i++) { acc = __fadd_rn(acc, v[i]); s[(size_t)(y + i) * W] = acc; }
    }
    for (; y < H; y++) { acc = __fadd_rn(acc, s[(size_t)y * W]); s[(size_t)y * W] = acc; }
}

// four-corner combine + min eigenvalue (pyx:17-31), operand types as in the Cython-generated C
__device__ __forceinline__ float window_sum(const float *__restrict__ s, int W, int x, int y, int hw, int hh) {
    const float a = s[(size_t)(y - hh - 1) * W + (x - hw - 1)], b = s[(size_t)(y - hh - 1) * W + (x + hw)];
    const float c = s[(size_t)(y + hh) * W + (x + hw)], d = s[(size_t)(y + hh) * W + (x - hw - 1)];
    return __fsub_rn(__fsub_rn(__fadd_rn(c, a), b), d);
}
__device__ __forceinline__ float min_eigenvalue(float gxx, float gxy, float gyy) {
    const float d = __fsub_rn(gxx, gyy);
    const float dd = __fmul_rn(d, d);
    const double t = __dadd_rn((double)dd, __dmul_rn(__dmul_rn(4.0, (double)gxy), (double)gxy));
    const float sqrtTerm = __double2float_rn(sqrt(t));   // reference: pow(t, 0.5); IEEE sqrt is the correctly rounded value
    const float s = __fsub_rn(__fadd_rn(gxx, gyy), sqrtTerm);
    return __double2float_rn(__ddiv_rn((double)s, 2.0));
}

// key layout (ascending sort of ~key == descending (val, x, y)): [val bits 32][x 13][y 13]
__device__ __forceinline__ unsigned long long make_key(float val, int x, int y) {
    const unsigned long long k = ((unsigned long long)__float_as_uint(val) << 26) | ((unsigned long long)x << 13) |
                                 (unsigned long long)y;
    return ~k;
}

__global__ void __launch_bounds__(256)
eigen_kernel(const float *__restrict__ sxx, const float *__restrict__ sxy, const float *__restrict__ syy, int W,
             int bx, int by, int hw, int hh, int step, int nx, int ny, float *__restrict__ val_out,
             unsigned long long *__restrict__ keys, unsigned int *__restrict__ nkeys, float min_val) {
    __shared__ unsigned int warp_cnt[8];
    __shared__ unsigned int block_base;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const bool inside = i < nx && j < ny;
    const int x = bx + i * step, y = by + j * step;
    float v = 0.f;
    if (inside) {
        v = min_eigenvalue(window_sum(sxx, W, x, y, hw, hh), window_sum(sxy, W, x, y, hw, hh),
                           window_sum(syy, W, x, y, hw, hh));
        if (val_out) val_out[(size_t)j * nx + i] = v;
    }
    if (!keys) return;
    // block-aggregated compaction: candidates below min_eigenvalue can never be accepted (:116), drop them
    const bool keep = inside && v >= min_val;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tot = 0;
        for (int w = 0; w < 8; w++) { const unsigned int c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
        block_base = tot ? atomicAdd(nkeys, tot) : 0u;
    }
    __syncthreads();
    if (keep) keys[block_base + warp_cnt[warp] + __popc(m & ((1u << lane) - 1u))] = make_key(v, x, y);
}

// ---- LSD radix sort of 64-bit keys (8-bit digits, stable) -----------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 16
#define RS_CHUNK (RS_THREADS * RS_ITEMS)

__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long *__restrict__ keys, const unsigned int *__restrict__ n_ptr, int shift,
               unsigned int *__restrict__ hist, int nblocks) {
    __shared__ unsigned int h[256];
    const unsigned int n = *n_ptr;
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_CHUNK;
    for (int r = 0; r < RS_ITEMS; r++) {
        const size_t i = base + (size_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned int)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of hist[0..total) in place, one CTA
__global__ void __launch_bounds__(1024)
rs_scan_kernel(unsigned int *__restrict__ hist, int total) {
    __shared__ unsigned int part[1024];
    const int t = threadIdx.x;
    const int per = (total + 1023) / 1024;
    const int lo = min(t * per, total), hi = min(lo + per, total);
    unsigned int s = 0;
    for (int i = lo; i < hi; i++) s += hist[i];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {       // Hillis-Steele inclusive scan
        unsigned int v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned int run = part[t] - s;
    for (int i = lo; i < hi; i++) { const unsigned int v = hist[i]; hist[i] = run; run += v; }
}

__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out,
                  const unsigned int *__restrict__ n_ptr, int shift, const unsigned int *__restrict__ hist, int nblocks) {
    __shared__ unsigned int running[256];
    __shared__ unsigned int wh[RS_THREADS / 32][256];
    const unsigned int n = *n_ptr;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    running[t] = hist[(size_t)t * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; w++) wh[w][t] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_CHUNK;
    for (int r = 0; r < RS_ITEMS; r++) {
        const size_t i = base + (size_t)r * RS_THREADS + t;
        const bool valid = i < n;
        const unsigned long long key = valid ? in[i] : 0ull;
        const unsigned int d = valid ? ((unsigned int)(key >> shift) & 255u) : (0x1000u + lane);
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const unsigned int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) wh[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            unsigned int off = running[d] + rank;
            for (int w = 0; w < warp; w++) off += wh[w][d];
            out[off] = key;
        }
        __syncthreads();
        unsigned int tot = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) { tot += wh[w][t]; wh[w][t] = 0; }
        running[t] += tot;
        __syncthreads();
    }
}

// ---- greedy minimum-distance suppression (_enforceMinimumDistance) ----------------------------------------
// One warp walks the sorted candidates 32 at a time.  A candidate is dead if the byte map says a better
// feature already claimed its pixel, or if a feature accepted earlier IN THE SAME batch lies within
// Chebyshev distance r = mindist-1 (which is exactly what the map would say after that feature was marked).
// This reproduces the sequential greedy walk exactly, including the order in which slots are filled.
struct GreedyArgs {
    const unsigned long long *keys;
    const unsigned int *nkeys;
    unsigned char *map;
    int W, H, r, n_features, overwrite;
    double *fx, *fy;
    int *fval;
    unsigned long long *consumed;
};

__device__ __forceinline__ void mark_region(unsigned char *map, int W, int H, int x, int y, int r, int lane) {
    const int side = 2 * r + 1;
    for (int idx = lane; idx < side * side; idx += 32) {
        const int iy = y - r + idx / side, ix = x - r + idx % side;
        if (ix >= 0 && ix < W && iy >= 0 && iy < H) map[(size_t)iy * W + ix] = 1;
    }
}

__global__ void __launch_bounds__(32)
greedy_kernel(const __grid_constant__ GreedyArgs A) {
    const int lane = threadIdx.x;
    const unsigned int n = *A.nkeys;
    volatile unsigned char *vmap = A.map;
    int indx = 0;
    if (!A.overwrite) {
        for (int f = 0; f < A.n_features; f++)               // :64-69 pre-mark surviving features
            if (A.fval[f] >= 0) mark_region(A.map, A.W, A.H, (int)A.fx[f], (int)A.fy[f], A.r, lane);
        __syncwarp();
        __threadfence_block();
        while (indx < A.n_features && A.fval[indx] >= 0) indx++;
    }
    unsigned long long pi = 0;
    bool full = indx >= A.n_features;
    // the reference reads one more candidate before noticing that every slot is taken (:96-112)
    while (pi < n && !full) {
        const unsigned long long i = pi + lane;
        const bool valid = i < n;
        const unsigned long long k = valid ? ~A.keys[i] : 0ull;
        const int x = (int)((k >> 13) & 8191ull), y = (int)(k & 8191ull);
        const float val = __uint_as_float((unsigned int)(k >> 26));
        bool live = valid && vmap[(size_t)y * A.W + x] == 0;
        unsigned int m;
        int last = -1;
        while ((m = __ballot_sync(0xffffffffu, live)) != 0u) {
            const int leader = __ffs(m) - 1;
            const int lx = __shfl_sync(0xffffffffu, x, leader), ly = __shfl_sync(0xffffffffu, y, leader);
            const float lval = __shfl_sync(0xffffffffu, val, leader);
            if (lane == 0) { A.fx[indx] = (double)lx; A.fy[indx] = (double)ly; A.fval[indx] = (int)lval; }
            mark_region(A.map, A.W, A.H, lx, ly, A.r, lane);
            if (lane == leader || (abs(x - lx) <= A.r && abs(y - ly) <= A.r)) live = false;
            indx++;
            if (!A.overwrite) while (indx < A.n_features && A.fval[indx] >= 0) indx++;
            last = leader;
            if (indx >= A.n_features) { full = true; break; }
        }
        __syncwarp();
        __threadfence_block();
        if (full) { pi += (unsigned long long)last + 1; if (pi < n) pi += 1; }
        else pi += 32;
    }
    if (pi > n) pi = n;
    if (!full && A.overwrite && lane == 0)
        for (int f = indx; f < A.n_features; f++) { A.fx[f] = -1.0; A.fy[f] = -1.0; A.fval[f] = KLT_NOT_FOUND; }
    if (lane == 0 && A.consumed) *A.consumed = pi;
}

// ---------------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int klt_launch_scan(klt_ctx *ctx, const float *gx, const float *gy, size_t pitch, int w, int h, int bx, int by, int hw,
                    int hh, int skip, float *val_dev, int nx, int ny) {
    // workspace: 3 SATs
    const size_t plane = align_up((size_t)w * h * sizeof(float), 256);
    int rc = klt_ws_reserve(ctx, 3 * plane);
    if (rc) return rc;
    float *sxx = (float *)ctx->ws, *sxy = (float *)((char *)ctx->ws + plane), *syy = (float *)((char *)ctx->ws + 2 * plane);
    sat_rows_kernel<<<(h + 31) / 32, 32, 0, ctx->stream>>>(gx, gy, pitch, w, h, sxx, sxy, syy);
    KLT_CHECK_LAUNCH(ctx);
    sat_cols_kernel<<<dim3((w + 127) / 128, 3), 128, 0, ctx->stream>>>(sxx, sxy, syy, w, h);
    KLT_CHECK_LAUNCH(ctx);
    if (nx > 0 && ny > 0) {
        eigen_kernel<<<dim3((nx + 255) / 256, ny), 256, 0, ctx->stream>>>(sxx, sxy, syy, w, bx, by, hw, hh, skip + 1, nx, ny,
                                                                        val_dev, nullptr, nullptr, 0.f);
        KLT_CHECK_LAUNCH(ctx);
    }
    return KLT_OK;
}

int klt_select_device(klt_ctx *ctx, const klt_params *p, const float *gx, const float *gy, size_t pitch, int w, int h,
                      int n_features, int replace, double *x, double *y, int32_t *val, int64_t *n_consumed) {
    if (w > 8191 || h > 8191) return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "image larger than 8191 pixels per side");
    // border and window exactly as selectGoodFeatures.py:168-169,215-221,230 (true division, then int truncation)
    double window_hw = p->window_width / 2.0, window_hh = p->window_height / 2.0;
    double bxd = p->borderx, byd = p->bordery;
    if (bxd < window_hw) bxd = window_hw;
    if (byd < window_hh) byd = window_hh;
    const int bx = (int)bxd, by = (int)byd, hw = (int)window_hw, hh = (int)window_hh;
    if (bx < hw + 1 || by < hh + 1)
        return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "border (%d,%d) smaller than window half-size + 1: the reference reads out of bounds here", bx, by);
    const int step = p->n_skipped_pixels + 1;
    int nx = 0, ny = 0;
    if (w - bx > bx) nx = (w - 2 * bx + step - 1) / step;
    if (h - by > by) ny = (h - 2 * by + step - 1) / step;
    const size_t ncand = (size_t)nx * ny;
    int mindist = p->mindist < 0 ? 0 : p->mindist;            // :241-243
    int min_eig = p->min_eigenvalue < 1 ? 1 : p->min_eigenvalue;   // :53

    const size_t plane = align_up((size_t)w * h * sizeof(float), 256);
    const size_t keys_b = align_up((ncand + 1) * sizeof(unsigned long long), 256);
    const int nblocks = (int)((ncand + RS_CHUNK - 1) / RS_CHUNK) + 1;
    const size_t hist_b = align_up((size_t)256 * nblocks * sizeof(unsigned int), 256);
    const size_t map_b = align_up((size_t)w * h, 256);
    const size_t feat_b = align_up((size_t)n_features * (2 * sizeof(double) + sizeof(int)) + 64, 256);
    const size_t total = 3 * plane + 2 * keys_b + hist_b + map_b + feat_b + 256;
    int rc = klt_ws_reserve(ctx, total);
    if (rc) return rc;
    char *wsp = (char *)ctx->ws;
    float *sxx = (float *)wsp; wsp += plane;
    float *sxy = (float *)wsp; wsp += plane;
    float *syy = (float *)wsp; wsp += plane;
    unsigned long long *keys0 = (unsigned long long *)wsp; wsp += keys_b;
    unsigned long long *keys1 = (unsigned long long *)wsp; wsp += keys_b;
    unsigned int *hist = (unsigned int *)wsp; wsp += hist_b;
    unsigned char *map = (unsigned char *)wsp; wsp += map_b;
    double *fx = (double *)wsp; double *fy = fx + n_features; int *fval = (int *)(fy + n_features); wsp += feat_b;
    unsigned int *nkeys = (unsigned int *)wsp;
    unsigned long long *consumed = (unsigned long long *)(wsp + 8);

    KLT_CUDA(ctx, cudaMemsetAsync(nkeys, 0, 16, ctx->stream));
    KLT_CUDA(ctx, cudaMemsetAsync(map, 0, (size_t)w * h, ctx->stream));
    if (replace) {
        KLT_CUDA(ctx, cudaMemcpyAsync(fx, x, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemcpyAsync(fy, y, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemcpyAsync(fval, val, n_features * sizeof(int), cudaMemcpyDefault, ctx->stream));
    }
    sat_rows_kernel<<<(h + 31) / 32, 32, 0, ctx->stream>>>(gx, gy, pitch, w, h, sxx, sxy, syy);
    KLT_CHECK_LAUNCH(ctx);
    sat_cols_kernel<<<dim3((w + 127) / 128, 3), 128, 0, ctx->stream>>>(sxx, sxy, syy, w, h);
    KLT_CHECK_LAUNCH(ctx);
    if (ncand) {
        eigen_kernel<<<dim3((nx + 255) / 256, ny), 256, 0, ctx->stream>>>(sxx, sxy, syy, w, bx, by, hw, hh, step, nx, ny, nullptr,
                                                                        keys0, nkeys, (float)min_eig);
        KLT_CHECK_LAUNCH(ctx);
        // 58 significant key bits -> 8 passes of 8 bits (the launch geometry covers the worst case ncand;
        // blocks beyond the actual key count see no valid items)
        unsigned long long *src = keys0, *dst = keys1;
        for (int pass = 0; pass < 8; pass++) {
            rs_hist_kernel<<<nblocks, RS_THREADS, 0, ctx->stream>>>(src, nkeys, pass * 8, hist, nblocks);
            KLT_CHECK_LAUNCH(ctx);
            rs_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, 256 * nblocks);
            KLT_CHECK_LAUNCH(ctx);
            rs_scatter_kernel<<<nblocks, RS_THREADS, 0, ctx->stream>>>(src, dst, nkeys, pass * 8, hist, nblocks);
            KLT_CHECK_LAUNCH(ctx);
            unsigned long long *t = src; src = dst; dst = t;
        }
        GreedyArgs G;
        G.keys = src; G.nkeys = nkeys; G.map = map; G.W = w; G.H = h; G.r = mindist - 1; G.n_features = n_features;
        G.overwrite = replace ? 0 : 1; G.fx = fx; G.fy = fy; G.fval = fval; G.consumed = consumed;
        greedy_kernel<<<1, 32, 0, ctx->stream>>>(G);
        KLT_CHECK_LAUNCH(ctx);
    } else {
        GreedyArgs G;
        G.keys = keys0; G.nkeys = nkeys; G.map = map; G.W = w; G.H = h; G.r = mindist - 1; G.n_features = n_features;
        G.overwrite = replace ? 0 : 1; G.fx = fx; G.fy = fy; G.fval = fval; G.consumed = consumed;
        greedy_kernel<<<1, 32, 0, ctx->stream>>>(G);
        KLT_CHECK_LAUNCH(ctx);
    }
    KLT_CUDA(ctx, cudaMemcpyAsync(x, fx, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(y, fy, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(val, fval, n_features * sizeof(int), cudaMemcpyDefault, ctx->stream));
    unsigned long long cons = 0;
    if (n_consumed) KLT_CUDA(ctx, cudaMemcpyAsync(&cons, consumed, sizeof(cons), cudaMemcpyDeviceToHost, ctx->stream));
    KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_consumed) *n_consumed = (int64_t)cons;
    return KLT_OK;
}
