// Good-feature selection on the GPU (sm_100a).
//
// Replaces goodFeaturesUtils.ScanImageForGoodFeatures (goodFeaturesUtils.pyx:35-73), the candidate sort
// (selectGoodFeatures.py:234-236) and _enforceMinimumDistance (selectGoodFeatures.py:45-135).
//
// The reference's eigenvalues carry the rounding of three float32 summed-area tables built by strictly
// sequential additions (np.cumsum along rows, then along columns); selection ORDER depends on that rounding
// (SURVEY 7.3), so the tables are rebuilt here with the same chains: one lane per row (tiles transposed through
// shared memory so that global traffic stays coalesced), then one thread per column.  A parallel prefix scan would
// be faster per element but produces different float32 sums.
//
// The reference sorts every candidate (1.87 M at 1080p) although greedy suppression consumes only the best ~8.5 N.
// Here a histogram of the eigenvalues picks a threshold that keeps roughly the best M = 32 N + 8192 candidates, only
// those are sorted (stable LSD radix sort on 64-bit keys) and walked; if the walk runs out of candidates before every
// slot is filled, the selection is repeated without the threshold, so the result is always the exact greedy result.
#include <cuda_pipeline.h>

#include "klt_common.cuh"

// ---- summed-area tables -------------------------------------------------------------------------------
// rows: s[y][x] = s[y][x-1] + p[y][x], p = exact fp32 product (np.power(g,2.) / g*g, pyx:49-51).
// One warp owns 32 rows.  32x32 tiles of gx, gy are brought in with cp.async (double buffered, coalesced), each lane
// then walks ITS row of the tile sequentially (the float32 chain of np.cumsum), and the three result tiles go back
// transposed so that the stores are coalesced too.
#define SAT_T 32
#define SAT_NBUF 4                      // cp.async ring: tiles are requested 3 chunks ahead of their use
struct SatSmem {
    float in[SAT_NBUF][2][SAT_T][SAT_T + 1];   // [buffer][gx|gy][row][col]
    float out[3][SAT_T][SAT_T + 1];
};

__global__ void __launch_bounds__(32)
sat_rows_kernel(const float *__restrict__ gx, const float *__restrict__ gy, size_t pitch, int W, int H,
                float *__restrict__ sxx, float *__restrict__ sxy, float *__restrict__ syy) {
    __shared__ SatSmem sm;
    const int lane = threadIdx.x;
    const int y0 = blockIdx.x * SAT_T;
    const int nchunks = (W + SAT_T - 1) / SAT_T;
    auto issue = [&](int chunk, int buf) {
        const int x = chunk * SAT_T + lane;
        if (x < W) {
#pragma unroll 8
            for (int i = 0; i < SAT_T; i++) {
                const int y = min(y0 + i, H - 1);
                __pipeline_memcpy_async(&sm.in[buf][0][i][lane], gx + (size_t)y * pitch + x, 4);
                __pipeline_memcpy_async(&sm.in[buf][1][i][lane], gy + (size_t)y * pitch + x, 4);
            }
        }
        __pipeline_commit();
    };
    float axx = 0.f, axy = 0.f, ayy = 0.f;     // 0 + p == p exactly, so starting from 0 equals cumsum's first copy
    for (int c = 0; c < SAT_NBUF - 1; c++) {
        if (c < nchunks) issue(c, c); else __pipeline_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        const int buf = c % SAT_NBUF;
        if (c + SAT_NBUF - 1 < nchunks) issue(c + SAT_NBUF - 1, (c + SAT_NBUF - 1) % SAT_NBUF); else __pipeline_commit();
        __pipeline_wait_prior(SAT_NBUF - 1);
        __syncwarp();
        const int ncol = min(SAT_T, W - c * SAT_T);
        for (int k = 0; k < ncol; k++) {       // lane = row: sequential chain along x
            const float a = sm.in[buf][0][lane][k], b = sm.in[buf][1][lane][k];
            axx = __fadd_rn(axx, __fmul_rn(a, a));
            axy = __fadd_rn(axy, __fmul_rn(a, b));
            ayy = __fadd_rn(ayy, __fmul_rn(b, b));
            sm.out[0][lane][k] = axx; sm.out[1][lane][k] = axy; sm.out[2][lane][k] = ayy;
        }
        __syncwarp();
        const int x = c * SAT_T + lane;
        if (x < W) {
#pragma unroll 8
            for (int i = 0; i < SAT_T; i++) {
                const int y = y0 + i;
                if (y < H) {
                    const size_t o = (size_t)y * W + x;
                    sxx[o] = sm.out[0][i][lane]; sxy[o] = sm.out[1][i][lane]; syy[o] = sm.out[2][i][lane];
                }
            }
        }
        __syncwarp();
    }
}

// Same algorithm for W % 32 == 0 and 16-byte aligned rows: 128-bit cp.async, 128-bit shared loads/stores (row stride 36
// floats keeps every quarter-warp on distinct banks), fully unrolled so that only the 32-step float32 add chains are serial.
#define SAT_S 36
struct SatSmemV {
    float in[SAT_NBUF][2][SAT_T][SAT_S];
    float out[3][SAT_T][SAT_S];
};
__global__ void __launch_bounds__(32)
sat_rows_vec_kernel(const float *__restrict__ gx, const float *__restrict__ gy, size_t pitch, int W, int H,
                    float *__restrict__ sxx, float *__restrict__ sxy, float *__restrict__ syy) {
    extern __shared__ __align__(16) unsigned char sat_raw[];
    SatSmemV &sm = *reinterpret_cast<SatSmemV *>(sat_raw);
    const int lane = threadIdx.x;
    const int y0 = blockIdx.x * SAT_T;
    const int nchunks = W / SAT_T;
    const int q = lane & 7, r8 = lane >> 3;            // 16-byte chunk within a tile row, row within a group of 4
    auto issue = [&](int chunk, int buf) {
        const int x = chunk * SAT_T + 4 * q;
#pragma unroll
        for (int g = 0; g < SAT_T / 4; g++) {
            const int i = 4 * g + r8;
            const int y = min(y0 + i, H - 1);
            __pipeline_memcpy_async(&sm.in[buf][0][i][4 * q], gx + (size_t)y * pitch + x, 16);
            __pipeline_memcpy_async(&sm.in[buf][1][i][4 * q], gy + (size_t)y * pitch + x, 16);
        }
        __pipeline_commit();
    };
    float axx = 0.f, axy = 0.f, ayy = 0.f;
    for (int c = 0; c < SAT_NBUF - 1; c++) {
        if (c < nchunks) issue(c, c); else __pipeline_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        const int buf = c % SAT_NBUF;
        if (c + SAT_NBUF - 1 < nchunks) issue(c + SAT_NBUF - 1, (c + SAT_NBUF - 1) % SAT_NBUF); else __pipeline_commit();
        __pipeline_wait_prior(SAT_NBUF - 1);
        __syncwarp();
        float4 va[8], vb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            va[k] = *reinterpret_cast<const float4 *>(&sm.in[buf][0][lane][4 * k]);
            vb[k] = *reinterpret_cast<const float4 *>(&sm.in[buf][1][lane][4 * k]);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {          // lane = row: the sequential float32 chain of np.cumsum along x
            float4 oxx, oxy, oyy;
            axx = __fadd_rn(axx, __fmul_rn(va[k].x, va[k].x)); axy = __fadd_rn(axy, __fmul_rn(va[k].x, vb[k].x)); ayy = __fadd_rn(ayy, __fmul_rn(vb[k].x, vb[k].x));
            oxx.x = axx; oxy.x = axy; oyy.x = ayy;
            axx = __fadd_rn(axx, __fmul_rn(va[k].y, va[k].y)); axy = __fadd_rn(axy, __fmul_rn(va[k].y, vb[k].y)); ayy = __fadd_rn(ayy, __fmul_rn(vb[k].y, vb[k].y));
            oxx.y = axx; oxy.y = axy; oyy.y = ayy;
            axx = __fadd_rn(axx, __fmul_rn(va[k].z, va[k].z)); axy = __fadd_rn(axy, __fmul_rn(va[k].z, vb[k].z)); ayy = __fadd_rn(ayy, __fmul_rn(vb[k].z, vb[k].z));
            oxx.z = axx; oxy.z = axy; oyy.z = ayy;
            axx = __fadd_rn(axx, __fmul_rn(va[k].w, va[k].w)); axy = __fadd_rn(axy, __fmul_rn(va[k].w, vb[k].w)); ayy = __fadd_rn(ayy, __fmul_rn(vb[k].w, vb[k].w));
            oxx.w = axx; oxy.w = axy; oyy.w = ayy;
            *reinterpret_cast<float4 *>(&sm.out[0][lane][4 * k]) = oxx;
            *reinterpret_cast<float4 *>(&sm.out[1][lane][4 * k]) = oxy;
            *reinterpret_cast<float4 *>(&sm.out[2][lane][4 * k]) = oyy;
        }
        __syncwarp();
        const int x = c * SAT_T + 4 * q;
#pragma unroll
        for (int g = 0; g < SAT_T / 4; g++) {
            const int i = 4 * g + r8, y = y0 + i;
            if (y < H) {
                const size_t o = (size_t)y * W + x;
                *reinterpret_cast<float4 *>(sxx + o) = *reinterpret_cast<const float4 *>(&sm.out[0][i][4 * q]);
                *reinterpret_cast<float4 *>(sxy + o) = *reinterpret_cast<const float4 *>(&sm.out[1][i][4 * q]);
                *reinterpret_cast<float4 *>(syy + o) = *reinterpret_cast<const float4 *>(&sm.out[2][i][4 * q]);
            }
        }
        __syncwarp();
    }
}
// columns: s[y][x] = s[y-1][x] + s[y][x]; blockIdx.y selects the table
__global__ void __launch_bounds__(64)
sat_cols_kernel(float *__restrict__ s0, float *__restrict__ s1, float *__restrict__ s2, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    float *s = (blockIdx.y == 0 ? s0 : blockIdx.y == 1 ? s1 : s2) + x;
    float acc = s[0];
    int y = 1;
    for (; y + 16 <= H; y += 16) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = s[(size_t)(y + i) * W];
#pragma unroll
        for (int i = 0; i < 16; i++) { acc = __fadd_rn(acc, v[i]); s[(size_t)(y + i) * W] = acc; }
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

// histogram bin of an eigenvalue >= 1: exponent and 5 mantissa bits (3 % resolution), 2048 bins
#define EIG_BINS 2048
__device__ __forceinline__ int eig_bin(float v) {
    const int b = (int)(__float_as_uint(v) >> 18) - (0x3F800000 >> 18);
    return min(max(b, 0), EIG_BINS - 1);
}

// eigenvalue map (+ optional histogram of the values >= min_val).  One block = 256 columns x EIG_ROWS candidate rows.
#define EIG_ROWS 8
__global__ void __launch_bounds__(256)
eigen_kernel(const float *__restrict__ sxx, const float *__restrict__ sxy, const float *__restrict__ syy, int W,
             int bx, int by, int hw, int hh, int step, int nx, int ny, float *__restrict__ val_out,
             unsigned int *__restrict__ hist, float min_val) {
    __shared__ unsigned int h[EIG_BINS];
    if (hist) {
        for (int b = threadIdx.x; b < EIG_BINS; b += 256) h[b] = 0;
        __syncthreads();
    }
    const int i = blockIdx.x * 256 + threadIdx.x;
    for (int jj = 0; jj < EIG_ROWS; jj++) {
        const int j = blockIdx.y * EIG_ROWS + jj;
        if (i < nx && j < ny) {
            const int x = bx + i * step, y = by + j * step;
            const float v = min_eigenvalue(window_sum(sxx, W, x, y, hw, hh), window_sum(sxy, W, x, y, hw, hh),
                                           window_sum(syy, W, x, y, hw, hh));
            val_out[(size_t)j * nx + i] = v;
            if (hist && v >= min_val) atomicAdd(&h[eig_bin(v)], 1u);
        }
    }
    if (hist) {
        __syncthreads();
        for (int b = threadIdx.x; b < EIG_BINS; b += 256)
            if (h[b]) atomicAdd(&hist[b], h[b]);
    }
}

// picks the highest bin T such that at least `target` candidates lie in bins >= T (T = 0 if there are fewer)
__global__ void __launch_bounds__(32)
threshold_kernel(const unsigned int *__restrict__ hist, unsigned int target, int force_all, unsigned int *__restrict__ out /* [0]=T */) {
    const int lane = threadIdx.x;
    constexpr int PER = EIG_BINS / 32;
    unsigned int mine = 0;
    for (int b = 0; b < PER; b++) mine += hist[lane * PER + b];
    unsigned int suffix = mine;                       // inclusive suffix sum over lanes (lane 31 = highest bins)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += v;
    }
    const unsigned int above = suffix - mine;         // candidates in lanes above this one
    const bool here = above < target && suffix >= target;
    const unsigned int m = __ballot_sync(0xffffffffu, here);
    if (force_all || m == 0u) { if (lane == 0) out[0] = 0; return; }
    if (here) {
        unsigned int acc = above;
        int T = lane * PER;
        for (int b = PER - 1; b >= 0; b--) {
            acc += hist[lane * PER + b];
            if (acc >= target) { T = lane * PER + b; break; }
        }
        out[0] = (unsigned int)T;
    }
}

// compaction of the candidates with bin >= T into 64-bit keys (block-aggregated append; order is irrelevant, the
// keys are unique and get sorted)
__global__ void __launch_bounds__(256)
compact_kernel(const float *__restrict__ val, int bx, int by, int step, int nx, int ny, float min_val,
               const unsigned int *__restrict__ thr, unsigned long long *__restrict__ keys, unsigned int *__restrict__ nkeys) {
    __shared__ unsigned int warp_cnt[8];
    __shared__ unsigned int block_base;
    const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
    const int T = (int)thr[0];
    float v = 0.f;
    bool keep = false;
    if (i < nx && j < ny) {
        v = val[(size_t)j * nx + i];
        keep = v >= min_val && eig_bin(v) >= T;     // candidates below min_eigenvalue can never be accepted (:116)
    }
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
    if (keep) keys[block_base + warp_cnt[warp] + __popc(m & ((1u << lane) - 1u))] = make_key(v, bx + i * step, by + j * step);
}

// ---- LSD radix sort of 64-bit keys (8-bit digits, stable) -----------------------------------------------
#define RS_THREADS 256
#define RS_ITEMS 4
#define RS_CHUNK (RS_THREADS * RS_ITEMS)
#define RS_FUSED_SCAN_BLOCKS 96   // up to this many blocks each scatter block scans the raw counts itself

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

template <bool SCANNED>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out,
                  const unsigned int *__restrict__ n_ptr, int shift, const unsigned int *__restrict__ hist, int nblocks) {
    __shared__ unsigned int running[256];
    __shared__ unsigned int wh[RS_THREADS / 32][256];
    __shared__ unsigned int wtot[RS_THREADS / 32];
    const unsigned int n = *n_ptr;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (SCANNED) {
        running[t] = hist[(size_t)t * nblocks + blockIdx.x];
    } else {
        // few blocks: every block derives its own offsets from the raw per-block counts (saves the scan launch).
        // offset(digit t, this block) = keys with a smaller digit + keys with digit t in earlier blocks
        unsigned int total = 0, before = 0;
        for (int b = 0; b < nblocks; b++) {
            const unsigned int c = hist[(size_t)t * nblocks + b];
            total += c;
            if (b < (int)blockIdx.x) before += c;
        }
        unsigned int incl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) wtot[warp] = incl;
        __syncthreads();
        unsigned int wbase = 0;
        for (int w = 0; w < warp; w++) wbase += wtot[w];
        running[t] = wbase + incl - total + before;
    }
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
// One warp walks the sorted candidates 32 at a time.  Accepted features are remembered in a grid of cells of side
// cs = r + 1 (r = mindist - 1): two features in one cell would be closer than mindist, so a cell holds at most one and
// a candidate only has to look at its 3x3 cell neighbourhood (shared memory, or global memory for very large images).
// Features that survive from a previous frame (replacement mode, :64-69) are pre-marked in a byte map by a separate
// kernel (they may be closer to each other than mindist); the map value travels with the prefetched key.
// A candidate also dies if a feature accepted earlier IN THE SAME batch lies within Chebyshev distance r.  This
// reproduces the sequential greedy walk exactly, including the order in which slots are filled.
struct GreedyArgs {
    const unsigned long long *keys;
    const unsigned int *nkeys;
    const unsigned char *premap;     // NULL in SELECTING_ALL mode
    unsigned short *grid_global;     // used when the cell grid does not fit in shared memory
    int W, H, r, n_features, overwrite;
    int cs, gw, gh, grid_in_smem;
    double *fx, *fy;
    int *fval;
    int *free_slots;                 // [n_features] scratch: the fillable slots in list order (replacement mode)
    unsigned long long *consumed;    // [0] candidates consumed, [1] 1 if the keys ran out before all slots were filled
};

__global__ void __launch_bounds__(256)
premark_kernel(const double *__restrict__ fx, const double *__restrict__ fy, const int *__restrict__ fval, int n,
               unsigned char *__restrict__ map, int W, int H, int r) {
    const int f = blockIdx.x;
    if (f >= n || fval[f] < 0) return;
    const int x = (int)fx[f], y = (int)fy[f], side = 2 * r + 1;
    for (int idx = threadIdx.x; idx < side * side; idx += blockDim.x) {
        const int iy = y - r + idx / side, ix = x - r + idx % side;
        if (ix >= 0 && ix < W && iy >= 0 && iy < H) map[(size_t)iy * W + ix] = 1;
    }
}

#define GREEDY_THREADS 1024
// is candidate (x, y) within Chebyshev distance r of a feature registered in the 3x3 cell neighbourhood?
__device__ __forceinline__ bool grid_conflict(const unsigned short *grid, int gw, int gh, int cs, int r, int x, int y) {
    const int cx = x / cs, cy = y / cs;
    bool hit = false;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            const int ncx = cx + dx, ncy = cy + dy;
            if (ncx >= 0 && ncx < gw && ncy >= 0 && ncy < gh) {
                const unsigned int g = grid[ncy * gw + ncx];
                if (g != 0xFFFFu) {
                    const int fx = ncx * cs + (int)(g >> 8), fy = ncy * cs + (int)(g & 255u);
                    hit |= abs(fx - x) <= r && abs(fy - y) <= r;
                }
            }
        }
    return hit;
}

// One CTA.  Candidates are taken 1024 at a time: (1) all 32 warps test their candidate against the features accepted in
// EARLIER super-batches (most candidates die here) and compact the survivors, in order, into shared memory; (2) warp 0
// walks the survivors 32 at a time with exactly the result of the sequential reference loop: re-test against the grid
// (features accepted earlier in this super-batch), then resolve the batch in parallel -- lane i is accepted iff no EARLIER
// accepted lane lies within distance r.  That recurrence is solved by iteration: an undecided lane is rejected as soon as
// an accepted earlier lane conflicts with it and accepted as soon as all its earlier conflicting lanes are rejected; the
// lowest undecided lane is decided in every round, typical batches need 2-4 rounds.  Accepted lanes then fill their slots
// (the k-th accepted candidate takes the k-th fillable slot) and register in the grid all at once.
__global__ void __launch_bounds__(GREEDY_THREADS)
greedy_kernel(const __grid_constant__ GreedyArgs A) {
    extern __shared__ unsigned short grid_smem[];
    __shared__ unsigned long long surv_key[GREEDY_THREADS];
    __shared__ unsigned int surv_idx[GREEDY_THREADS];
    __shared__ unsigned int warp_cnt[32];
    __shared__ int s_filled, s_slots, s_full, s_nsurv;
    __shared__ unsigned long long s_consumed;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int n = *A.nkeys;
    unsigned short *grid = A.grid_in_smem ? grid_smem : A.grid_global;
    if (A.r >= 0)
        for (int c = tid; c < A.gw * A.gh; c += GREEDY_THREADS) grid[c] = 0xFFFFu;
    // fillable slots: every slot (SELECTING_ALL) or, in list order, the slots of lost features (:64-69, :110-112)
    if (tid == 0) { s_slots = A.overwrite ? A.n_features : 0; s_filled = 0; s_consumed = 0ull; }
    __syncthreads();
    if (!A.overwrite) {
        for (int f0 = 0; f0 < A.n_features; f0 += GREEDY_THREADS) {
            const int f = f0 + tid;
            const bool lost = f < A.n_features && A.fval[f] < 0;
            const unsigned int m = __ballot_sync(0xffffffffu, lost);
            if (lane == 0) warp_cnt[warp] = __popc(m);
            __syncthreads();
            if (warp == 0) {
                unsigned int c = warp_cnt[lane], incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                warp_cnt[lane] = incl - c;
                if (lane == 31) s_nsurv = (int)incl;
            }
            __syncthreads();
            if (lost) A.free_slots[s_slots + warp_cnt[warp] + __popc(m & ((1u << lane) - 1u))] = f;
            __syncthreads();
            if (tid == 0) s_slots += s_nsurv;
            __syncthreads();
        }
    }
    if (tid == 0) s_full = s_slots == 0;
    __syncthreads();
    for (unsigned long long base = 0; base < n && !s_full; base += GREEDY_THREADS) {
        // ---- phase 1: parallel test against earlier super-batches, ordered compaction of the survivors ----
        const unsigned long long i = base + tid;
        bool live = i < n;
        unsigned long long k = 0ull;
        if (live) {
            k = ~A.keys[i];
            const int x = (int)((k >> 13) & 8191ull), y = (int)(k & 8191ull);
            if (A.premap && A.premap[(size_t)y * A.W + x]) live = false;
            if (live && A.r >= 0 && grid_conflict(grid, A.gw, A.gh, A.cs, A.r, x, y)) live = false;
        }
        const unsigned int m = __ballot_sync(0xffffffffu, live);
        if (lane == 0) warp_cnt[warp] = __popc(m);
        __syncthreads();
        if (warp == 0) {
            unsigned int c = warp_cnt[lane], incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            warp_cnt[lane] = incl - c;
            if (lane == 31) s_nsurv = (int)incl;
        }
        __syncthreads();
        if (live) {
            const unsigned int pos = warp_cnt[warp] + __popc(m & ((1u << lane) - 1u));
            surv_key[pos] = k; surv_idx[pos] = (unsigned int)(i - base);
        }
        __syncthreads();
        // ---- phase 2: warp 0 walks the survivors in rank order ----
        if (warp == 0) {
            int filled = s_filled;
            const int slots = s_slots;
            bool full = false;
            unsigned int last_off = 0;
            const int ns = s_nsurv;
            const unsigned int lt = (1u << lane) - 1u;
            for (int b0 = 0; b0 < ns && !full; b0 += 32) {
                const bool valid = b0 + lane < ns;
                const unsigned long long kk = valid ? surv_key[b0 + lane] : 0ull;
                const unsigned int off = valid ? surv_idx[b0 + lane] : 0u;
                const int x = (int)((kk >> 13) & 8191ull), y = (int)(kk & 8191ull);
                const float val = __uint_as_float((unsigned int)(kk >> 26));
                bool lv = valid;
                if (lv && A.r >= 0 && grid_conflict(grid, A.gw, A.gh, A.cs, A.r, x, y)) lv = false;
                const unsigned int live_mask = __ballot_sync(0xffffffffu, lv);
                if (live_mask == 0u) continue;
                // earlier live lanes of the batch within distance r of this one
                unsigned int confl = 0u;
                if (A.r >= 0) {
                    for (unsigned int mm = live_mask; mm; mm &= mm - 1u) {
                        const int j = __ffs(mm) - 1;
                        const int xj = __shfl_sync(0xffffffffu, x, j), yj = __shfl_sync(0xffffffffu, y, j);
                        if (j < lane && abs(x - xj) <= A.r && abs(y - yj) <= A.r) confl |= 1u << j;
                    }
                }
                unsigned int acc = 0u, rej = 0u, undecided = live_mask;
                while (undecided) {
                    const bool mine = (undecided >> lane) & 1u;
                    const bool r_now = mine && (confl & acc) != 0u;
                    const bool a_now = mine && !r_now && (confl & ~rej) == 0u;      // (confl & ~rej) has no accepted bit here
                    const unsigned int na = __ballot_sync(0xffffffffu, a_now), nr = __ballot_sync(0xffffffffu, r_now);
                    acc |= na; rej |= nr; undecided &= ~(na | nr);
                }
                // the k-th accepted candidate takes the k-th fillable slot; stop where the slots run out
                const int nacc = __popc(acc), room = slots - filled;
                const int take = nacc < room ? nacc : room;
                const int rank = __popc(acc & lt);
                const bool store = ((acc >> lane) & 1u) && rank < take;
                __syncwarp();                     // every lane has finished reading the grid before anyone registers in it
                if (store) {
                    const int slot = A.overwrite ? filled + rank : A.free_slots[filled + rank];
                    A.fx[slot] = (double)x; A.fy[slot] = (double)y; A.fval[slot] = (int)val;
                    if (A.r >= 0) {
                        const int cx = x / A.cs, cy = y / A.cs;
                        grid[cy * A.gw + cx] = (unsigned short)(((x - cx * A.cs) << 8) | (y - cy * A.cs));
                    }
                }
                const unsigned int stored = __ballot_sync(0xffffffffu, store);
                last_off = __shfl_sync(0xffffffffu, off, 31 - __clz(stored));     // the last candidate that was accepted
                filled += take;
                if (filled >= slots) full = true;
                __syncwarp();
            }
            if (lane == 0) {
                s_filled = filled;
                if (full) {
                    // the reference reads one more candidate before noticing that every slot is taken (:96-112)
                    unsigned long long c = base + last_off + 1;
                    if (c < n) c += 1;
                    s_consumed = c; s_full = 1;
                } else {
                    s_consumed = base + GREEDY_THREADS < n ? base + GREEDY_THREADS : (unsigned long long)n;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        A.consumed[0] = s_consumed;
        A.consumed[1] = s_full ? 0ull : 1ull;
        A.consumed[2] = (unsigned long long)s_filled;      // SELECTING_ALL: also the first unfilled slot
    }
}

// fills the slots the walk could not fill (SELECTING_ALL only): x = y = -1, val = KLT_NOT_FOUND (C-KLT behaviour, quirk Q6)
__global__ void fill_not_found_kernel(double *fx, double *fy, int *fval, int n, const unsigned long long *consumed) {
    const int f = (int)consumed[2] + blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n) { fx[f] = -1.0; fy[f] = -1.0; fval[f] = KLT_NOT_FOUND; }
}

// ---------------------------------------------------------------------------------------------------------
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int launch_sat(klt_ctx *ctx, const float *gx, const float *gy, size_t pitch, int w, int h, float *sxx, float *sxy, float *syy) {
    const bool vec = (w % SAT_T) == 0 && (pitch % 4) == 0 && ((reinterpret_cast<uintptr_t>(gx) | reinterpret_cast<uintptr_t>(gy)) & 15) == 0;
    if (vec) {
        KLT_CUDA(ctx, cudaFuncSetAttribute(sat_rows_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SatSmemV)));
        KLT_LAUNCH(ctx, "sat_rows", 20.0 * w * h, (sat_rows_vec_kernel<<<(h + SAT_T - 1) / SAT_T, 32, sizeof(SatSmemV), ctx->stream>>>(gx, gy, pitch, w, h, sxx, sxy, syy)));
    } else
    KLT_LAUNCH(ctx, "sat_rows", 20.0 * w * h, (sat_rows_kernel<<<(h + SAT_T - 1) / SAT_T, 32, 0, ctx->stream>>>(gx, gy, pitch, w, h, sxx, sxy, syy)));
    KLT_LAUNCH(ctx, "sat_cols", 24.0 * w * h, (sat_cols_kernel<<<dim3((w + 63) / 64, 3), 64, 0, ctx->stream>>>(sxx, sxy, syy, w, h)));
    return KLT_OK;
}

int klt_launch_scan(klt_ctx *ctx, const float *gx, const float *gy, size_t pitch, int w, int h, int bx, int by, int hw,
                    int hh, int skip, float *val_dev, int nx, int ny) {
    const size_t plane = align_up((size_t)w * h * sizeof(float), 256);
    int rc = klt_ws_reserve(ctx, 3 * plane);
    if (rc) return rc;
    float *sxx = (float *)ctx->ws, *sxy = (float *)((char *)ctx->ws + plane), *syy = (float *)((char *)ctx->ws + 2 * plane);
    if ((rc = launch_sat(ctx, gx, gy, pitch, w, h, sxx, sxy, syy))) return rc;
    if (nx > 0 && ny > 0)
        KLT_LAUNCH(ctx, "eigen", 52.0 * nx * ny, (eigen_kernel<<<dim3((nx + 255) / 256, (ny + EIG_ROWS - 1) / EIG_ROWS), 256, 0, ctx->stream>>>(
                                                      sxx, sxy, syy, w, bx, by, hw, hh, skip + 1, nx, ny, val_dev, nullptr, 0.f)));
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
    const int mindist = p->mindist < 0 ? 0 : p->mindist;                 // :241-243
    const int min_eig = p->min_eigenvalue < 1 ? 1 : p->min_eigenvalue;   // :53
    const int r = mindist - 1;                                           // :61
    if (r > 254) return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "mindist larger than 255");

    const size_t plane = align_up((size_t)w * h * sizeof(float), 256);
    const size_t val_b = align_up((ncand + 1) * sizeof(float), 256);
    const size_t keys_b = align_up((ncand + 1) * sizeof(unsigned long long), 256);
    const int max_blocks = (int)((ncand + RS_CHUNK - 1) / RS_CHUNK) + 1;
    const size_t hist_b = align_up((size_t)256 * max_blocks * sizeof(unsigned int), 256);
    const size_t map_b = align_up((size_t)w * h, 256);
    const size_t feat_b = align_up((size_t)n_features * (2 * sizeof(double) + 2 * sizeof(int)) + 64, 256);
    const int cs = r >= 0 ? r + 1 : 1, gw = (w + cs - 1) / cs, gh = (h + cs - 1) / cs;
    const size_t grid_b = align_up((size_t)gw * gh * sizeof(unsigned short), 256);
    const bool grid_in_smem = grid_b <= 180 * 1024;
    const size_t total = 3 * plane + val_b + 2 * keys_b + hist_b + map_b + feat_b + grid_b + EIG_BINS * 4 + 512;
    int rc = klt_ws_reserve(ctx, total);
    if (rc) return rc;
    char *wsp = (char *)ctx->ws;
    float *sxx = (float *)wsp; wsp += plane;
    float *sxy = (float *)wsp; wsp += plane;
    float *syy = (float *)wsp; wsp += plane;
    float *vmap = (float *)wsp; wsp += val_b;
    unsigned long long *keys0 = (unsigned long long *)wsp; wsp += keys_b;
    unsigned long long *keys1 = (unsigned long long *)wsp; wsp += keys_b;
    unsigned int *hist = (unsigned int *)wsp; wsp += hist_b;
    unsigned char *map = (unsigned char *)wsp; wsp += map_b;
    double *fx = (double *)wsp; double *fy = fx + n_features; int *fval = (int *)(fy + n_features); wsp += feat_b;
    unsigned short *grid_g = (unsigned short *)wsp; wsp += grid_b;
    unsigned int *ehist = (unsigned int *)wsp; wsp += EIG_BINS * 4;
    unsigned int *nkeys = (unsigned int *)wsp;                       // [0] key count, [1] threshold bin
    unsigned long long *consumed = (unsigned long long *)(wsp + 64); // [0] consumed, [1] ran out, [2] next slot

    KLT_CUDA(ctx, cudaMemsetAsync(ehist, 0, EIG_BINS * 4 + 512, ctx->stream));
    if (replace) {
        KLT_CUDA(ctx, cudaMemcpyAsync(fx, x, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemcpyAsync(fy, y, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemcpyAsync(fval, val, n_features * sizeof(int), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemsetAsync(map, 0, (size_t)w * h, ctx->stream));
        if (r >= 0 && n_features > 0)
            KLT_LAUNCH(ctx, "premark", 0.0, (premark_kernel<<<n_features, 128, 0, ctx->stream>>>(fx, fy, fval, n_features, map, w, h, r)));
    }
    if ((rc = launch_sat(ctx, gx, gy, pitch, w, h, sxx, sxy, syy))) return rc;
    if (ncand)
        KLT_LAUNCH(ctx, "eigen", 52.0 * ncand, (eigen_kernel<<<dim3((nx + 255) / 256, (ny + EIG_ROWS - 1) / EIG_ROWS), 256, 0, ctx->stream>>>(
                                                    sxx, sxy, syy, w, bx, by, hw, hh, step, nx, ny, vmap, ehist, (float)min_eig)));
    GreedyArgs G;
    G.nkeys = nkeys; G.premap = replace ? map : nullptr; G.grid_global = grid_g; G.W = w; G.H = h; G.r = r;
    G.n_features = n_features; G.overwrite = replace ? 0 : 1; G.cs = cs; G.gw = gw; G.gh = gh; G.grid_in_smem = grid_in_smem ? 1 : 0;
    G.fx = fx; G.fy = fy; G.fval = fval; G.free_slots = fval + n_features; G.consumed = consumed;
    if (grid_in_smem) KLT_CUDA(ctx, cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_b));
    unsigned long long cons[3] = {0, 0, 0};
    const unsigned int target = (unsigned int)(32u * (unsigned int)n_features + 8192u);
    for (int attempt = 0; attempt < 2; attempt++) {
        // attempt 0: only the best ~target candidates; attempt 1 (rare): every candidate >= min_eigenvalue
        unsigned int hk[2] = {0, 0};
        if (ncand) {
            KLT_CUDA(ctx, cudaMemsetAsync(nkeys, 0, 8, ctx->stream));
            KLT_LAUNCH(ctx, "threshold", 0.0, (threshold_kernel<<<1, 32, 0, ctx->stream>>>(ehist, target, attempt, nkeys + 1)));
            KLT_LAUNCH(ctx, "compact", 4.0 * ncand, (compact_kernel<<<dim3((nx + 255) / 256, ny), 256, 0, ctx->stream>>>(
                                                        vmap, bx, by, step, nx, ny, (float)min_eig, nkeys + 1, keys0, nkeys)));
            KLT_CUDA(ctx, cudaMemcpyAsync(hk, nkeys, 8, cudaMemcpyDeviceToHost, ctx->stream));
            KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));        // the sort's launch geometry follows the key count
        }
        const unsigned int nk = hk[0];
        const int nblocks = (int)((nk + RS_CHUNK - 1) / RS_CHUNK);
        unsigned long long *src = keys0, *dst = keys1;
        if (nblocks > 0) {
            // 58 significant key bits -> 8 passes of 8 bits
            for (int pass = 0; pass < 8; pass++) {
                KLT_LAUNCH(ctx, "rs_hist", 8.0 * nk, (rs_hist_kernel<<<nblocks, RS_THREADS, 0, ctx->stream>>>(src, nkeys, pass * 8, hist, nblocks)));
                if (nblocks <= RS_FUSED_SCAN_BLOCKS) {
                    KLT_LAUNCH(ctx, "rs_scatter", 16.0 * nk, (rs_scatter_kernel<false><<<nblocks, RS_THREADS, 0, ctx->stream>>>(src, dst, nkeys, pass * 8, hist, nblocks)));
                } else {
                    KLT_LAUNCH(ctx, "rs_scan", 0.0, (rs_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist, 256 * nblocks)));
                    KLT_LAUNCH(ctx, "rs_scatter", 16.0 * nk, (rs_scatter_kernel<true><<<nblocks, RS_THREADS, 0, ctx->stream>>>(src, dst, nkeys, pass * 8, hist, nblocks)));
                }
                unsigned long long *t = src; src = dst; dst = t;
            }
        }
        if (replace && attempt == 1) {      // the first walk may have filled slots: start again from the caller's list
            KLT_CUDA(ctx, cudaMemcpyAsync(fval, val, n_features * sizeof(int), cudaMemcpyDefault, ctx->stream));
            KLT_CUDA(ctx, cudaMemcpyAsync(fx, x, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
            KLT_CUDA(ctx, cudaMemcpyAsync(fy, y, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
        }
        G.keys = src;
        KLT_LAUNCH(ctx, "greedy", 0.0, (greedy_kernel<<<1, GREEDY_THREADS, grid_in_smem ? grid_b : 0, ctx->stream>>>(G)));
        KLT_CUDA(ctx, cudaMemcpyAsync(cons, consumed, sizeof(cons), cudaMemcpyDeviceToHost, ctx->stream));
        KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const bool thresholded = hk[1] > 0;
        if (!(cons[1] && thresholded)) break;   // all slots filled, or nothing was excluded: this IS the greedy result
    }
    if (!replace && cons[1] && (int)cons[2] < n_features) {
        const int rest = n_features - (int)cons[2];
        KLT_LAUNCH(ctx, "fill_not_found", 0.0, (fill_not_found_kernel<<<(rest + 127) / 128, 128, 0, ctx->stream>>>(fx, fy, fval, n_features, consumed)));
    }
    KLT_CUDA(ctx, cudaMemcpyAsync(x, fx, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(y, fy, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(val, fval, n_features * sizeof(int), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_consumed) *n_consumed = (int64_t)cons[0];
    return KLT_OK;
}

// _enforceMinimumDistance on a caller-provided, already ordered candidate list (selectGoodFeatures.py:45-135):
// keys_host[i] = ~((val_bits << 26) | (x << 13) | y) in walk order.
int klt_greedy_presorted(klt_ctx *ctx, const unsigned long long *keys_host, unsigned int nk, int w, int h, int mindist,
                         int n_features, int overwrite, double *x, double *y, int32_t *val) {
    if (w > 8191 || h > 8191) return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "image larger than 8191 pixels per side");
    if (mindist < 0) mindist = 0;
    const int r = mindist - 1;
    if (r > 254) return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "mindist larger than 255");
    const size_t keys_b = align_up(((size_t)nk + 1) * sizeof(unsigned long long), 256);
    const size_t map_b = align_up((size_t)w * h, 256);
    const size_t feat_b = align_up((size_t)n_features * (2 * sizeof(double) + 2 * sizeof(int)) + 64, 256);
    const int cs = r >= 0 ? r + 1 : 1, gw = (w + cs - 1) / cs, gh = (h + cs - 1) / cs;
    const size_t grid_b = align_up((size_t)gw * gh * sizeof(unsigned short), 256);
    const bool grid_in_smem = grid_b <= 180 * 1024;
    int rc = klt_ws_reserve(ctx, keys_b + map_b + feat_b + grid_b + 512);
    if (rc) return rc;
    char *wsp = (char *)ctx->ws;
    unsigned long long *keys = (unsigned long long *)wsp; wsp += keys_b;
    unsigned char *map = (unsigned char *)wsp; wsp += map_b;
    double *fx = (double *)wsp; double *fy = fx + n_features; int *fval = (int *)(fy + n_features); wsp += feat_b;
    unsigned short *grid_g = (unsigned short *)wsp; wsp += grid_b;
    unsigned int *nkeys = (unsigned int *)wsp;
    unsigned long long *consumed = (unsigned long long *)(wsp + 64);
    KLT_CUDA(ctx, cudaMemsetAsync(nkeys, 0, 512, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(nkeys, &nk, sizeof(nk), cudaMemcpyHostToDevice, ctx->stream));
    if (nk) KLT_CUDA(ctx, cudaMemcpyAsync(keys, keys_host, (size_t)nk * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
    if (!overwrite) {
        KLT_CUDA(ctx, cudaMemcpyAsync(fx, x, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemcpyAsync(fy, y, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemcpyAsync(fval, val, n_features * sizeof(int), cudaMemcpyDefault, ctx->stream));
        KLT_CUDA(ctx, cudaMemsetAsync(map, 0, (size_t)w * h, ctx->stream));
        if (r >= 0 && n_features > 0)
            KLT_LAUNCH(ctx, "premark", 0.0, (premark_kernel<<<n_features, 128, 0, ctx->stream>>>(fx, fy, fval, n_features, map, w, h, r)));
    }
    GreedyArgs G;
    G.keys = keys; G.nkeys = nkeys; G.premap = overwrite ? nullptr : map; G.grid_global = grid_g; G.W = w; G.H = h; G.r = r;
    G.n_features = n_features; G.overwrite = overwrite; G.cs = cs; G.gw = gw; G.gh = gh; G.grid_in_smem = grid_in_smem ? 1 : 0;
    G.fx = fx; G.fy = fy; G.fval = fval; G.free_slots = fval + n_features; G.consumed = consumed;
    if (grid_in_smem) KLT_CUDA(ctx, cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)grid_b));
    KLT_LAUNCH(ctx, "greedy", 0.0, (greedy_kernel<<<1, GREEDY_THREADS, grid_in_smem ? grid_b : 0, ctx->stream>>>(G)));
    unsigned long long cons[3] = {0, 0, 0};
    KLT_CUDA(ctx, cudaMemcpyAsync(cons, consumed, sizeof(cons), cudaMemcpyDeviceToHost, ctx->stream));
    KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (overwrite && cons[1] && (int)cons[2] < n_features) {
        const int rest = n_features - (int)cons[2];
        KLT_LAUNCH(ctx, "fill_not_found", 0.0, (fill_not_found_kernel<<<(rest + 127) / 128, 128, 0, ctx->stream>>>(fx, fy, fval, n_features, consumed)));
    }
    KLT_CUDA(ctx, cudaMemcpyAsync(x, fx, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(y, fy, n_features * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(val, fval, n_features * sizeof(int), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KLT_OK;
}
