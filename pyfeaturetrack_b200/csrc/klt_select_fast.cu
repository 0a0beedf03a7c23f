// FAST eigenvalue map: gradients, window sums and the minimum eigenvalue from ONE pass over the level-0 image
// (north_star (2): "the per-pixel 2x2 gradient-matrix minimum-eigenvalue scan fused with the gradient pass").
//
// Replaces, for select="fast", KLTComputeGradients (convolve.py:226-248) + ScanImageForGoodFeatures
// (goodFeaturesUtils.pyx:35-73): the gradient planes and the three summed-area tables never exist.  Arithmetic is plain
// fp32 with direct window sums, so the values differ from the reference's SAT-rounded ones by the SAT's own rounding error
// (same feature SET for ~99.8 %, not the same slots) -- the mode is judged by set overlap and labelled as such.
//
// Shape: a block of 128 threads owns a strip of 128 columns of one image and marches down a row segment; a thread owns one
// column.  Per input row: (1) the row goes through a shared-memory line, every thread forms d_h(I) and g_h(I) from its 7
// neighbours; (2) the vertical 7-tap filters run as accumulate-and-shift FMAs in registers (no memory), yielding one row of
// gx, gy; (3) gx^2, gx*gy, gy^2 enter the vertical window sums, again accumulate-and-shift in registers; (4) the completed row
// of column sums goes through shared memory once for the horizontal window sum; (5) eigenvalue, map store, histogram.
// Compulsory traffic: 4 B/px read + 4 B/candidate written (the strict path moves 52 B/px).
#include "klt_common.cuh"
#include "klt_select.cuh"

#define FS_THREADS 128
#define FS_R 3                  // gradient kernel radius (grad_sigma = 1.0 -> 7 taps)
#define FS_PF 4                 // rows loaded ahead

struct FastTaps { float g[7], d[7]; };   // c[j] multiplies in[x + j - 3]

template <int HH>
__global__ void __launch_bounds__(FS_THREADS)
eigen_fast_kernel(const float *__restrict__ img0, size_t img_stride, size_t pitch, const __grid_constant__ SelDev S,
                  const __grid_constant__ FastTaps T, int rows_per_seg) {
    __shared__ float rowI[2][FS_THREADS];
    __shared__ float rowV[2][3][FS_THREADS];
    __shared__ unsigned int h[SEL_BINS];
    const int t = threadIdx.x, b = blockIdx.z;
    const int W = S.W, H = S.H, hw = S.hw;
    const int halo = FS_R + hw, span = FS_THREADS - 2 * halo;
    const int x = S.bx - halo + (int)blockIdx.x * span + t;            // this thread's column
    const int xr = klt_reflect(x, W);
    const float *img = img0 + (size_t)b * img_stride + xr;
    for (int k = t; k < SEL_BINS; k += FS_THREADS) h[k] = 0;
    // candidate rows of this segment
    const int ys = S.by + (int)blockIdx.y * rows_per_seg, ye = min(ys + rows_per_seg, H - S.by);
    if (ys >= ye) return;
    const int ci = x - S.bx;
    const bool cand_col = t >= halo && t < FS_THREADS - halo && x < W - S.bx && (ci % S.step) == 0;
    const int i_c = ci / S.step;
    float *vmap = S.vmap + (size_t)b * S.ncand;
    const unsigned char *pm = S.premap ? S.premap + (size_t)b * S.map_stride : nullptr;
    float Px[6], Py[6];                   // pending gx / gy rows (vertical 7-tap filters)
    float Qxx[2 * HH], Qxy[2 * HH], Qyy[2 * HH];   // pending vertical window sums
#pragma unroll
    for (int m = 0; m < 6; m++) { Px[m] = 0.f; Py[m] = 0.f; }
#pragma unroll
    for (int m = 0; m < 2 * HH; m++) { Qxx[m] = 0.f; Qxy[m] = 0.f; Qyy[m] = 0.f; }
    const int y0 = ys - HH - FS_R;        // first input row; input row y completes gradient row y-3 and window row y-3-HH
    float q[FS_PF];
#pragma unroll
    for (int i = 0; i < FS_PF; i++) q[i] = img[(size_t)klt_reflect(y0 + i, H) * pitch];
    __syncthreads();
    const int y_end = ye - 1 + HH + FS_R;  // last input row (inclusive)
    for (int y = y0; y <= y_end; y++) {
        const int par = (y - y0) & 1;
        const float cur = q[0];
#pragma unroll
        for (int i = 0; i < FS_PF - 1; i++) q[i] = q[i + 1];
        q[FS_PF - 1] = img[(size_t)klt_reflect(min(y + FS_PF, y_end), H) * pitch];
        rowI[par][t] = cur;
        __syncthreads();
        float dh = 0.f, gh = 0.f;
        if (t >= FS_R && t < FS_THREADS - FS_R) {
            const float *r = &rowI[par][t];
            const float a1 = r[1] + r[-1], a2 = r[2] + r[-2], a3 = r[3] + r[-3];
            const float b1 = r[1] - r[-1], b2 = r[2] - r[-2], b3 = r[3] - r[-3];
            // the reference's kernels are symmetric (gauss) / antisymmetric (derivative): c[3+k] = +-c[3-k]
            gh = fmaf(T.g[6], a3, fmaf(T.g[5], a2, fmaf(T.g[4], a1, T.g[3] * r[0])));
            dh = fmaf(T.d[6], b3, fmaf(T.d[5], b2, T.d[4] * b1));
        }
        // vertical filters: gx = g_v(d_h), gy = d_v(g_h); row y completes output row y - 3
        const float gx = fmaf(T.g[6], dh, Px[0]), gy = fmaf(T.d[6], gh, Py[0]);
#pragma unroll
        for (int m = 0; m < 5; m++) { Px[m] = fmaf(T.g[5 - m], dh, Px[m + 1]); Py[m] = fmaf(T.d[5 - m], gh, Py[m + 1]); }
        Px[5] = T.g[0] * dh; Py[5] = T.d[0] * gh;
        // vertical window sums: gradient row y-3 completes window row y - 3 - HH
        const float pxx = gx * gx, pxy = gx * gy, pyy = gy * gy;
        const float vxx = Qxx[0] + pxx, vxy = Qxy[0] + pxy, vyy = Qyy[0] + pyy;
#pragma unroll
        for (int m = 0; m < 2 * HH - 1; m++) { Qxx[m] = Qxx[m + 1] + pxx; Qxy[m] = Qxy[m + 1] + pxy; Qyy[m] = Qyy[m + 1] + pyy; }
        Qxx[2 * HH - 1] = pxx; Qxy[2 * HH - 1] = pxy; Qyy[2 * HH - 1] = pyy;
        const int yc = y - FS_R - HH;
        if (yc >= ys) {                                   // block-uniform
            rowV[par][0][t] = vxx; rowV[par][1][t] = vxy; rowV[par][2][t] = vyy;
            __syncthreads();
            const int cj = yc - S.by;
            if (cand_col && (cj % S.step) == 0) {
                float gxx = 0.f, gxy = 0.f, gyy = 0.f;
                for (int k = -hw; k <= hw; k++) { gxx += rowV[par][0][t + k]; gxy += rowV[par][1][t + k]; gyy += rowV[par][2][t + k]; }
                const float dd = gxx - gyy;
                const float v = 0.5f * ((gxx + gyy) - sqrtf(fmaf(dd, dd, 4.f * gxy * gxy)));
                vmap[(size_t)(cj / S.step) * S.nx + i_c] = v;
                if (v >= S.min_val && !(pm && pm[(size_t)yc * W + x])) atomicAdd(&h[eig_rbin(v)], 1u);
            }
        }
    }
    __syncthreads();
    unsigned int *hist = S.hist + (size_t)b * SEL_BINS;
    for (int k = t; k < SEL_BINS; k += FS_THREADS)
        if (h[k]) atomicAdd(&hist[k], h[k]);
}

// ---- windows up to 7x7: the same pass with no shared memory and no barriers -------------------------------------------
// A WARP owns 128 adjacent columns and marches down its row segment; a lane owns 4 adjacent columns (one 128-bit load per
// row).  Horizontal neighbours -- three image columns for the 7-tap gradient filters, HH columns of window sums -- come from
// the adjacent lanes by shuffle; lanes 0, 1, 30, 31 are halo lanes (112 of 128 columns produce output).  Everything else is
// register arithmetic on four independent columns, so the kernel needs neither occupancy nor barriers to stay busy.
#define FQ_PF 3
template <int HH>
__global__ void __launch_bounds__(FS_THREADS)
eigen_fast_quad_kernel(const float *__restrict__ img0, size_t img_stride, size_t pitch, const __grid_constant__ SelDev S,
                       const __grid_constant__ FastTaps T, int rows_per_seg, int n_strips, int vec_ok) {
    __shared__ unsigned int h[SEL_BINS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, b = blockIdx.z;
    const int W = S.W, H = S.H;
    for (int k = t; k < SEL_BINS; k += FS_THREADS) h[k] = 0;
    __syncthreads();
    const int strip = blockIdx.x * (FS_THREADS / 32) + warp;
    const int ys = S.by + (int)blockIdx.y * rows_per_seg, ye = min(ys + rows_per_seg, H - S.by);
    if (strip < n_strips && ys < ye) {
        const int xa = S.bx & ~3;
        const int xl = xa - 8 + strip * 112 + 4 * lane;             // first of this lane's 4 columns (multiple of 4)
        const bool inside = xl >= 0 && xl + 3 < W;
        const bool vec = inside && vec_ok;
        const int c0 = klt_reflect(xl, W), c1 = klt_reflect(xl + 1, W), c2 = klt_reflect(xl + 2, W), c3 = klt_reflect(xl + 3, W);
        const float *img = img0 + (size_t)b * img_stride;
        auto load_row = [&](int y) -> float4 {
            const float *r = img + (size_t)klt_reflect(y, H) * pitch;
            if (vec) return *reinterpret_cast<const float4 *>(r + xl);
            return make_float4(r[c0], r[c1], r[c2], r[c3]);
        };
        const bool out_lane = lane >= 2 && lane < 30;
        float *vmap = S.vmap + (size_t)b * S.ncand;
        const unsigned char *pm = S.premap ? S.premap + (size_t)b * S.map_stride : nullptr;
        float Px[4][6], Py[4][6], Qxx[4][2 * HH], Qxy[4][2 * HH], Qyy[4][2 * HH];
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int m = 0; m < 6; m++) { Px[c][m] = 0.f; Py[c][m] = 0.f; }
#pragma unroll
            for (int m = 0; m < 2 * HH; m++) { Qxx[c][m] = 0.f; Qxy[c][m] = 0.f; Qyy[c][m] = 0.f; }
        }
        const int y0 = ys - HH - FS_R, y_end = ye - 1 + HH + FS_R;
        float4 q[FQ_PF];
#pragma unroll
        for (int i = 0; i < FQ_PF; i++) q[i] = load_row(min(y0 + i, y_end));
        for (int y = y0; y <= y_end; y++) {
            const float4 cur = q[0];
#pragma unroll
            for (int i = 0; i < FQ_PF - 1; i++) q[i] = q[i + 1];
            q[FQ_PF - 1] = load_row(min(y + FQ_PF, y_end));
            float I[10];                                            // columns xl-3 .. xl+6
            I[0] = __shfl_up_sync(0xffffffffu, cur.y, 1); I[1] = __shfl_up_sync(0xffffffffu, cur.z, 1); I[2] = __shfl_up_sync(0xffffffffu, cur.w, 1);
            I[3] = cur.x; I[4] = cur.y; I[5] = cur.z; I[6] = cur.w;
            I[7] = __shfl_down_sync(0xffffffffu, cur.x, 1); I[8] = __shfl_down_sync(0xffffffffu, cur.y, 1); I[9] = __shfl_down_sync(0xffffffffu, cur.z, 1);
            float vxx[4], vxy[4], vyy[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float *r = &I[3 + c];
                const float a1 = r[1] + r[-1], a2 = r[2] + r[-2], a3 = r[3] + r[-3];
                const float b1 = r[1] - r[-1], b2 = r[2] - r[-2], b3 = r[3] - r[-3];
                const float gh = fmaf(T.g[6], a3, fmaf(T.g[5], a2, fmaf(T.g[4], a1, T.g[3] * r[0])));
                const float dh = fmaf(T.d[6], b3, fmaf(T.d[5], b2, T.d[4] * b1));
                const float gx = fmaf(T.g[6], dh, Px[c][0]), gy = fmaf(T.d[6], gh, Py[c][0]);
#pragma unroll
                for (int m = 0; m < 5; m++) { Px[c][m] = fmaf(T.g[5 - m], dh, Px[c][m + 1]); Py[c][m] = fmaf(T.d[5 - m], gh, Py[c][m + 1]); }
                Px[c][5] = T.g[0] * dh; Py[c][5] = T.d[0] * gh;
                const float pxx = gx * gx, pxy = gx * gy, pyy = gy * gy;
                vxx[c] = Qxx[c][0] + pxx; vxy[c] = Qxy[c][0] + pxy; vyy[c] = Qyy[c][0] + pyy;
#pragma unroll
                for (int m = 0; m < 2 * HH - 1; m++) { Qxx[c][m] = Qxx[c][m + 1] + pxx; Qxy[c][m] = Qxy[c][m + 1] + pxy; Qyy[c][m] = Qyy[c][m + 1] + pyy; }
                Qxx[c][2 * HH - 1] = pxx; Qxy[c][2 * HH - 1] = pxy; Qyy[c][2 * HH - 1] = pyy;
            }
            const int yc = y - FS_R - HH;
            if (yc >= ys) {                                          // warp-uniform
                // horizontal window sums: columns xl-HH .. xl+3+HH of the three planes
                float sx[4], sxy_[4], sy[4];
                auto hbox = [&](const float (&v)[4], float (&o)[4]) {
                    float V[4 + 2 * HH];
#pragma unroll
                    for (int k = 0; k < HH; k++) {
                        V[k] = __shfl_up_sync(0xffffffffu, v[4 - HH + k], 1);
                        V[4 + HH + k] = __shfl_down_sync(0xffffffffu, v[k], 1);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) V[HH + k] = v[k];
                    float s0 = V[0];
#pragma unroll
                    for (int k = 1; k <= 2 * HH; k++) s0 += V[k];
                    o[0] = s0;
#pragma unroll
                    for (int c = 1; c < 4; c++) o[c] = o[c - 1] + V[c + 2 * HH] - V[c - 1];
                };
                hbox(vxx, sx); hbox(vxy, sxy_); hbox(vyy, sy);
                const int cj = yc - S.by;
                if (out_lane && (cj % S.step) == 0) {
                    float *vrow = vmap + (size_t)(cj / S.step) * S.nx;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const int x = xl + c, ci = x - S.bx;
                        if (ci >= 0 && x < W - S.bx && (ci % S.step) == 0) {
                            const float dd = sx[c] - sy[c];
                            const float v = 0.5f * ((sx[c] + sy[c]) - sqrtf(fmaf(dd, dd, 4.f * sxy_[c] * sxy_[c])));
                            vrow[ci / S.step] = v;
                            if (v >= S.min_val && !(pm && pm[(size_t)yc * W + x])) atomicAdd(&h[eig_rbin(v)], 1u);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    unsigned int *hist = S.hist + (size_t)b * SEL_BINS;
    for (int k = t; k < SEL_BINS; k += FS_THREADS)
        if (h[k]) atomicAdd(&hist[k], h[k]);
}

template <int HH>
static int launch_fast_quad(klt_ctx *ctx, const SelDev *S, int B, const float *img0, size_t img_stride, size_t pitch, const FastTaps &T) {
    const int ncols = S->W - S->bx - (S->bx & ~3), nrows = S->H - 2 * S->by;
    if (S->W - 2 * S->bx <= 0 || nrows <= 0) return 1;
    const int n_strips = (ncols + 111) / 112;
    const int strip_blocks = (n_strips + FS_THREADS / 32 - 1) / (FS_THREADS / 32);
    // about two blocks per SM; a segment re-reads 2*(HH+3) warm-up rows, so keep segments >= 64 rows
    long nseg = ((long)ctx->num_sms * 2) / ((long)strip_blocks * B);
    if (nseg < 1) nseg = 1;
    int rows = (int)((nrows + nseg - 1) / nseg);
    if (rows < 64) rows = 64;
    if (rows > nrows) rows = nrows;
    const dim3 grid(strip_blocks, (nrows + rows - 1) / rows, B);
    const double bytes = (4.0 * S->W * S->H + 4.0 * S->ncand) * B;
    const int vec_ok = (pitch % 4) == 0 && (img_stride % 4) == 0 && (reinterpret_cast<uintptr_t>(img0) & 15) == 0;
    KLT_LAUNCH(ctx, "eigen_fast", bytes, (eigen_fast_quad_kernel<HH><<<grid, FS_THREADS, 0, ctx->stream>>>(img0, img_stride, pitch, *S, T, rows, n_strips, vec_ok)));
    return 1;
}

template <int HH>
static int launch_fast(klt_ctx *ctx, const SelDev *S, int B, const float *img0, size_t img_stride, size_t pitch, const FastTaps &T) {
    const int span = FS_THREADS - 2 * (FS_R + S->hw);
    const int ncols = S->W - 2 * S->bx, nrows = S->H - 2 * S->by;
    if (ncols <= 0 || nrows <= 0) return 1;
    const int strips = (ncols + span - 1) / span;
    // about two waves of 8 resident blocks per SM; a segment re-reads 2*(HH+3) warm-up rows, so keep segments >= 64 rows
    long nseg = ((long)ctx->num_sms * 16) / ((long)strips * B);
    if (nseg < 1) nseg = 1;
    int rows = (int)((nrows + nseg - 1) / nseg);
    if (rows < 64) rows = 64;
    if (rows > nrows) rows = nrows;
    const dim3 grid(strips, (nrows + rows - 1) / rows, B);
    const double bytes = (4.0 * S->W * S->H + 4.0 * S->ncand) * B;
    KLT_LAUNCH(ctx, "eigen_fast", bytes, (eigen_fast_kernel<HH><<<grid, FS_THREADS, 0, ctx->stream>>>(img0, img_stride, pitch, *S, T, rows)));
    return 1;
}

// 1 = launched, 0 = configuration not covered (caller builds gradient planes and takes the table-based pass), < 0 error
int klt_sel_launch_eigen_fast(klt_ctx *ctx, const SelDev *S, int B, const float *img0, size_t img_stride, size_t pitch,
                              const klt_kernel1d *gauss, const klt_kernel1d *deriv) {
    if (!gauss || !deriv || gauss->n != 2 * FS_R + 1 || deriv->n != 2 * FS_R + 1) return 0;
    if (S->hw != S->hh || S->hw < 1 || S->hw > 7) return 0;
    FastTaps T;
    for (int j = 0; j < 7; j++) { T.g[j] = (float)gauss->taps[6 - j]; T.d[j] = (float)deriv->taps[6 - j]; }
    for (int k = 1; k <= 3; k++)            // the folded form needs the symmetry the reference's kernels have
        if (fabs(gauss->taps[3 + k] - gauss->taps[3 - k]) > 1e-12 || fabs(deriv->taps[3 + k] + deriv->taps[3 - k]) > 1e-12) return 0;
    if (fabs(deriv->taps[3]) > 1e-12) return 0;
    switch (S->hh) {
        case 1: return launch_fast_quad<1>(ctx, S, B, img0, img_stride, pitch, T);
        case 2: return launch_fast_quad<2>(ctx, S, B, img0, img_stride, pitch, T);
        case 3: return launch_fast_quad<3>(ctx, S, B, img0, img_stride, pitch, T);
        case 4: return launch_fast<4>(ctx, S, B, img0, img_stride, pitch, T);
        case 5: return launch_fast<5>(ctx, S, B, img0, img_stride, pitch, T);
        case 6: return launch_fast<6>(ctx, S, B, img0, img_stride, pitch, T);
        case 7: return launch_fast<7>(ctx, S, B, img0, img_stride, pitch, T);
    }
    return 0;
}
