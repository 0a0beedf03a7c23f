// Affine consistency check of KLTTrackFeatures (trackFeatures.py:347-399) on the GPU (sm_100a).
//
// The reference's call site exists but its callees (_KLTCreateFloatImage, _am_getSubFloatImage,
// _am_trackFeatureAffine) are undefined -- it dies with NameError on the first tracked feature.  This kernel
// implements the routines of those names from Birchfield's C KLT 1.3.4 (trackFeatures.c), which the call site
// mirrors argument for argument.  Parity is therefore checked against the in-repo
// restatement (oracle/klt_oracle.c: orc_affine_step), NOT against the reference: "parity unpinned".
//
// Mapping: one warp per feature; the 15x15 window's 225 pixels are spread over the 32 lanes (8 per lane).  Every
// lane accumulates its share of the 21 + 6 sums of the 6x6 normal equations (10 + 4 for the similarity model, 5 for
// pure translation), a butterfly of warp shuffles leaves the totals in every lane, and the warp solves the system
// cooperatively: lane r owns row r of the Gauss-Jordan elimination with full pivoting (registers and shuffles only).
// Per-feature state (template of (aw+2)x(ah+2) pixels for image / gradx / grady, template centre, the 2x2 map A)
// lives in device memory inside a klt_affine object and persists from frame to frame like the fields of KLT_Feature.
#include "klt_common.cuh"

struct AffineArgs {
    klt_pyr p1, p2;
    int n_per_image, total;
    int affine_map, width, height, max_iterations;
    float step_factor, small_det, th, th_aff, max_residue, mdd;
    int tw, th_, tn;                  // template dims (aw+2, ah+2) and pixel count
    int *has;
    float *aff_x, *aff_y, *A, *tmpl;
};

__device__ __forceinline__ float interp_f(const float *__restrict__ img, int nc, float x, float y) {
    const int xt = (int)x, yt = (int)y;
    const float ax = x - xt, ay = y - yt;
    const float *p = img + (size_t)nc * yt + xt;
    return (1.f - ax) * (1.f - ay) * p[0] + ax * (1.f - ay) * p[1] + (1.f - ax) * ay * p[nc] + ax * ay * p[nc + 1];
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// C-KLT _am_gauss_jordan_elimination (Numerical Recipes gaussj with full pivoting); b is the single right-hand side.
// Warp-cooperative and register-only: lane r < 6 owns row r of the augmented matrix (rw[0..5] | rb); the column of an
// access is the only dynamic index left and is resolved by a 6-way select, so nothing goes to local memory (the earlier
// per-lane version indexed a[6][6] dynamically: 877 LDL/STL in its SASS).  Pivot search = per-row maximum, then a
// butterfly arg-max over the rows; ties resolve to the LAST element in (row, column) order like the reference's `>=` scan.
// The solution comes back in every lane's x[0..5].
__device__ __forceinline__ float sel6(const float (&v)[6], int k) {
    return k == 0 ? v[0] : (k == 1 ? v[1] : (k == 2 ? v[2] : (k == 3 ? v[3] : (k == 4 ? v[4] : v[5]))));
}
__device__ int gauss_jordan_warp(const float (&T)[6][6], const float (&rhs)[6], int n, int lane, float (&x)[6]) {
    const unsigned int full = 0xffffffffu;
    float rw[6], rb = sel6(rhs, lane < 6 ? lane : 0);
#pragma unroll
    for (int c = 0; c < 6; c++) {
        const float col[6] = {T[0][c], T[1][c], T[2][c], T[3][c], T[4][c], T[5][c]};
        rw[c] = sel6(col, lane < 6 ? lane : 0);
    }
    unsigned int used = 0u;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (i < n) {                                            // warp-uniform
            float best = -1.0f;
            int bk = 0;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (k < n && !((used >> k) & 1u)) { const float v = fabsf(rw[k]); if (v >= best) { best = v; bk = k; } }
            if (!(lane < n) || ((used >> lane) & 1u)) best = -2.0f;      // rows already used as pivots (ipiv[j] == 1) and idle lanes
            int br = lane;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {                   // arg-max over lanes 0..7
                const float ob = __shfl_xor_sync(full, best, o);
                const int orow = __shfl_xor_sync(full, br, o), ok = __shfl_xor_sync(full, bk, o);
                if (ob > best || (ob == best && orow > br)) { best = ob; br = orow; bk = ok; }
            }
            const int row = __shfl_sync(full, br, 0), col = __shfl_sync(full, bk, 0);
            used |= 1u << col;
            if (row != col) {                                   // swap rows `row` and `col`
                const int src = lane == row ? col : (lane == col ? row : lane);
#pragma unroll
                for (int l = 0; l < 6; l++) rw[l] = __shfl_sync(full, rw[l], src);
                rb = __shfl_sync(full, rb, src);
            }
            const float piv = __shfl_sync(full, sel6(rw, col), col);
            if (piv == 0.0f) return KLT_SMALL_DET;
            const float pivinv = 1.0f / piv;
            if (lane == col) {
#pragma unroll
                for (int l = 0; l < 6; l++) rw[l] = (l == col ? 1.0f : rw[l]) * pivinv;
                rb *= pivinv;
            }
            float pr[6];
#pragma unroll
            for (int l = 0; l < 6; l++) pr[l] = __shfl_sync(full, rw[l], col);
            const float pb = __shfl_sync(full, rb, col);
            if (lane != col) {
                const float dum = sel6(rw, col);
#pragma unroll
                for (int l = 0; l < 6; l++) rw[l] = (l == col ? 0.0f : rw[l]) - pr[l] * dum;
                rb -= pb * dum;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 6; r++) x[r] = __shfl_sync(full, rb, r);
    return KLT_TRACKED;
}

__device__ __forceinline__ bool oob1(float v, int n) { return v < 0.0f || (float)n - v < 1.001f; }

__global__ void __launch_bounds__(128, 3)
lk_affine_kernel(const __grid_constant__ AffineArgs G, const double *__restrict__ x_in, const double *__restrict__ y_in,
                 const int *__restrict__ val_in, double *__restrict__ xs, double *__restrict__ ys, int *__restrict__ vals,
                 int *__restrict__ assert_flag) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * (blockDim.x >> 5) + warp;
    if (f >= G.total) return;
    if (val_in[f] < 0) return;                                   // not tracked in this call (trackFeatures.py:253)
    if (vals[f] != KLT_TRACKED) {                                // lost by the translational tracker: templates freed
        if (lane == 0) G.has[f] = 0;
        return;
    }
    const int image = f / G.n_per_image;
    const int nc = G.p1.lv[0].w, nr = G.p1.lv[0].h, pitch = G.p1.lv[0].pitch;
    float *tm = G.tmpl + (size_t)f * 3 * G.tn;
    if (!G.has[f]) {
        // first successful track: store the integer-aligned template of image 1 (_am_getSubFloatImage)
        const float xl = (float)x_in[f], yl = (float)y_in[f];
        const int hw = G.tw / 2, hh = G.th_ / 2, x0 = (int)xl, y0 = (int)yl;
        if (!(x0 - hw >= 0 && y0 - hh >= 0 && x0 + hw <= nc && y0 + hh <= nr)) {
            if (lane == 0) atomicExch(assert_flag, 1);
            return;
        }
        for (int c = 0; c < 3; c++) {
            const float *src = G.p1.level(c, image, 0);
            for (int k = lane; k < G.tn; k += 32) {
                const int j = k / G.tw - hh, i = k % G.tw - hw;
                tm[c * G.tn + k] = src[(size_t)(j + y0) * pitch + (i + x0)];
            }
        }
        if (lane == 0) {
            G.aff_x[f] = xl - (float)x0 + (float)((G.width + 2) / 2);
            G.aff_y[f] = yl - (float)y0 + (float)((G.height + 2) / 2);
            G.has[f] = 1;
        }
        return;
    }
    // ---- _am_trackFeatureAffine ----
    const float *I2 = G.p2.level(0, image, 0), *GX2 = G.p2.level(1, image, 0), *GY2 = G.p2.level(2, image, 0);
    const float *T1 = tm, *TX1 = tm + G.tn, *TY1 = tm + 2 * G.tn;
    const int nc1 = G.tw, nr1 = G.th_;
    const int width = G.width, height = G.height, hw = width / 2, hh = height / 2, npx = width * height;
    const float x1 = G.aff_x[f], y1 = G.aff_y[f];
    float x2 = (float)xs[f], y2 = (float)ys[f];
    const float old_x2 = x2, old_y2 = y2;
    float Axx = G.A[4 * f + 0], Ayx = G.A[4 * f + 1], Axy = G.A[4 * f + 2], Ayy = G.A[4 * f + 3];
    int status = KLT_TRACKED, iteration = 0;
    bool convergence = false;
    const bool tmpl_oob = x1 - hw < 0.0f || nc1 - (x1 + hw) < 1.001f || y1 - hh < 0.0f || nr1 - (y1 + hh) < 1.001f;
    do {
        float dx = 0.f, dy = 0.f;
        if (G.affine_map == 0) {
            if (tmpl_oob || x2 - hw < 0.0f || nc - (x2 + hw) < 1.001f || y2 - hh < 0.0f || nr - (y2 + hh) < 1.001f) { status = KLT_OOB; break; }
            float gxx = 0, gxy = 0, gyy = 0, ex = 0, ey = 0;
            for (int k = lane; k < npx; k += 32) {
                const int j = k / width - hh, i = k % width - hw;
                const float diff = interp_f(T1, nc1, x1 + i, y1 + j) - interp_f(I2, pitch, x2 + i, y2 + j);
                const float gx = interp_f(TX1, nc1, x1 + i, y1 + j) + interp_f(GX2, pitch, x2 + i, y2 + j);
                const float gy = interp_f(TY1, nc1, x1 + i, y1 + j) + interp_f(GY2, pitch, x2 + i, y2 + j);
                gxx += gx * gx; gxy += gx * gy; gyy += gy * gy; ex += diff * gx; ey += diff * gy;
            }
            gxx = warp_sum(gxx); gxy = warp_sum(gxy); gyy = warp_sum(gyy);
            ex = warp_sum(ex) * G.step_factor; ey = warp_sum(ey) * G.step_factor;
            const float det = gxx * gyy - gxy * gxy;
            if (det < G.small_det) status = KLT_SMALL_DET;
            else { dx = (gyy * ex - gxy * ey) / det; dy = (gxx * ey - gxy * ex) / det; status = KLT_TRACKED; }
            convergence = fabsf(dx) < G.th && fabsf(dy) < G.th;
            x2 += dx; y2 += dy;
        } else {
            float ul_x = Axx * (-hw) + Axy * hh + x2, ul_y = Ayx * (-hw) + Ayy * hh + y2;
            float ll_x = Axx * (-hw) + Axy * (-hh) + x2, ll_y = Ayx * (-hw) + Ayy * (-hh) + y2;
            float ur_x = Axx * hw + Axy * hh + x2, ur_y = Ayx * hw + Ayy * hh + y2;
            float lr_x = Axx * hw + Axy * (-hh) + x2, lr_y = Ayx * hw + Ayy * (-hh) + y2;
            if (tmpl_oob || oob1(ul_x, nc) || oob1(ll_x, nc) || oob1(ur_x, nc) || oob1(lr_x, nc) || oob1(ul_y, nr) ||
                oob1(ll_y, nr) || oob1(ur_y, nr) || oob1(lr_y, nr)) { status = KLT_OOB; break; }
            float T[6][6], a[6];
#pragma unroll
            for (int r = 0; r < 6; r++) { a[r] = 0.f;
#pragma unroll
                for (int c = 0; c < 6; c++) T[r][c] = 0.f; }
            for (int k = lane; k < npx; k += 32) {
                const int jj = k / width - hh, ii = k % width - hw;
                const float x = (float)ii, y = (float)jj;
                const float mx = x2 + (Axx * x + Axy * y), my = y2 + (Ayx * x + Ayy * y);
                const float diff = interp_f(T1, nc1, x1 + x, y1 + y) - interp_f(I2, pitch, mx, my);
                const float gx = interp_f(GX2, pitch, mx, my), gy = interp_f(GY2, pitch, mx, my);
                if (G.affine_map == 1) {
                    const float t1 = x * gx + y * gy, t2 = x * gy - y * gx;
                    a[0] += diff * t1; a[1] += diff * t2; a[2] += diff * gx; a[3] += diff * gy;
                    T[0][0] += t1 * t1; T[0][1] += t1 * t2; T[0][2] += gx * t1; T[0][3] += gy * t1;
                    T[1][1] += t2 * t2; T[1][2] += gx * t2; T[1][3] += gy * t2;
                    T[2][2] += gx * gx; T[2][3] += gx * gy; T[3][3] += gy * gy;
                } else {
                    const float gxx = gx * gx, gxy = gx * gy, gyy = gy * gy, xx = x * x, xy = x * y, yy = y * y;
                    const float dgx = diff * gx, dgy = diff * gy;
                    a[0] += dgx * x; a[1] += dgy * x; a[2] += dgx * y; a[3] += dgy * y; a[4] += dgx; a[5] += dgy;
                    T[0][0] += xx * gxx; T[0][1] += xx * gxy; T[0][2] += xy * gxx; T[0][3] += xy * gxy; T[0][4] += x * gxx; T[0][5] += x * gxy;
                    T[1][1] += xx * gyy; T[1][2] += xy * gxy; T[1][3] += xy * gyy; T[1][4] += x * gxy; T[1][5] += x * gyy;
                    T[2][2] += yy * gxx; T[2][3] += yy * gxy; T[2][4] += y * gxx; T[2][5] += y * gxy;
                    T[3][3] += yy * gyy; T[3][4] += y * gxy; T[3][5] += y * gyy;
                    T[4][4] += gxx; T[4][5] += gxy; T[5][5] += gyy;
                }
            }
            const int n = G.affine_map == 1 ? 4 : 6;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                a[r] = warp_sum(a[r]) * 0.5f;
#pragma unroll
                for (int c = r; c < 6; c++) T[r][c] = warp_sum(T[r][c]);
            }
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < r; c++) T[r][c] = T[c][r];
            {
                float sol[6];
                status = gauss_jordan_warp(T, a, n, lane, sol);
#pragma unroll
                for (int r = 0; r < 6; r++) a[r] = sol[r];
            }
            if (G.affine_map == 1) { Axx += a[0]; Ayx += a[1]; Ayy = Axx; Axy = -Ayx; dx = a[2]; dy = a[3]; }
            else { Axx += a[0]; Ayx += a[1]; Axy += a[2]; Ayy += a[3]; dx = a[4]; dy = a[5]; }
            x2 += dx; y2 += dy;
            ul_x -= Axx * (-hw) + Axy * hh + x2; ul_y -= Ayx * (-hw) + Ayy * hh + y2;
            ll_x -= Axx * (-hw) + Axy * (-hh) + x2; ll_y -= Ayx * (-hw) + Ayy * (-hh) + y2;
            ur_x -= Axx * hw + Axy * hh + x2; ur_y -= Ayx * hw + Ayy * hh + y2;
            lr_x -= Axx * hw + Axy * (-hh) + x2; lr_y -= Ayx * hw + Ayy * (-hh) + y2;
            convergence = fabsf(dx) < G.th && fabsf(dy) < G.th && fabsf(ul_x) < G.th_aff && fabsf(ul_y) < G.th_aff &&
                          fabsf(ll_x) < G.th_aff && fabsf(ll_y) < G.th_aff && fabsf(ur_x) < G.th_aff &&
                          fabsf(ur_y) < G.th_aff && fabsf(lr_x) < G.th_aff && fabsf(lr_y) < G.th_aff;
        }
        if (status == KLT_SMALL_DET) break;
        iteration++;
    } while (!convergence && iteration < G.max_iterations);

    if (x2 - hw < 0.0f || nc - (x2 + hw) < 1.001f || y2 - hh < 0.0f || nr - (y2 + hh) < 1.001f) status = KLT_OOB;
    if ((x2 - old_x2) > G.mdd || (y2 - old_y2) > G.mdd) status = KLT_OOB;
    if (status == KLT_TRACKED) {
        float sum = 0.f;
        for (int k = lane; k < npx; k += 32) {
            const int jj = k / width - hh, ii = k % width - hw;
            const float g1 = interp_f(T1, nc1, x1 + ii, y1 + jj);
            const float g2 = G.affine_map == 0 ? interp_f(I2, pitch, x2 + ii, y2 + jj)
                                               : interp_f(I2, pitch, x2 + (Axx * ii + Axy * jj), y2 + (Ayx * ii + Ayy * jj));
            sum += fabsf(g1 - g2);
        }
        sum = warp_sum(sum);
        if (sum / (float)npx > G.max_residue) status = KLT_LARGE_RESIDUE;
    }
    if (lane == 0) {
        G.A[4 * f + 0] = Axx; G.A[4 * f + 1] = Ayx; G.A[4 * f + 2] = Axy; G.A[4 * f + 3] = Ayy;
        if (status != KLT_TRACKED) {                             // trackFeatures.py:386-395
            vals[f] = status; xs[f] = -1.0; ys[f] = -1.0;
            G.aff_x[f] = -1.0f; G.aff_y[f] = -1.0f; G.has[f] = 0;
        }
    }
}

int klt_launch_affine(klt_ctx *ctx, const klt_params *p, const klt_pyr *p1, const klt_pyr *p2, int n_per_image,
                      const double *x_in, const double *y_in, const int32_t *val_in, double *x, double *y, int32_t *val,
                      klt_affine *st, int *assert_dev) {
    AffineArgs G;
    G.p1 = *p1; G.p2 = *p2;
    G.n_per_image = n_per_image; G.total = n_per_image * p1->batch;
    G.affine_map = p->affine_consistency_check; G.width = p->affine_window_width; G.height = p->affine_window_height;
    G.max_iterations = p->affine_max_iterations;
    G.step_factor = p->step_factor; G.small_det = p->min_determinant; G.th = p->min_displacement;
    G.th_aff = p->affine_min_displacement; G.max_residue = p->affine_max_residue; G.mdd = p->affine_max_displacement_differ;
    G.tw = st->aw + 2; G.th_ = st->ah + 2; G.tn = G.tw * G.th_;
    G.has = st->has; G.aff_x = st->aff_x; G.aff_y = st->aff_y; G.A = st->A; G.tmpl = st->tmpl;
    if (G.total <= 0) return KLT_OK;
    KLT_LAUNCH(ctx, "lk_affine", 0.0, (lk_affine_kernel<<<(G.total + 3) / 4, 128, 0, ctx->stream>>>(G, x_in, y_in, val_in, x, y, val, assert_dev)));
    return KLT_OK;
}

// reset: features with mask != 0 lose their template, aff_x = aff_y = -1, A = identity (selectGoodFeatures.py:120-128)
__global__ void affine_reset_kernel(int n, const int *__restrict__ mask, int *has, float *aff_x, float *aff_y, float *A) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n || (mask && !mask[f])) return;
    has[f] = 0; aff_x[f] = -1.0f; aff_y[f] = -1.0f;
    A[4 * f + 0] = 1.0f; A[4 * f + 1] = 0.0f; A[4 * f + 2] = 0.0f; A[4 * f + 3] = 1.0f;
}
int klt_launch_affine_reset(klt_ctx *ctx, klt_affine *st, const int *mask_dev) {
    if (st->n <= 0) return KLT_OK;
    KLT_LAUNCH(ctx, "affine_reset", 0.0, (affine_reset_kernel<<<(st->n + 127) / 128, 128, 0, ctx->stream>>>(st->n, mask_dev, st->has, st->aff_x, st->aff_y, st->A)));
    return KLT_OK;
}
