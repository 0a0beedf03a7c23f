/* The drop-in boundary from plain C: select features on one synthetic frame, track them into a shifted copy.
 *   gcc -std=c11 -Iinclude examples/c_abi_demo.c -Lpyfeaturetrack_b200 -lkltb200 -Wl,-rpath,$PWD/pyfeaturetrack_b200 -lm -o c_abi_demo
 * Prints "tracked N of M features, median shift (dx, dy)"; exits 3 with the library's message when no B200 is present
 * (there is no CPU fallback).  The kernel taps are what convolve.py:_computeKernels gives for the default context. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "klt_b200.h"

/* gauss / gaussderiv taps of _computeKernels(sigma) (convolve.py:44-102), restated: factor 0.01, normalised */
static void kernels(double sigma, klt_kernel1d *g, klt_kernel1d *d) {
    const int MAXW = KLT_MAX_TAPS, hw = MAXW / 2;
    double gg[KLT_MAX_TAPS], dd[KLT_MAX_TAPS];
    for (int i = -hw; i <= hw; i++) {
        gg[i + hw] = exp(-(double)i * i / (2 * sigma * sigma));
        dd[i + hw] = -i * gg[i + hw];
    }
    const double max_gauss = 1.0, max_deriv = sigma * exp(-0.5), factor = 0.01;
    int gw = MAXW, dw = MAXW;
    for (int i = -hw; fabs(gg[i + hw] / max_gauss) < factor; i++) gw -= 2;
    for (int i = -hw; fabs(dd[i + hw] / max_deriv) < factor; i++) dw -= 2;
    memset(g, 0, sizeof *g); memset(d, 0, sizeof *d);
    g->n = gw; d->n = dw;
    for (int i = 0; i < gw; i++) g->taps[i] = gg[i + (MAXW - gw) / 2];
    for (int i = 0; i < dw; i++) d->taps[i] = dd[i + (MAXW - dw) / 2];
    double den = 0;
    for (int i = 0; i < gw; i++) den += g->taps[i];
    for (int i = 0; i < gw; i++) g->taps[i] /= den;
    const int dhw = dw / 2;
    den = 0;
    for (int i = -dhw; i <= dhw; i++) den -= i * d->taps[i + dhw];
    for (int i = -dhw; i <= dhw; i++) d->taps[i + dhw] /= den;
}

static int cmp(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

int main(void) {
    enum { W = 640, H = 480, N = 200, L = 2, SS = 2 };
    klt_ctx *ctx = NULL;
    if (klt_ctx_create(0, NULL, &ctx) != KLT_OK) {
        fprintf(stderr, "klt_ctx_create: %s\n", klt_last_error(NULL));
        return 3;
    }
    /* a smooth random texture and the same texture shifted by (3, 2) pixels */
    static unsigned char f1[H][W], f2[H][W];
    static float tex[H + 8][W + 8];
    srand(7);
    for (int y = 0; y < H + 8; y++) for (int x = 0; x < W + 8; x++) tex[y][x] = (float)(rand() % 256);
    for (int pass = 0; pass < 3; pass++)
        for (int y = 1; y < H + 7; y++) for (int x = 1; x < W + 7; x++)
            tex[y][x] = 0.2f * (tex[y][x] + tex[y - 1][x] + tex[y + 1][x] + tex[y][x - 1] + tex[y][x + 1]);
    float lo = 1e9f, hi = -1e9f;
    for (int y = 2; y < H + 6; y++) for (int x = 2; x < W + 6; x++) { if (tex[y][x] < lo) lo = tex[y][x]; if (tex[y][x] > hi) hi = tex[y][x]; }
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        f1[y][x] = (unsigned char)(255.f * (tex[y + 4][x + 4] - lo) / (hi - lo));
        f2[y][x] = (unsigned char)(255.f * (tex[y + 2][x + 1] - lo) / (hi - lo));      /* content moves by (+3, +2) */
    }
    klt_params p; memset(&p, 0, sizeof p);
    p.window_width = p.window_height = 7; p.n_levels = L; p.subsampling = SS;
    p.mindist = 10; p.min_eigenvalue = 1; p.max_iterations = 10;
    p.min_determinant = 0.01f; p.min_displacement = 0.1f; p.step_factor = 1.0f;
    p.has_max_residue = 1; p.max_residue = 10.0f;
    p.borderx = p.bordery = 24.0;              /* >= KLTUpdateTCBorder's value for this context; any larger border is legal */
    p.affine_consistency_check = -1; p.affine_window_width = p.affine_window_height = 15;
    klt_taps taps; klt_kernel1d unused;
    kernels(0.7, &taps.smooth, &unused);       /* smooth_sigma_fact * window = 0.1 * 7 */
    kernels(0.9 * SS, &taps.pyramid, &unused); /* pyramid_sigma_fact * subsampling */
    kernels(1.0, &taps.grad_gauss, &taps.grad_deriv);
    klt_pyr *p1 = NULL, *p2 = NULL;
    static double x[N], y[N], x0[N], y0[N];
    static int32_t val[N];
    int rc = klt_pyr_create(ctx, W, H, L, SS, 1, &p1);
    if (!rc) rc = klt_pyr_create(ctx, W, H, L, SS, 1, &p2);
    /* selection works on the gradients of a (strict) one-level build of the first frame */
    klt_pyr *ps = NULL;
    if (!rc) rc = klt_pyr_create(ctx, W, H, 1, SS, 1, &ps);
    if (!rc) rc = klt_pyr_build_u8(ctx, ps, &f1[0][0], W, (size_t)W * H, &taps, KLT_PRECISION_STRICT);
    if (!rc) rc = klt_select_good_features(ctx, &p, ps, 0, NULL, NULL, 0, 0, N, 0, x, y, val, NULL);
    memcpy(x0, x, sizeof x); memcpy(y0, y, sizeof y);
    if (!rc) rc = klt_track_pairs_u8(ctx, &p, &taps, KLT_PRECISION_FAST_WINDOWED, p1, p2, &f1[0][0], &f2[0][0], W, (size_t)W * H, N, x, y, val);
    if (rc) { fprintf(stderr, "error %d: %s\n", rc, klt_last_error(ctx)); return 1; }
    static double dx[N], dy[N];
    int n = 0;
    for (int i = 0; i < N; i++) if (val[i] == 0) { dx[n] = x[i] - x0[i]; dy[n] = y[i] - y0[i]; n++; }
    qsort(dx, n, sizeof(double), cmp); qsort(dy, n, sizeof(double), cmp);
    printf("tracked %d of %d features, median shift (%.2f, %.2f), %lld kernel launches\n", n, N, n ? dx[n / 2] : 0.0, n ? dy[n / 2] : 0.0,
           (long long)klt_launch_count(ctx));
    klt_pyr_destroy(ctx, ps); klt_pyr_destroy(ctx, p1); klt_pyr_destroy(ctx, p2);
    klt_ctx_destroy(ctx);
    return 0;
}
