// Pyramidal Lucas-Kanade on IMAGE-ONLY pyramids (KLT_PRECISION_FAST_WINDOWED): the gradient planes of the reference
// (trackFeatures.py:171-176, `_KLTComputeGradients` per level) are never written to HBM.  A tracked feature only ever
// reads gradients inside its (W+1)^2 window -- ~3 % of a 1080p frame for 1000 features -- so each warp evaluates the
// 7-tap separable gradient pair (convolve.py:245-246) just for the pixels its feature visits:
//
//   * one WARP per feature; per pyramid level the warp stages, with cp.async, a square region of the smoothed image of
//     each frame into shared memory (first image: bilinear footprint of the window + filter radius; second image: the
//     same + MARGIN pixels of slack; SciPy 'reflect' indices at the image border) and meanwhile prefetches the next
//     level's regions into L2;
//   * evaluates the gradient pair on the (W+1)^2 footprint of each window: horizontal pass = one region ROW per lane
//     (sliding 7-value register window, both kernels at once, the rows of both images in one pass where 32 lanes
//     suffice), vertical pass = one COLUMN per lane, the planes gx2, gy2, gx1, gy1 on separate lane groups;
//   * iterates exactly like the dense FAST kernel (lk_track_rows_kernel), except that every bilinear sample comes from
//     shared memory: one global round trip per level instead of one per Newton step;
//   * when the window moves to another integer position the second image's gradients are re-evaluated there (0.2 times
//     per level on the benchmark), and when it walks out of the staged region that is re-staged around the current
//     position (warp-uniform branches; the values do not depend on where the region sits).
//   * The shared-memory layout is bank-aware: the four horizontal-result planes start 8 banks apart, the gradients of an
//     image are one plane of (gx, gy) pairs (64-bit bilinear taps, the two images 16 banks apart), and for 7x7 windows 8 lanes serve a window row and the staged regions use pitch 25, which puts the
//     four window rows of a round on disjoint banks (ncu: conflict replays 40 % -> 16 % of the wavefronts).
//
// Arithmetic: float32 FMA, the same tap order as the dense FAST kernels (c[0..6] left to right / top to bottom), so the
// window gradients agree with the planes `stream_grad_kernel` would have written to ~1 ulp; the tracked positions agree
// with the dense FAST path to ~1e-5 px.  STRICT pyramids never come here.
#include <stdlib.h>

#include "klt_common.cuh"
#include "klt_track_args.cuh"

namespace {

constexpr int RG = 3;      // gradient kernel radius served here (grad_sigma = 1.0: 7 taps); other radii use the planes
constexpr int MARGIN = 2;  // pixels of slack around the start window of the second image
constexpr int WIN_WARPS = 4;     // features (warps) per CTA
constexpr int WIN_MIN_CTAS = 5;  // 4 or 5 CTAs/SM measured equal, 6..8 are 4-6 % slower: more warps only add shared-memory contention

template <int W>
struct Cfg {
    static constexpr int S = W + 1;                         // gradient region of either image: the bilinear footprint of the window
    static constexpr int NG = S + 2 * RG;                   // smoothed-image rows/columns that region needs
    // Window pixels -> lanes.  Linear (k = lane + 32 i) in general; where it costs no extra round, LPR lanes per window row
    // (7x7: 8 lanes per row, 4 rows per round), which lets the pitches below make every bilinear tap a conflict-free access.
    static constexpr int LPR = W <= 4 ? 4 : (W <= 8 ? 8 : 16), RPR = 32 / LPR;
    static constexpr bool ROWMAP = LPR <= 8 && ((W + RPR - 1) / RPR) * 32 <= ((W * W + 31) / 32) * 32;   // 3x3 and 7x7
    // (15x15 qualifies arithmetically, but its linear mapping is already nearly conflict-free and measured 6 % faster)
    static constexpr int PX = ROWMAP ? (W + RPR - 1) / RPR : (W * W + 31) / 32;           // rounds = window pixels per lane
    // staged regions: odd pitches keep the one-row-per-lane walks of the horizontal pass conflict-free; for 7x7, 25 also
    // puts the 4 window rows of a round on disjoint banks (0, 25, 18, 11 + 7 columns)
    static constexpr int N1 = NG, NP1 = W == 7 ? 25 : (N1 | 1);             // first image
    static constexpr int N2 = NG + 2 * MARGIN, NP2 = W == 7 ? 25 : (N2 | 1);   // second image: MARGIN pixels of slack all round
    static constexpr int SP = S | 1;                        // horizontal results (row-per-lane stores, column-per-lane loads)
    static constexpr int GP = (ROWMAP && LPR == S) ? S : SP;   // gradient planes: pitch = lanes per window row where possible
    // per-warp shared memory (floats): staged inputs, horizontal results (deriv, gauss) and gradx / grady of both images
    // The four horizontal-result planes (and the four gradient planes) start 8 banks apart (QUAD_V) or 16 (half-warp
    // mapping): the lane groups of the vertical pass, which walk one plane each, then never share a bank, and neither do
    // the two row groups of the merged horizontal pass when they store.
    static constexpr int GROUP_BANKS = S <= 8 ? 8 : 16;
    static constexpr int pad_to(int n, int banks) { return n + ((banks - n % 32) % 32 + 32) % 32; }
    static constexpr int TSZ = pad_to(NG * SP, GROUP_BANKS);
    static constexpr int GSZ = pad_to(2 * S * GP, 16);      // gradients: ONE plane of (gx, gy) pairs per image -> 64-bit loads
    static constexpr int IN1 = 0, IN2 = IN1 + N1 * NP1;
    static constexpr int TD2 = pad_to(IN2 + N2 * NP2, 0), TG2 = TD2 + TSZ, TD1 = TG2 + TSZ, TG1 = TD1 + TSZ;
    static constexpr int G2 = pad_to(TG1 + TSZ, 0), G1 = G2 + GSZ;
    static constexpr int FLOATS = G1 + GSZ;
    static constexpr bool MERGED_H = 2 * NG <= 32;          // one lane per row of BOTH regions in the horizontal pass
    static constexpr bool QUAD_V = S <= 8;                  // vertical pass: 4 groups of 8 lanes (gx2, gy2, gx1, gy1)
};

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// issue the copies of the N x N input region whose top-left pixel is (sx0, sy0) (SciPy 'reflect' outside the image).
// Interior regions (the common case): half-warp h copies rows h, h+2, ... -- lane column fixed, one pointer bump per
// row, no index arithmetic; columns beyond 16 go through a short flat loop.
template <int N, int NP>
__device__ __forceinline__ void stage_region(float *__restrict__ dst, const float *__restrict__ img, int pitch, int nc, int nr,
                                             int sx0, int sy0, int lane) {
    const bool interior = sx0 >= 0 && sy0 >= 0 && sx0 + N <= nc && sy0 + N <= nr;       // warp-uniform
    if (interior) {
        const float *src = img + (size_t)sy0 * pitch + sx0;
        constexpr int C0 = N < 16 ? N : 16;
        const int half = lane >> 4, col = lane & 15;
        const ptrdiff_t step2 = 2 * (ptrdiff_t)pitch;
        if (col < C0) {
            // running pointers: one 64-bit add per copy instead of a multiply-add chain per (row, pitch) pair
            const float *p = src + (ptrdiff_t)half * pitch + col;
            float *q = dst + half * NP + col;
#pragma unroll
            for (int ry = 0; ry < N; ry += 2) {
                if (ry + 1 < N || half == 0) cp_async4(q + ry * NP, p);
                p += step2;
            }
        }
        if (N > 16) {
            // the remaining REM = N - 16 columns: lane -> (row parity, column) once, then the same row walk
            constexpr int REM = N - 16;
            static_assert(2 * REM <= 32, "region wider than 32 columns");
            const int h2 = lane / REM, c2 = 16 + lane % REM;
            if (lane < 2 * REM) {
                const float *p = src + (ptrdiff_t)h2 * pitch + c2;
                float *q = dst + h2 * NP + c2;
#pragma unroll
                for (int ry = 0; ry < N; ry += 2) {
                    if (ry + 1 < N || h2 == 0) cp_async4(q + ry * NP, p);
                    p += step2;
                }
            }
        }
    } else {
#pragma unroll 1
        for (int idx = lane; idx < N * N; idx += 32) {
            const int ry = idx / N, rx = idx - ry * N;
            cp_async4(dst + ry * NP + rx, img + (size_t)klt_reflect(sy0 + ry, nr) * pitch + klt_reflect(sx0 + rx, nc));
        }
    }
}

// horizontal pass of one region row: S outputs of both kernels from S + 6 consecutive staged values
template <int S>
__device__ __forceinline__ void hrow(const float *__restrict__ row, float *__restrict__ td, float *__restrict__ tg,
                                     const float (&g)[7], const float (&d)[7]) {
    constexpr int N = S + 2 * RG;
    float v[N];
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = row[i];
#pragma unroll
    for (int c = 0; c < S; c++) {
        float hd = d[0] * v[c], hg = g[0] * v[c];
#pragma unroll
        for (int j = 1; j < 7; j++) { hd = fmaf(d[j], v[c + j], hd); hg = fmaf(g[j], v[c + j], hg); }
        td[c] = hd; tg[c] = hg;
    }
}

// vertical pass of one output column: S outputs from S + 6 rows of the horizontal result
template <int S, int SP, int GP>
__device__ __forceinline__ void vcol(const float *__restrict__ src, float *__restrict__ dst, const float (&t)[7]) {
    constexpr int N = S + 2 * RG;
    float v[N];
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = src[i * SP];
#pragma unroll
    for (int r = 0; r < S; r++) {
        float o = t[0] * v[r];
#pragma unroll
        for (int j = 1; j < 7; j++) o = fmaf(t[j], v[r + j], o);
        dst[2 * r * GP] = o;                               // interleaved (gx, gy) plane
    }
}

// Gradients of the window of the second image at offset (ox, oy) inside its staged region and, unless only2, of the
// first image's region.  gx = gauss_v(deriv_h(img)), gy = deriv_v(gauss_h(img))  (convolve.py:245-246).
// tq: this lane's vertical taps for the 4 x 8 mapping (QUAD_V) -- group 0: gx2, 1: gy2, 2: gx1, 3: gy1;
// th1 / th2: for the half-warp mapping (lower half: gauss -> gx, upper half: deriv -> gy).
template <int W>
__device__ __forceinline__ void window_gradients(float *__restrict__ s, const WindowedTaps &K, const float (&tq)[7],
                                                 const float (&th1)[7], const float (&th2)[7], bool same_taps, bool only2,
                                                 int ox, int oy, int lane) {
    using C = Cfg<W>;
    const float *in2 = s + C::IN2 + oy * C::NP2 + ox;
    if (C::MERGED_H && same_taps) {
        const bool second = lane < C::NG;                   // lanes [0, NG): second image, [NG, 2 NG): first image
        const int row = second ? lane : lane - C::NG;
        if (second || (!only2 && row < C::NG)) {
            const float *src = second ? in2 + row * C::NP2 : s + C::IN1 + row * C::NP1;
            float *td = s + (second ? C::TD2 : C::TD1) + row * C::SP, *tg = s + (second ? C::TG2 : C::TG1) + row * C::SP;
            hrow<C::S>(src, td, tg, K.g2, K.d2);
        }
    } else {
        if (lane < C::NG) hrow<C::S>(in2 + lane * C::NP2, s + C::TD2 + lane * C::SP, s + C::TG2 + lane * C::SP, K.g2, K.d2);
        if (!only2 && lane < C::NG) hrow<C::S>(s + C::IN1 + lane * C::NP1, s + C::TD1 + lane * C::SP, s + C::TG1 + lane * C::SP, K.g1, K.d1);
    }
    __syncwarp();
    if (C::QUAD_V) {
        const int grp = lane >> 3, col = lane & 7;
        if (col < C::S && (!only2 || grp < 2)) {
            const int src = grp == 0 ? C::TD2 : (grp == 1 ? C::TG2 : (grp == 2 ? C::TD1 : C::TG1));
            const int dst = (grp < 2 ? C::G2 : C::G1) + (grp & 1);            // gx at even, gy at odd floats
            vcol<C::S, C::SP, C::GP>(s + src + col, s + dst + 2 * col, tq);
        }
    } else {
        const bool upper = lane >= 16;
        const int col = lane & 15;
        if (col < C::S) {
            vcol<C::S, C::SP, C::GP>(s + (upper ? C::TG2 : C::TD2) + col, s + C::G2 + (upper ? 1 : 0) + 2 * col, th2);
            if (!only2) vcol<C::S, C::SP, C::GP>(s + (upper ? C::TG1 : C::TD1) + col, s + C::G1 + (upper ? 1 : 0) + 2 * col, th1);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ float bil(const float *p, int P, float w00, float w01, float w10, float w11) {
    return fmaf(w11, p[P + 1], fmaf(w10, p[P], fmaf(w01, p[1], w00 * p[0])));
}
// bilinear sample of an interleaved (gx, gy) plane: four 64-bit loads for both gradients; P = pitch in pairs
__device__ __forceinline__ float2 bil2(const float *p, int P, float w00, float w01, float w10, float w11) {
    const float2 a = *reinterpret_cast<const float2 *>(p), b = *reinterpret_cast<const float2 *>(p + 2);
    const float2 c = *reinterpret_cast<const float2 *>(p + 2 * P), d = *reinterpret_cast<const float2 *>(p + 2 * P + 2);
    return make_float2(fmaf(w11, d.x, fmaf(w10, c.x, fmaf(w01, b.x, w00 * a.x))),
                       fmaf(w11, d.y, fmaf(w10, c.y, fmaf(w01, b.y, w00 * a.y))));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Five warp-wide sums at once, 16 shuffles instead of 25: after the first exchange the lower half-warp owns a, b, c and
// the upper one d, e; after the second each quarter owns one of a, b, d, e (c rides along on the lower half); three
// plain butterfly rounds finish them and five broadcasts hand the totals to every lane.
__device__ __forceinline__ void warp_sum5(float &a, float &b, float &c, float &d, float &e, int lane) {
    const unsigned int full = 0xffffffffu;
    const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
    const float t0 = __shfl_xor_sync(full, up16 ? a : d, 16);
    const float t1 = __shfl_xor_sync(full, up16 ? b : e, 16);
    const float t2 = __shfl_xor_sync(full, up16 ? c : 0.f, 16);
    const float A = (up16 ? d : a) + t0, B = (up16 ? e : b) + t1;
    float Cc = up16 ? 0.f : c + t2;
    float X = (up8 ? B : A) + __shfl_xor_sync(full, up8 ? A : B, 8);
    Cc += __shfl_xor_sync(full, Cc, 8);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { X += __shfl_xor_sync(full, X, o); Cc += __shfl_xor_sync(full, Cc, o); }
    a = __shfl_sync(full, X, 0); b = __shfl_sync(full, X, 8); d = __shfl_sync(full, X, 16); e = __shfl_sync(full, X, 24);
    c = __shfl_sync(full, Cc, 0);
}

// L2 prefetch of one row of an N-wide region (top-left input pixel (sx0, sy0))
template <int N>
__device__ __forceinline__ void prefetch_region(const float *__restrict__ img, int pitch, int nc, int nr, int sx0, int sy0, int row) {
    const int y = min(max(sy0 + row, 0), nr - 1);
    const int xa = min(max(sx0, 0), nc - 1), xb = min(max(sx0 + N - 1, 0), nc - 1);
    const float *p = img + (size_t)y * pitch;
#pragma unroll
    for (int k = 0; k * 16 < N; k++) prefetch_l2(p + min(xa + 16 * k, xb));      // one touch per 64 bytes
    prefetch_l2(p + xb);
}

template <int W, int MINB>
__global__ void __launch_bounds__(WIN_WARPS * 32, MINB)
lk_windowed_kernel(const __grid_constant__ TrackArgs A, const __grid_constant__ WindowedTaps K, double *__restrict__ xs,
                   double *__restrict__ ys, int *__restrict__ vals, unsigned long long *__restrict__ iters_total,
                   int *__restrict__ assert_flag) {
    using C = Cfg<W>;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = A.f_begin + blockIdx.x * WIN_WARPS + warp;
    if (f >= A.total) return;                     // one feature per warp: all control flow below is warp-uniform
    if (vals[f] < 0) return;                      // trackFeatures.py:253
    float *s = smem + warp * C::FLOATS;
    const int image = f / A.n_per_image;
    const float *const plane1 = A.p1.base + (size_t)image * A.p1.plane_floats;      // intensity planes of this feature's pair
    const float *const plane2 = A.p2.base + (size_t)image * A.p2.plane_floats;
    constexpr int hw = W / 2;
    const double ss = (double)A.ss, inv_ss = 1.0 / ss;        // subsampling is a power of two: x * inv_ss == x / ss exactly
    double xloc = xs[f], yloc = ys[f];
    for (int r = A.n_levels - 1; r >= 0; r--) { xloc *= inv_ss; yloc *= inv_ss; }
    double xout = xloc, yout = yloc;
    int st = KLT_TRACKED;
    unsigned int my_iters = 0;
    bool alive = true;
    // this lane's vertical taps for the two lane mappings of the vertical pass, and whether both images share their kernels
    float tq[7], th1[7], th2[7];
    const bool same_taps = K.same != 0;
#pragma unroll
    for (int j = 0; j < 7; j++) {
        const int grp = lane >> 3;
        tq[j] = grp == 0 ? K.g2[j] : (grp == 1 ? K.d2[j] : (grp == 2 ? K.g1[j] : K.d1[j]));
        th1[j] = lane >= 16 ? K.d1[j] : K.g1[j];
        th2[j] = lane >= 16 ? K.d2[j] : K.g2[j];
    }

    for (int r = A.n_levels - 1; r >= 0; r--) {
        xloc *= ss; yloc *= ss; xout *= ss; yout *= ss;
        if (!alive) continue;
        const int nc = A.p1.lv[r].w, nr = A.p1.lv[r].h, pitch = A.p1.lv[r].pitch;
        const float *I1 = plane1 + A.p1.lv[r].off, *I2 = plane2 + A.p2.lv[r].off;
        const float x1 = (float)xloc, y1 = (float)yloc;
        const int ix1 = (int)x1, iy1 = (int)y1;
        if (!(ix1 - hw >= 0 && iy1 - hw >= 0 && ix1 + hw + 2 <= nc && iy1 + hw + 2 <= nr)) {   // pyx:35
            if (lane == 0) atomicExch(assert_flag, 1);
            st = KLT_INTERNAL_ASSERT;
            alive = false;
            continue;
        }
        float x2 = (float)xout, y2 = (float)yout;
        int status = KLT_TRACKED, iteration = 0;
        const float fnc = (float)nc, fnr = (float)nr, fhw = (float)hw;
        // (rx0, ry0): window position at offset (0, 0) of the staged region of the second image; the start window sits
        // MARGIN pixels inside.  (gix, giy): window position the gradients in shared memory belong to.
        int rx0 = (int)x2 - hw - MARGIN, ry0 = (int)y2 - hw - MARGIN;
        int gix = -0x40000000, giy = -0x40000000;
        const bool start_inside = !(x2 - fhw < 0.f || fnc - (x2 + fhw) < 1.001f || y2 - fhw < 0.f || fnr - (y2 + fhw) < 1.001f);
        // one memory round trip: both regions of this level
        stage_region<C::N1, C::NP1>(s + C::IN1, I1, pitch, nc, nr, ix1 - hw - RG, iy1 - hw - RG, lane);
        if (start_inside) stage_region<C::N2, C::NP2>(s + C::IN2, I2, pitch, nc, nr, rx0 - RG, ry0 - RG, lane);
        if (r > 0) {
            // pull the next (finer) level's regions towards L2 while this level computes
            const int ncn = A.p1.lv[r - 1].w, nrn = A.p1.lv[r - 1].h, pn = A.p1.lv[r - 1].pitch;
            const float xn1 = (float)(xloc * ss), yn1 = (float)(yloc * ss);
            const float xn2 = (float)((double)x2 * ss), yn2 = (float)((double)y2 * ss);
            if (lane < C::N2)
                prefetch_region<C::N2>(plane2 + A.p2.lv[r - 1].off, pn, ncn, nrn, (int)xn2 - hw - MARGIN - RG, (int)yn2 - hw - MARGIN - RG, lane);
            if (lane < C::N1)
                prefetch_region<C::N1>(plane1 + A.p1.lv[r - 1].off, pn, ncn, nrn, (int)xn1 - hw - RG, (int)yn1 - hw - RG, lane);
        }
        cp_async_wait_all();
        __syncwarp();
        if (start_inside) {
            window_gradients<W>(s, K, tq, th1, th2, same_taps, false, MARGIN, MARGIN, lane);
            gix = (int)x2; giy = (int)y2;
        } else {
            // the loop below leaves with OOB at once; only the template is needed.  (The second image's buffer holds
            // stale but finite data; its results are never used.)
            window_gradients<W>(s, K, tq, th1, th2, same_taps, false, MARGIN, MARGIN, lane);
        }
        // template: the first image's window, gradients interpolated like the image (trackFeatures.py:87-92)
        float T[C::PX], Tgx[C::PX], Tgy[C::PX];
        int po[C::PX], pi[C::PX];                  // this lane's window pixels: offsets in a gradient array / in the staged region
        bool pon[C::PX];
        {
            const float ax = x1 - (float)ix1, ay = y1 - (float)iy1;
            const float w11 = ax * ay, w01 = ax - w11, w10 = ay - w11, w00 = 1.f - ax - ay + w11;
#pragma unroll
            for (int i = 0; i < C::PX; i++) {
                int pr, pc;
                bool on;
                if (C::ROWMAP) { pr = C::RPR * i + lane / C::LPR; pc = lane % C::LPR; on = pr < W && pc < W; }
                else { const int k = lane + 32 * i; on = k < W * W; pr = k / W; pc = k - pr * W; }   // (true at compile time for most i)
                if (!on) { pr = 0; pc = 0; }
                pon[i] = on;
                po[i] = pr * C::GP + pc;
                pi[i] = (pr + RG) * C::NP2 + pc + RG;
                T[i] = on ? bil(s + C::IN1 + (pr + RG) * C::NP1 + pc + RG, C::NP1, w00, w01, w10, w11) : 0.f;
                const float2 tg = on ? bil2(s + C::G1 + 2 * po[i], C::GP, w00, w01, w10, w11) : make_float2(0.f, 0.f);
                Tgx[i] = tg.x; Tgy[i] = tg.y;
            }
        }
        int ox = MARGIN, oy = MARGIN;
        for (;;) {
            if (x2 - fhw < 0.f || fnc - (x2 + fhw) < 1.001f || y2 - fhw < 0.f || fnr - (y2 + fhw) < 1.001f) {
                status = KLT_OOB;
                break;
            }
            const int ix = (int)x2, iy = (int)y2;
            if (ix != gix || iy != giy) {                   // the window moved to another pixel: new gradients
                ox = ix - hw - rx0; oy = iy - hw - ry0;
                __syncwarp();
                if (ox < 0 || oy < 0 || ox > 2 * MARGIN || oy > 2 * MARGIN) {      // it even left the staged region
                    rx0 = ix - hw - MARGIN; ry0 = iy - hw - MARGIN;
                    stage_region<C::N2, C::NP2>(s + C::IN2, I2, pitch, nc, nr, rx0 - RG, ry0 - RG, lane);
                    cp_async_wait_all();
                    __syncwarp();
                    ox = MARGIN; oy = MARGIN;
                }
                window_gradients<W>(s, K, tq, th1, th2, same_taps, true, ox, oy, lane);
                gix = ix; giy = iy;
            }
            const float ax = x2 - (float)ix, ay = y2 - (float)iy;
            const float w11 = ax * ay, w01 = ax - w11, w10 = ay - w11, w00 = 1.f - ax - ay + w11;
            const float *bi = s + C::IN2 + oy * C::NP2 + ox;
            float gxx = 0.f, gxy = 0.f, gyy = 0.f, ex = 0.f, ey = 0.f;
#pragma unroll
            for (int i = 0; i < C::PX; i++) {
                if (pon[i]) {
                    const float P = bil(bi + pi[i], C::NP2, w00, w01, w10, w11);
                    const float2 Pg = bil2(s + C::G2 + 2 * po[i], C::GP, w00, w01, w10, w11);
                    const float diff = T[i] - P, gx = Tgx[i] + Pg.x, gy = Tgy[i] + Pg.y;
                    gxx = fmaf(gx, gx, gxx); gxy = fmaf(gx, gy, gxy); gyy = fmaf(gy, gy, gyy);
                    ex = fmaf(diff, gx, ex); ey = fmaf(diff, gy, ey);
                }
            }
            warp_sum5(gxx, gxy, gyy, ex, ey, lane);
            ex *= A.step_factor; ey *= A.step_factor;
            const float det = __fsub_rn(__fmul_rn(gxx, gyy), __fmul_rn(gxy, gxy));
            if (det < A.small_det) { status = KLT_SMALL_DET; break; }
            const float inv = __frcp_rn(det);             // one reciprocal instead of two divisions (within 1 ulp of them)
            const float dx = __fsub_rn(__fmul_rn(gyy, ex), __fmul_rn(gxy, ey)) * inv;
            const float dy = __fsub_rn(__fmul_rn(gxx, ey), __fmul_rn(gxy, ex)) * inv;
            x2 += dx; y2 += dy;
            iteration++;
            if (!((fabsf(dx) >= A.th || fabsf(dy) >= A.th) && iteration < A.max_iterations)) break;
        }
        my_iters += iteration;
        {
            const double x2d = (double)x2, y2d = (double)y2, hwd = W / 2.0;
            if (x2d - hwd < 0.0 || (double)nc - (x2d + hwd) < 1.001 || y2d - hwd < 0.0 || (double)nr - (y2d + hwd) < 1.001)
                status = KLT_OOB;
        }
        if (status == KLT_TRACKED && A.has_max_residue) {
            const int ix = (int)x2, iy = (int)y2;
            ox = ix - hw - rx0; oy = iy - hw - ry0;
            if (ox < 0 || oy < 0 || ox > 2 * MARGIN || oy > 2 * MARGIN) {
                rx0 = ix - hw - MARGIN; ry0 = iy - hw - MARGIN;
                __syncwarp();
                stage_region<C::N2, C::NP2>(s + C::IN2, I2, pitch, nc, nr, rx0 - RG, ry0 - RG, lane);
                cp_async_wait_all();
                __syncwarp();
                ox = MARGIN; oy = MARGIN;
            }
            const float ax = x2 - (float)ix, ay = y2 - (float)iy;
            const float w11 = ax * ay, w01 = ax - w11, w10 = ay - w11, w00 = 1.f - ax - ay + w11;
            const float *bi = s + C::IN2 + oy * C::NP2 + ox;
            float res = 0.f;
#pragma unroll
            for (int i = 0; i < C::PX; i++)
                if (pon[i]) res += fabsf(T[i] - bil(bi + pi[i], C::NP2, w00, w01, w10, w11));
            res = warp_sum(res) / (float)(W * W);
            if (res > A.max_residue) status = KLT_LARGE_RESIDUE;
        }
        __syncwarp();                              // all lanes are done with the regions before the next level overwrites them
        xout = (double)x2; yout = (double)y2;
        if (A.retain) st = KLT_TRACKED;
        else if (status == KLT_SMALL_DET || status == KLT_OOB || status == KLT_LARGE_RESIDUE) st = status;
        else if (iteration >= A.max_iterations) st = KLT_MAX_ITERATIONS;
        else st = KLT_TRACKED;
        if (st == KLT_SMALL_DET || st == KLT_OOB) alive = false;                               // :284-285
    }
    if (lane == 0) {
        if (my_iters) atomicAdd(iters_total, (unsigned long long)my_iters);
        if (st == KLT_INTERNAL_ASSERT) return;
        const int W0 = A.p1.lv[0].w, H0 = A.p1.lv[0].h;
        const bool oob = xout < A.borderx || xout > (double)(W0 - 1) - A.borderx || yout < A.bordery ||
                         yout > (double)(H0 - 1) - A.bordery;
        if (st == KLT_OOB || oob) { xs[f] = -1.0; ys[f] = -1.0; vals[f] = KLT_OOB; }
        else if (st == KLT_SMALL_DET || st == KLT_LARGE_RESIDUE || st == KLT_MAX_ITERATIONS) { xs[f] = -1.0; ys[f] = -1.0; vals[f] = st; }
        else { xs[f] = xout; ys[f] = yout; vals[f] = KLT_TRACKED; }
    }
}

template <int W>
int launch_windowed(klt_ctx *ctx, const TrackArgs &A, const WindowedTaps &K, double *x, double *y, int32_t *v,
                    unsigned long long *it, int *af) {
    const int nfeat = A.total - A.f_begin;
    const int blocks = (nfeat + WIN_WARPS - 1) / WIN_WARPS;
    // algorithmic bytes: the staged regions of both images on every level (restaging not counted) + the feature records
    const double bytes = (double)nfeat * (A.n_levels * 4.0 * (Cfg<W>::N1 * Cfg<W>::N1 + Cfg<W>::N2 * Cfg<W>::N2) + 40.0);
    const size_t smem = (size_t)WIN_WARPS * Cfg<W>::FLOATS * sizeof(float);
    constexpr int MINB = W <= 11 ? WIN_MIN_CTAS : 3;
    if (smem > 48 * 1024) KLT_CUDA(ctx, cudaFuncSetAttribute(lk_windowed_kernel<W, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KLT_LAUNCH(ctx, "lk_windowed", bytes, (lk_windowed_kernel<W, MINB><<<blocks, WIN_WARPS * 32, smem, ctx->stream>>>(A, K, x, y, v, it, af)));
    return KLT_OK;
}

bool radius3(const klt_kernel1d &k) { return k.n == 2 * RG + 1; }
void flip7(const klt_kernel1d &k, float *dst) { for (int j = 0; j < 7; j++) dst[j] = (float)k.taps[6 - j]; }

}  // namespace

bool klt_windowed_supported(const klt_params *p, const klt_pyr *p1, const klt_pyr *p2) {
    if (p1->precision == KLT_PRECISION_STRICT || p2->precision == KLT_PRECISION_STRICT) return false;
    if (p->window_width != p->window_height || (p->window_width & 1) == 0 || p->window_width < 3 || p->window_width > 15) return false;
    if (!p1->hx || !p2->hx || !p1->hx->taps_valid || !p2->hx->taps_valid) return false;
    return radius3(p1->hx->taps.grad_gauss) && radius3(p1->hx->taps.grad_deriv) && radius3(p2->hx->taps.grad_gauss) &&
           radius3(p2->hx->taps.grad_deriv);
}

int klt_launch_track_windowed(klt_ctx *ctx, const TrackArgs &A, const klt_pyr *p1, const klt_pyr *p2, double *x_dev,
                              double *y_dev, int32_t *val_dev, unsigned long long *iters_dev, int *assert_dev) {
    WindowedTaps K;
    flip7(p1->hx->taps.grad_gauss, K.g1); flip7(p1->hx->taps.grad_deriv, K.d1);
    flip7(p2->hx->taps.grad_gauss, K.g2); flip7(p2->hx->taps.grad_deriv, K.d2);
    K.same = 1;
    for (int j = 0; j < 7; j++) K.same = K.same && K.g1[j] == K.g2[j] && K.d1[j] == K.d2[j];
    // second generation (two features per warp, TMA staging) unless $KLT_B200_WINDOWED_GEN=1; windows wider than
    // $KLT_B200_WINDOWED2_MAXW stay on this file's kernel
    static const int gen = getenv("KLT_B200_WINDOWED_GEN") ? atoi(getenv("KLT_B200_WINDOWED_GEN")) : 2;
    static const int maxw2 = getenv("KLT_B200_WINDOWED2_MAXW") ? atoi(getenv("KLT_B200_WINDOWED2_MAXW")) : 15;
    if (gen != 1 && A.w <= maxw2) {
        const int rc = klt_launch_track_windowed2(ctx, A, K, p1, p2, x_dev, y_dev, val_dev, iters_dev, assert_dev);
        if (rc <= 0) return rc;             // 1: no tensor maps on this driver -> first generation
    }
    switch (A.w) {
        case 3: return launch_windowed<3>(ctx, A, K, x_dev, y_dev, val_dev, iters_dev, assert_dev);
        case 5: return launch_windowed<5>(ctx, A, K, x_dev, y_dev, val_dev, iters_dev, assert_dev);
        case 7: return launch_windowed<7>(ctx, A, K, x_dev, y_dev, val_dev, iters_dev, assert_dev);
        case 9: return launch_windowed<9>(ctx, A, K, x_dev, y_dev, val_dev, iters_dev, assert_dev);
        case 11: return launch_windowed<11>(ctx, A, K, x_dev, y_dev, val_dev, iters_dev, assert_dev);
        case 13: return launch_windowed<13>(ctx, A, K, x_dev, y_dev, val_dev, iters_dev, assert_dev);
        case 15: return launch_windowed<15>(ctx, A, K, x_dev, y_dev, val_dev, iters_dev, assert_dev);
    }
    return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "windowed tracking: window %d not covered", A.w);
}
