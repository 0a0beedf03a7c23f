// Warp-streaming pyramid kernels for the FAST (fp32) path (sm_100a).
//
// Why this shape: at level 0 the pipeline is u8 -> 5-tap smooth (h, v) -> 7-tap gradient pair (h, v for gx and gy):
// 38 FMAs per pixel for 13 bytes of compulsory traffic.  At 6.5 TB/s each SM must retire ~1.8 pixels per clock, i.e.
// ~70 FMA/clk of its 128 -- the kernel is as much issue-bound as HBM-bound, so the design goal is to spend issue slots
// on FMAs only:
//   * one WARP owns a strip of columns and marches down the rows of its segment; every lane owns 4 adjacent columns;
//   * vertical filters never touch memory: each lane keeps the (2r+1) partially accumulated output rows of its 4 columns
//     in registers; a new input row updates them with one FMA each, and the FMA's destination register performs the
//     shift (acc[m] = fma(c, v, acc[m+1])), so there are no moves and no unrolling by the kernel length;
//   * horizontal filters get the 3-5 neighbour columns from the adjacent lanes with warp shuffles (no shared memory,
//     no block barriers); lanes 0 and 31 are halo lanes whose own outputs are discarded (30/32 = 94 % efficiency);
//   * no vertical halo recomputation inside a segment (a 2-D tile design recomputes ~30 %); segments overlap by the
//     filter radii only (warm-up rows);
//   * loads and stores are 128-bit (32-bit for the u8 frame: 4 pixels), coalesced along the warp's strip.
//
// Borders: the reference filters with SciPy's mode='reflect' (half-sample symmetric).  A SYMMETRIC filter commutes with
// that extension, so the fused level-0 kernel simply reads the u8 frame through reflected row/column indices and treats
// the smoothed image on the extended domain as the reflect-extension of the smoothed image (exact in real arithmetic,
// ~1e-7 relative in fp32).  Decimation does not commute with it, so levels >= 1 run as two streaming kernels (decimate,
// then gradients) and the gradient kernel reflects its input indices directly.
//
// These kernels serve KLT_PRECISION_FAST only.  STRICT (bit-exact) mode and unusual kernel radii use klt_conv.cu.
#include "klt_common.cuh"

#define WARPS_PER_CTA 4
#define FULLMASK 0xffffffffu

struct StreamTaps {
    // c[j] multiplies in[x + j - r] (convolution order), zero padded
    float s[9];      // smoothing (level 0) or pyramid gauss; up to 9 / 11 taps
    float p[11];
    float g[7];      // gradient gauss / deriv, radius 3
    float d[7];
};

__device__ __forceinline__ float u8_to_f32(unsigned int word, int byte) {
    // 0x4B000000 | b is the float 8388608 + b; exact for b in 0..255, and runs on the ALU/FMA pipes (no I2F)
    const unsigned int sel = 0x7650u | (unsigned int)byte;   // result byte0 = word.byte, bytes1..3 = 0x4B0000
    return __uint_as_float(__byte_perm(word, 0x4B000000u, sel)) - 8388608.0f;
}

__device__ __forceinline__ unsigned int load_u8_quad(const unsigned char *__restrict__ row, int c, int W, bool fast) {
    if (fast) return __ldg(reinterpret_cast<const unsigned int *>(row + c));
    unsigned int w = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) w |= (unsigned int)row[klt_reflect(c + i, W)] << (8 * i);
    return w;
}

__device__ __forceinline__ float4 load_f32_quad(const float *__restrict__ row, int c, int W, bool fast) {
    if (fast) return __ldg(reinterpret_cast<const float4 *>(row + c));
    float4 v;
    v.x = row[klt_reflect(c, W)]; v.y = row[klt_reflect(c + 1, W)];
    v.z = row[klt_reflect(c + 2, W)]; v.w = row[klt_reflect(c + 3, W)];
    return v;
}

// vertical accumulate-and-shift: acc[m] <- acc[m+1] + c[2R-m]*v ; returns the completed output (old acc[0] + c[2R]*v)
template <int R>
__device__ __forceinline__ float vacc(float (&acc)[2 * R], const float *c, float v) {
    const float out = fmaf(c[2 * R], v, acc[0]);
#pragma unroll
    for (int m = 0; m < 2 * R - 1; m++) acc[m] = fmaf(c[2 * R - 1 - m], v, acc[m + 1]);
    acc[2 * R - 1] = c[0] * v;
    return out;
}

// ---- gradient stage shared by the level-0 kernel and the gradient-only kernel -----------------------------------
// s[0..3] = this lane's 4 columns of the (smoothed) image row; neighbours come from lanes +-1.
struct GradState {
    float ax[4][6], ay[4][6];   // pending gx / gy rows (radius 3)
};

__device__ __forceinline__ void grad_row(GradState &st, const StreamTaps &T, const float (&s)[4], float (&gx)[4], float (&gy)[4]) {
    float e[10];                 // columns c-3 .. c+6
    e[0] = __shfl_up_sync(FULLMASK, s[1], 1);
    e[1] = __shfl_up_sync(FULLMASK, s[2], 1);
    e[2] = __shfl_up_sync(FULLMASK, s[3], 1);
    e[3] = s[0]; e[4] = s[1]; e[5] = s[2]; e[6] = s[3];
    e[7] = __shfl_down_sync(FULLMASK, s[0], 1);
    e[8] = __shfl_down_sync(FULLMASK, s[1], 1);
    e[9] = __shfl_down_sync(FULLMASK, s[2], 1);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float hd = T.d[0] * e[i], hg = T.g[0] * e[i];
#pragma unroll
        for (int j = 1; j < 7; j++) { hd = fmaf(T.d[j], e[i + j], hd); hg = fmaf(T.g[j], e[i + j], hg); }
        gx[i] = vacc<3>(st.ax[i], T.g, hd);      // gx = gauss_v( deriv_h )
        gy[i] = vacc<3>(st.ay[i], T.d, hg);      // gy = deriv_v( gauss_h )
    }
}

__device__ __forceinline__ void store_quad(float *__restrict__ row, int c, int W, const float (&v)[4]) {
    if (c + 3 < W) {
        *reinterpret_cast<float4 *>(row + c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (c + i < W) row[c + i] = v[i];
    }
}

// ---- level 0: u8 frame -> smoothed image, gradx, grady -----------------------------------------------------------
// Lane l of a warp owns columns c = x0 + 4*(l-1) .. c+3 where x0 = 120*strip; lanes 1..30 store outputs.
template <int RS>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
stream_level0_kernel(const unsigned char *__restrict__ frames, size_t pitch, size_t frame_stride, float *__restrict__ img,
                     float *__restrict__ gxo, float *__restrict__ gyo, int out_pitch, size_t out_stride, int W, int H,
                     int rows_per_seg, int n_strips, const __grid_constant__ StreamTaps T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * WARPS_PER_CTA + warp;
    if (strip >= n_strips) return;
    const int ys = blockIdx.y * rows_per_seg, ye = min(H, ys + rows_per_seg);
    const int c = strip * 120 + 4 * (lane - 1);
    const unsigned char *src = frames + (size_t)blockIdx.z * frame_stride;
    const size_t ob = (size_t)blockIdx.z * out_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | pitch) & 3) == 0;
    const bool fast = aligned && c >= 0 && c + 3 < W;
    // halo lanes also need the quad beyond them for the horizontal smooth
    const bool edge = lane == 0 || lane == 31;
    const int ce = lane == 0 ? c - 4 : c + 4;
    const bool fast_e = aligned && ce >= 0 && ce + 3 < W;
    const bool writer = lane >= 1 && lane <= 30 && c < W;

    float sa[4][2 * RS];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int m = 0; m < 2 * RS; m++) sa[i][m] = 0.f;
    GradState gs;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int m = 0; m < 6; m++) { gs.ax[i][m] = 0.f; gs.ay[i][m] = 0.f; }

    const int t0 = ys - 3 - RS, t1 = ye + 3 + RS;       // input rows [t0, t1)
    unsigned int wq = load_u8_quad(src + (size_t)klt_reflect(t0, H) * pitch, c, W, fast);
    unsigned int we = edge ? load_u8_quad(src + (size_t)klt_reflect(t0, H) * pitch, ce, W, fast_e) : 0u;
    for (int t = t0; t < t1; t++) {
        const unsigned int w_cur = wq, e_cur = we;
        if (t + 1 < t1) {                                // software prefetch of the next row
            const unsigned char *nrow = src + (size_t)klt_reflect(t + 1, H) * pitch;
            wq = load_u8_quad(nrow, c, W, fast);
            if (edge) we = load_u8_quad(nrow, ce, W, fast_e);
        }
        // ---- horizontal smooth ----
        float u[4 + 2 * RS];                             // columns c-RS .. c+3+RS
        float f[4];
#pragma unroll
        for (int i = 0; i < 4; i++) f[i] = u8_to_f32(w_cur, i);
#pragma unroll
        for (int i = 0; i < 4; i++) u[RS + i] = f[i];
#pragma unroll
        for (int k = 0; k < RS; k++) {                   // left neighbours: lane-1's f[4-RS+k]; right: lane+1's f[k]
            u[k] = __shfl_up_sync(FULLMASK, f[4 - RS + k], 1);
            u[RS + 4 + k] = __shfl_down_sync(FULLMASK, f[k], 1);
        }
        if (edge) {
#pragma unroll
            for (int k = 0; k < RS; k++) {
                if (lane == 0) u[k] = u8_to_f32(e_cur, 4 - RS + k);
                else u[RS + 4 + k] = u8_to_f32(e_cur, k);
            }
        }
        float s[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float h = T.s[0] * u[i];
#pragma unroll
            for (int j = 1; j < 2 * RS + 1; j++) h = fmaf(T.s[j], u[i + j], h);
            s[i] = vacc<RS>(sa[i], T.s, h);              // vertical smooth: completes row t - RS
        }
        const int r = t - RS;
        if (r < ys - 3) continue;                        // vertical smooth still warming up
        if (writer && r >= ys && r < ye) store_quad(img + ob + (size_t)r * out_pitch, c, W, s);
        float gx[4], gy[4];
        grad_row(gs, T, s, gx, gy);
        const int q = r - 3;
        if (writer && q >= ys) {                         // q < ye by construction
            store_quad(gxo + ob + (size_t)q * out_pitch, c, W, gx);
            store_quad(gyo + ob + (size_t)q * out_pitch, c, W, gy);
        }
    }
}

// ---- gradients only: float image -> gradx, grady (levels >= 1) ---------------------------------------------------
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
stream_grad_kernel(const float *__restrict__ in, int in_pitch, size_t in_stride, float *__restrict__ gxo,
                   float *__restrict__ gyo, int out_pitch, size_t out_stride, int W, int H, int rows_per_seg,
                   int n_strips, const __grid_constant__ StreamTaps T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * WARPS_PER_CTA + warp;
    if (strip >= n_strips) return;
    const int ys = blockIdx.y * rows_per_seg, ye = min(H, ys + rows_per_seg);
    const int c = strip * 120 + 4 * (lane - 1);
    const float *src = in + (size_t)blockIdx.z * in_stride;
    const size_t ob = (size_t)blockIdx.z * out_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (in_pitch & 3) == 0;
    const bool fast = aligned && c >= 0 && c + 3 < W;
    const bool writer = lane >= 1 && lane <= 30 && c < W;
    GradState gs;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int m = 0; m < 6; m++) { gs.ax[i][m] = 0.f; gs.ay[i][m] = 0.f; }
    const int t0 = ys - 3, t1 = ye + 3;
    float4 nxt = load_f32_quad(src + (size_t)klt_reflect(t0, H) * in_pitch, c, W, fast);
    for (int t = t0; t < t1; t++) {
        const float4 cur = nxt;
        if (t + 1 < t1) nxt = load_f32_quad(src + (size_t)klt_reflect(t + 1, H) * in_pitch, c, W, fast);
        const float s[4] = {cur.x, cur.y, cur.z, cur.w};
        float gx[4], gy[4];
        grad_row(gs, T, s, gx, gy);
        const int q = t - 3;
        if (writer && q >= ys) {
            store_quad(gxo + ob + (size_t)q * out_pitch, c, W, gx);
            store_quad(gyo + ob + (size_t)q * out_pitch, c, W, gy);
        }
    }
}

// ---- pyramid step for subsampling 2: out[Y][X] = smooth11(in)[2Y+1][2X+1] -----------------------------------------
// Lane owns output columns X..X+3 (X = 128*strip + 4*lane) = input columns 2X..2X+7; it needs 2X-4 .. 2X+12.
// Vertical: input rows arrive in (even, odd) pairs; output Y completes with even row 2Y+6.  Five partial outputs are
// pending per column; 11 FMAs per column per output row, no moves.
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
stream_down2_kernel(const float *__restrict__ in, int in_pitch, size_t in_stride, int W, int H, float *__restrict__ out,
                    int out_pitch, size_t out_stride, int OW, int OH, int rows_per_seg, int n_strips,
                    const __grid_constant__ StreamTaps T) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * WARPS_PER_CTA + warp;
    if (strip >= n_strips) return;
    const int ys = blockIdx.y * rows_per_seg, ye = min(OH, ys + rows_per_seg);
    const int X = strip * 128 + 4 * lane;
    const int ci = 2 * X;
    const float *src = in + (size_t)blockIdx.z * in_stride;
    float *dst = out + (size_t)blockIdx.z * out_stride;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (in_pitch & 3) == 0;
    const bool fast0 = aligned && ci >= 0 && ci + 3 < W, fast1 = aligned && ci + 7 < W;
    const bool fastl = aligned && ci - 4 >= 0 && ci - 1 < W;
    const bool fastr0 = aligned && ci + 11 < W, fastr1 = aligned && ci + 15 < W;

    float P[4][5];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int m = 0; m < 5; m++) P[i][m] = 0.f;

    // horizontal 11-tap at the 4 sampled columns of one input row (reflect-mapped row index)
    auto hrow = [&](int r, float (&h)[4]) {
        const float *row = src + (size_t)klt_reflect(r, H) * in_pitch;
        const float4 a = load_f32_quad(row, ci, W, fast0), b = load_f32_quad(row, ci + 4, W, fast1);
        float e[17];                                     // input columns ci-4 .. ci+12
        e[4] = a.x; e[5] = a.y; e[6] = a.z; e[7] = a.w; e[8] = b.x; e[9] = b.y; e[10] = b.z; e[11] = b.w;
        e[0] = __shfl_up_sync(FULLMASK, b.x, 1); e[1] = __shfl_up_sync(FULLMASK, b.y, 1);
        e[2] = __shfl_up_sync(FULLMASK, b.z, 1); e[3] = __shfl_up_sync(FULLMASK, b.w, 1);
        e[12] = __shfl_down_sync(FULLMASK, a.x, 1); e[13] = __shfl_down_sync(FULLMASK, a.y, 1);
        e[14] = __shfl_down_sync(FULLMASK, a.z, 1); e[15] = __shfl_down_sync(FULLMASK, a.w, 1);
        e[16] = __shfl_down_sync(FULLMASK, b.x, 1);
        if (lane == 0) {
            const float4 l = load_f32_quad(row, ci - 4, W, fastl);
            e[0] = l.x; e[1] = l.y; e[2] = l.z; e[3] = l.w;
        } else if (lane == 31) {
            const float4 r0 = load_f32_quad(row, ci + 8, W, fastr0);
            e[12] = r0.x; e[13] = r0.y; e[14] = r0.z; e[15] = r0.w;
            e[16] = fastr1 ? row[ci + 12] : row[klt_reflect(ci + 12, W)];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {                    // output column X+i is centred on input column ci + 2i + 1
            float acc = T.p[0] * e[2 * i];               // e index of (ci + 2i + 1 - 5) = 2i
#pragma unroll
            for (int j = 1; j < 11; j++) acc = fmaf(T.p[j], e[2 * i + j], acc);
            h[i] = acc;
        }
    };

    // c[j] multiplies in[2Y+1 + j - 5]; input row r contributes to output Y with j = r - 2Y + 4
    const int j0 = ys - 3;                               // first pair index: even row 2*j0 completes output j0-3 (discarded)
    for (int j = j0; j < ye + 3; j++) {
        float he[4], ho[4];
        hrow(2 * j, he);
        hrow(2 * j + 1, ho);
        const int Y = j - 3;
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // even row 2j: outputs Y=j-3..j+2 with j_tap = 10, 8, 6, 4, 2, 0
            o[i] = fmaf(T.p[10], he[i], P[i][0]);
            const float t0 = fmaf(T.p[8], he[i], P[i][1]);
            const float t1 = fmaf(T.p[6], he[i], P[i][2]);
            const float t2 = fmaf(T.p[4], he[i], P[i][3]);
            const float t3 = fmaf(T.p[2], he[i], P[i][4]);
            const float t4 = T.p[0] * he[i];
            // odd row 2j+1: outputs Y=j-2..j+2 with j_tap = 9, 7, 5, 3, 1
            P[i][0] = fmaf(T.p[9], ho[i], t0);
            P[i][1] = fmaf(T.p[7], ho[i], t1);
            P[i][2] = fmaf(T.p[5], ho[i], t2);
            P[i][3] = fmaf(T.p[3], ho[i], t3);
            P[i][4] = fmaf(T.p[1], ho[i], t4);
        }
        if (Y >= ys && Y < ye && X < OW) store_quad(dst + (size_t)Y * out_pitch, X, OW, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Segment height: fill the GPU with (at most) ONE wave of resident CTAs, so that equal-sized segments finish together
// and the warm-up rows (2*(radius) per segment) stay a small fraction of the work.
template <typename K>
static int pick_rows_per_seg(klt_ctx *ctx, K kernel, int H, int strip_ctas, int batch, int min_rows) {
    int per_sm = 4;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, WARPS_PER_CTA * 32, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 4;
    }
    const long capacity = (long)ctx->num_sms * per_sm;
    const long columns = (long)strip_ctas * batch;
    long nseg = capacity / columns;
    if (nseg < 1) nseg = 1;
    long rows = (H + nseg - 1) / nseg;
    if (rows < min_rows) rows = min_rows;
    if (rows > H) rows = H;
    return (int)rows;
}

static bool fill_taps(const klt_kernel1d *k, float *dst, int cap) {
    if (!k || k->n > cap || !(k->n & 1)) return false;
    const int pad = (cap - k->n) / 2;                   // centre shorter kernels inside the fixed-radius slot
    for (int j = 0; j < cap; j++) dst[j] = 0.f;
    for (int j = 0; j < k->n; j++) dst[pad + j] = (float)k->taps[k->n - 1 - j];
    return true;
}
static bool is_symmetric(const klt_kernel1d *k) {
    for (int i = 0; i < k->n / 2; i++)
        if (fabs(k->taps[i] - k->taps[k->n - 1 - i]) > 2.220446049250313e-16) return false;
    return true;
}

// Returns 1 if the streaming kernel was launched, 0 if the configuration is not covered (caller falls back), <0 on error.
int klt_stream_level0(klt_ctx *ctx, const uint8_t *frames, size_t pitch, size_t frame_stride, klt_pyr *p,
                      const klt_taps *taps) {
    StreamTaps T;
    const int ns = taps->smooth.n;
    if (ns != 3 && ns != 5 && ns != 7 && ns != 9) return 0;
    if (!is_symmetric(&taps->smooth)) return 0;          // the reflect-extension argument needs a symmetric smoother
    const int RS = ns / 2;
    if (!fill_taps(&taps->smooth, T.s, ns)) return 0;
    if (!fill_taps(&taps->grad_gauss, T.g, 7) || !fill_taps(&taps->grad_deriv, T.d, 7)) return 0;
    for (int j = 0; j < 11; j++) T.p[j] = 0.f;
    const int W = p->w, H = p->h;
    if (W < 16 || H < 16) return 0;
    const int n_strips = (W + 119) / 120;
    const int strip_ctas = (n_strips + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    int rows = 0;
    switch (RS) {
        case 1: rows = pick_rows_per_seg(ctx, stream_level0_kernel<1>, H, strip_ctas, p->batch, 32); break;
        case 2: rows = pick_rows_per_seg(ctx, stream_level0_kernel<2>, H, strip_ctas, p->batch, 32); break;
        case 3: rows = pick_rows_per_seg(ctx, stream_level0_kernel<3>, H, strip_ctas, p->batch, 32); break;
        default: rows = pick_rows_per_seg(ctx, stream_level0_kernel<4>, H, strip_ctas, p->batch, 32); break;
    }
    dim3 grid(strip_ctas, (H + rows - 1) / rows, p->batch), block(WARPS_PER_CTA * 32);
    const double bytes = 13.0 * W * H * p->batch;        // 1 B read + 3 x 4 B written per pixel
    float *img = p->level(0, 0, 0), *gx = p->level(1, 0, 0), *gy = p->level(2, 0, 0);
#define LAUNCH_L0(R)                                                                                                   \
    KLT_LAUNCH(ctx, "stream_level0", bytes,                                                                            \
               (stream_level0_kernel<R><<<grid, block, 0, ctx->stream>>>(frames, pitch, frame_stride, img, gx, gy,      \
                                                                         p->lv[0].pitch, p->plane_floats, W, H, rows,  \
                                                                         n_strips, T)))
    switch (RS) {
        case 1: LAUNCH_L0(1); break;
        case 2: LAUNCH_L0(2); break;
        case 3: LAUNCH_L0(3); break;
        case 4: LAUNCH_L0(4); break;
        default: return 0;
    }
#undef LAUNCH_L0
    return 1;
}

int klt_stream_grad(klt_ctx *ctx, klt_pyr *p, int level, const klt_taps *taps) {
    StreamTaps T;
    if (!fill_taps(&taps->grad_gauss, T.g, 7) || !fill_taps(&taps->grad_deriv, T.d, 7)) return 0;
    for (int j = 0; j < 9; j++) T.s[j] = 0.f;
    for (int j = 0; j < 11; j++) T.p[j] = 0.f;
    const LevelDesc &a = p->lv[level];
    if (a.w < 16 || a.h < 8) return 0;
    const int n_strips = (a.w + 119) / 120;
    const int strip_ctas = (n_strips + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int rows = pick_rows_per_seg(ctx, stream_grad_kernel, a.h, strip_ctas, p->batch, 24);
    dim3 grid(strip_ctas, (a.h + rows - 1) / rows, p->batch), block(WARPS_PER_CTA * 32);
    const double bytes = 12.0 * a.w * a.h * p->batch;
    KLT_LAUNCH(ctx, "stream_grad", bytes,
               (stream_grad_kernel<<<grid, block, 0, ctx->stream>>>(p->level(0, 0, level), a.pitch, p->plane_floats,
                                                                    p->level(1, 0, level), p->level(2, 0, level), a.pitch,
                                                                    p->plane_floats, a.w, a.h, rows, n_strips, T)));
    return 1;
}

int klt_stream_down2(klt_ctx *ctx, klt_pyr *p, int level, const klt_taps *taps) {
    if (p->ss != 2 || level < 1) return 0;
    StreamTaps T;
    if (taps->pyramid.n > 11 || !fill_taps(&taps->pyramid, T.p, 11)) return 0;
    for (int j = 0; j < 9; j++) T.s[j] = 0.f;
    for (int j = 0; j < 7; j++) { T.g[j] = 0.f; T.d[j] = 0.f; }
    const LevelDesc &a = p->lv[level - 1], &b = p->lv[level];
    if (b.w < 8 || b.h < 8 || a.w < 16) return 0;
    const int n_strips = (b.w + 127) / 128;
    const int strip_ctas = (n_strips + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const int rows = pick_rows_per_seg(ctx, stream_down2_kernel, b.h, strip_ctas, p->batch, 24);
    dim3 grid(strip_ctas, (b.h + rows - 1) / rows, p->batch), block(WARPS_PER_CTA * 32);
    const double bytes = 4.0 * ((double)a.w * a.h + (double)b.w * b.h) * p->batch;
    KLT_LAUNCH(ctx, "stream_down2", bytes,
               (stream_down2_kernel<<<grid, block, 0, ctx->stream>>>(p->level(0, 0, level - 1), a.pitch, p->plane_floats, a.w,
                                                                     a.h, p->level(0, 0, level), b.pitch, p->plane_floats, b.w,
                                                                     b.h, rows, n_strips, T)));
    return 1;
}
