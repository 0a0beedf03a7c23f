// Separable convolution, gradient-pair and pyramid-decimation kernels (sm_100a).
//
// Replaces scipy.ndimage.convolve1d as called from convolve.py:212-213 (reference), i.e. the arithmetic of
// KLTComputeSmoothedImage (convolve.py:254-264), KLTComputeGradients (:226-248) and KLTPyramid.Compute
// (pyramid.py:59-77).  Border handling is SciPy's default mode='reflect' (half-sample symmetric, quirk Q7).
//
// Two arithmetic modes (klt_b200.h):
//   FAST   : fp32 FMA accumulation; every 1-D pass rounds to fp32 like the reference does between passes.
//   STRICT : SciPy's exact recipe -- inputs widened to double, accumulation in double in NI_Correlate1D's
//            operation order (folded symmetric / antisymmetric form), no FMA contraction, one rounding to
//            fp32 per 1-D pass.  Bit-identical to the reference's images.
#include "klt_common.cuh"

template <bool STRICT> struct TapsSel { typedef TapsF type; };
template <> struct TapsSel<true> { typedef TapsD type; };

// a points at the centre sample; stride between neighbouring samples along the filtered axis
__device__ __forceinline__ float apply_taps(const float *a, int stride, const TapsF &t) {
    float o = 0.f;
    const int r = t.r;
#pragma unroll 4
    for (int j = 0; j < t.n; j++) o = fmaf(a[(j - r) * stride], t.c[j], o);
    return o;
}
__device__ __forceinline__ float apply_taps(const float *a, int stride, const TapsD &t) {
    const int r = t.r;
    double o;
    if (t.sym > 0) {
        o = __dmul_rn((double)a[0], t.c[r]);
        for (int jj = -r; jj < 0; jj++)
            o = __dadd_rn(o, __dmul_rn(__dadd_rn((double)a[jj * stride], (double)a[-jj * stride]), t.c[jj + r]));
    } else if (t.sym < 0) {
        o = __dmul_rn((double)a[0], t.c[r]);
        for (int jj = -r; jj < 0; jj++)
            o = __dadd_rn(o, __dmul_rn(__dsub_rn((double)a[jj * stride], (double)a[-jj * stride]), t.c[jj + r]));
    } else {
        o = __dmul_rn((double)a[r * stride], t.c[2 * r]);
        for (int jj = -r; jj < r; jj++) o = __dadd_rn(o, __dmul_rn((double)a[jj * stride], t.c[jj + r]));
    }
    return __double2float_rn(o);
}

// ---------------------------------------------------------------------------------------------------
// Generic tiled separable convolution: out = V_vk( H_hk(in) ).  One CTA = one TW x TH output tile of one
// image; the (TH+2rv) x (TW+2rh) input region is staged in shared memory with reflected indices.
// ---------------------------------------------------------------------------------------------------
template <typename InT, bool STRICT>
__global__ void __launch_bounds__(256)
conv_sep_kernel(const InT *__restrict__ in, size_t in_pitch, size_t in_stride, float *__restrict__ out,
                size_t out_pitch, size_t out_stride, int W, int H, int TW, int TH,
                const __grid_constant__ typename TapsSel<STRICT>::type hk,
                const __grid_constant__ typename TapsSel<STRICT>::type vk) {
    extern __shared__ float smem[];
    const int rh = hk.r, rv = vk.r;
    const int RW = TW + 2 * rh, RH = TH + 2 * rv;
    float *s_in = smem;
    float *s_tmp = smem + (size_t)RH * RW;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const InT *src = in + (size_t)blockIdx.z * in_stride;
    for (int ry = ty; ry < RH; ry += blockDim.y) {
        const InT *row = src + (size_t)klt_reflect(y0 - rv + ry, H) * in_pitch;
        for (int rx = tx; rx < RW; rx += blockDim.x) s_in[ry * RW + rx] = (float)row[klt_reflect(x0 - rh + rx, W)];
    }
    __syncthreads();
    for (int ry = ty; ry < RH; ry += blockDim.y)
        for (int x = tx; x < TW; x += blockDim.x) s_tmp[ry * TW + x] = apply_taps(s_in + ry * RW + x + rh, 1, hk);
    __syncthreads();
    float *dst = out + (size_t)blockIdx.z * out_stride;
    for (int y = ty; y < TH; y += blockDim.y) {
        if (y0 + y >= H) break;
        for (int x = tx; x < TW; x += blockDim.x)
            if (x0 + x < W) dst[(size_t)(y0 + y) * out_pitch + x0 + x] = apply_taps(s_tmp + (y + rv) * TW + x, TW, vk);
    }
}

// gx = V_g( H_d(in) ), gy = V_d( H_g(in) )  (convolve.py:245-246); the input tile is read once.
template <bool STRICT>
__global__ void __launch_bounds__(256)
grad_pair_kernel(const float *__restrict__ in, size_t in_pitch, size_t in_stride, float *__restrict__ gxo,
                 float *__restrict__ gyo, size_t out_pitch, size_t out_stride, int W, int H, int TW, int TH,
                 const __grid_constant__ typename TapsSel<STRICT>::type g,
                 const __grid_constant__ typename TapsSel<STRICT>::type d) {
    extern __shared__ float smem[];
    const int R = max(g.r, d.r);
    const int RW = TW + 2 * R, RH = TH + 2 * R;
    float *s_in = smem;
    float *s_d = s_in + (size_t)RH * RW;
    float *s_g = s_d + (size_t)RH * TW;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const float *src = in + (size_t)blockIdx.z * in_stride;
    for (int ry = ty; ry < RH; ry += blockDim.y) {
        const float *row = src + (size_t)klt_reflect(y0 - R + ry, H) * in_pitch;
        for (int rx = tx; rx < RW; rx += blockDim.x) s_in[ry * RW + rx] = row[klt_reflect(x0 - R + rx, W)];
    }
    __syncthreads();
    for (int ry = ty; ry < RH; ry += blockDim.y)
        for (int x = tx; x < TW; x += blockDim.x) {
            const float *a = s_in + ry * RW + x + R;
            s_d[ry * TW + x] = apply_taps(a, 1, d);
            s_g[ry * TW + x] = apply_taps(a, 1, g);
        }
    __syncthreads();
    const size_t ob = (size_t)blockIdx.z * out_stride;
    for (int y = ty; y < TH; y += blockDim.y) {
        if (y0 + y >= H) break;
        for (int x = tx; x < TW; x += blockDim.x)
            if (x0 + x < W) {
                size_t o = ob + (size_t)(y0 + y) * out_pitch + x0 + x;
                gxo[o] = apply_taps(s_d + (y + R) * TW + x, TW, g);
                gyo[o] = apply_taps(s_g + (y + R) * TW + x, TW, d);
            }
    }
}

// Pyramid step (pyramid.py:59-72): smooth then keep pixel (ss*y+ss/2, ss*x+ss/2).  Only the sampled
// columns are filtered horizontally and only the sampled rows vertically; rounding between the passes is
// the reference's (fp32 after each 1-D pass).
template <bool STRICT>
__global__ void __launch_bounds__(256)
pyr_down_kernel(const float *__restrict__ in, size_t in_pitch, size_t in_stride, int W, int H,
                float *__restrict__ out, size_t out_pitch, size_t out_stride, int OW, int OH, int ss, int TW, int TH,
                const __grid_constant__ typename TapsSel<STRICT>::type g) {
    extern __shared__ float smem[];
    const int r = g.r;
    const int RW = ss * (TW - 1) + 1 + 2 * r, RH = ss * (TH - 1) + 1 + 2 * r;
    float *s_in = smem;
    float *s_tmp = s_in + (size_t)RH * RW;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int ix0 = ss * ox0 + ss / 2 - r, iy0 = ss * oy0 + ss / 2 - r;
    const float *src = in + (size_t)blockIdx.z * in_stride;
    for (int ry = ty; ry < RH; ry += blockDim.y) {
        const float *row = src + (size_t)klt_reflect(iy0 + ry, H) * in_pitch;
        for (int rx = tx; rx < RW; rx += blockDim.x) s_in[ry * RW + rx] = row[klt_reflect(ix0 + rx, W)];
    }
    __syncthreads();
    for (int ry = ty; ry < RH; ry += blockDim.y)
        for (int x = tx; x < TW; x += blockDim.x) s_tmp[ry * TW + x] = apply_taps(s_in + ry * RW + ss * x + r, 1, g);
    __syncthreads();
    float *dst = out + (size_t)blockIdx.z * out_stride;
    for (int y = ty; y < TH; y += blockDim.y) {
        if (oy0 + y >= OH) break;
        for (int x = tx; x < TW; x += blockDim.x)
            if (ox0 + x < OW)
                dst[(size_t)(oy0 + y) * out_pitch + ox0 + x] = apply_taps(s_tmp + (ss * y + r) * TW + x, TW, g);
    }
}

// ---------------------------------------------------------------------------------------------------
int klt_make_taps(klt_ctx *ctx, const klt_kernel1d *k, TapsF *f, TapsD *d) {
    if (!k || k->n < 1 || k->n > KLT_MAX_TAPS) return klt_fail(ctx, KLT_ERR_INVALID, "kernel length %d out of range", k ? k->n : -1);
    if ((k->n & 1) == 0) return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "even kernel length %d (reference kernels are odd, convolve.py:118)", k->n);
    const int n = k->n, r = n / 2;
    int sym = 1;
    for (int ii = 1; ii <= r; ii++)
        if (fabs(k->taps[r + ii] - k->taps[r - ii]) > 2.220446049250313e-16) { sym = 0; break; }
    if (!sym) {
        sym = -1;
        for (int ii = 1; ii <= r; ii++)
            if (fabs(k->taps[r + ii] + k->taps[r - ii]) > 2.220446049250313e-16) { sym = 0; break; }
    }
    f->n = d->n = n; f->r = d->r = r; f->sym = d->sym = sym; d->pad = 0;
    for (int j = 0; j < KLT_MAX_TAPS; j++) { f->c[j] = 0.f; d->c[j] = 0.0; }
    for (int j = 0; j < n; j++) { d->c[j] = k->taps[n - 1 - j]; f->c[j] = (float)k->taps[n - 1 - j]; }
    return KLT_OK;
}

static const size_t kMaxSmem = 200 * 1024;

template <typename K> static int set_smem(klt_ctx *ctx, K kernel) {
    KLT_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    return KLT_OK;
}

template <typename InT>
static int launch_conv_sep(klt_ctx *ctx, const InT *in, size_t in_pitch, size_t in_stride, float *out, size_t out_pitch,
                           size_t out_stride, int w, int h, int batch, const klt_kernel1d *hk, const klt_kernel1d *vk,
                           int precision) {
    TapsF hf, vf; TapsD hd, vd;
    int rc;
    if ((rc = klt_make_taps(ctx, hk, &hf, &hd))) return rc;
    if ((rc = klt_make_taps(ctx, vk, &vf, &vd))) return rc;
    int TW = 64, TH = 32;
    auto smem_for = [&](int tw, int th) { return ((size_t)(th + 2 * vf.r) * (tw + 2 * hf.r) + (size_t)(th + 2 * vf.r) * tw) * sizeof(float); };
    while (smem_for(TW, TH) > kMaxSmem && TH > 8) TH /= 2;
    while (smem_for(TW, TH) > kMaxSmem && TW > 32) TW /= 2;
    size_t smem = smem_for(TW, TH);
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, batch), block(32, 8);
    const double bytes = (double)(sizeof(InT) + sizeof(float)) * w * h * batch;
    const bool u8 = sizeof(InT) == 1;
    if (precision == KLT_PRECISION_STRICT) {
        if ((rc = set_smem(ctx, conv_sep_kernel<InT, true>))) return rc;
        KLT_LAUNCH(ctx, u8 ? "conv_sep_u8_strict" : "conv_sep_f32_strict", bytes,
                   (conv_sep_kernel<InT, true><<<grid, block, smem, ctx->stream>>>(in, in_pitch, in_stride, out, out_pitch, out_stride, w, h, TW, TH, hd, vd)));
    } else {
        if ((rc = set_smem(ctx, conv_sep_kernel<InT, false>))) return rc;
        KLT_LAUNCH(ctx, u8 ? "conv_sep_u8" : "conv_sep_f32", bytes,
                   (conv_sep_kernel<InT, false><<<grid, block, smem, ctx->stream>>>(in, in_pitch, in_stride, out, out_pitch, out_stride, w, h, TW, TH, hf, vf)));
    }
    return KLT_OK;
}

int klt_launch_conv_sep_f32(klt_ctx *ctx, const float *in, size_t in_pitch, size_t in_stride, float *out, size_t out_pitch,
                            size_t out_stride, int w, int h, int batch, const klt_kernel1d *hk, const klt_kernel1d *vk,
                            int precision) {
    return launch_conv_sep<float>(ctx, in, in_pitch, in_stride, out, out_pitch, out_stride, w, h, batch, hk, vk, precision);
}
int klt_launch_conv_sep_u8(klt_ctx *ctx, const uint8_t *in, size_t in_pitch, size_t in_stride, float *out, size_t out_pitch,
                           size_t out_stride, int w, int h, int batch, const klt_kernel1d *hk, const klt_kernel1d *vk,
                           int precision) {
    return launch_conv_sep<uint8_t>(ctx, in, in_pitch, in_stride, out, out_pitch, out_stride, w, h, batch, hk, vk, precision);
}

int klt_launch_grad_pair(klt_ctx *ctx, const float *in, size_t in_pitch, size_t in_stride, float *gx, float *gy,
                         size_t out_pitch, size_t out_stride, int w, int h, int batch, const klt_kernel1d *g,
                         const klt_kernel1d *d, int precision) {
    TapsF gf, df; TapsD gd, dd;
    int rc;
    if ((rc = klt_make_taps(ctx, g, &gf, &gd))) return rc;
    if ((rc = klt_make_taps(ctx, d, &df, &dd))) return rc;
    const int R = gf.r > df.r ? gf.r : df.r;
    int TW = 64, TH = 32;
    auto smem_for = [&](int tw, int th) { return ((size_t)(th + 2 * R) * (tw + 2 * R) + 2 * (size_t)(th + 2 * R) * tw) * sizeof(float); };
    while (smem_for(TW, TH) > kMaxSmem && TH > 8) TH /= 2;
    while (smem_for(TW, TH) > kMaxSmem && TW > 32) TW /= 2;
    size_t smem = smem_for(TW, TH);
    dim3 grid((w + TW - 1) / TW, (h + TH - 1) / TH, batch), block(32, 8);
    const double bytes = 12.0 * w * h * batch;
    if (precision == KLT_PRECISION_STRICT) {
        if ((rc = set_smem(ctx, grad_pair_kernel<true>))) return rc;
        KLT_LAUNCH(ctx, "grad_pair_strict", bytes,
                   (grad_pair_kernel<true><<<grid, block, smem, ctx->stream>>>(in, in_pitch, in_stride, gx, gy, out_pitch, out_stride, w, h, TW, TH, gd, dd)));
    } else {
        if ((rc = set_smem(ctx, grad_pair_kernel<false>))) return rc;
        KLT_LAUNCH(ctx, "grad_pair", bytes,
                   (grad_pair_kernel<false><<<grid, block, smem, ctx->stream>>>(in, in_pitch, in_stride, gx, gy, out_pitch, out_stride, w, h, TW, TH, gf, df)));
    }
    return KLT_OK;
}

int klt_launch_pyr_down(klt_ctx *ctx, const float *in, size_t in_pitch, size_t in_stride, int w, int h, float *out,
                        size_t out_pitch, size_t out_stride, int ow, int oh, int ss, int batch, const klt_kernel1d *g,
                        int precision) {
    TapsF gf; TapsD gd;
    int rc;
    if ((rc = klt_make_taps(ctx, g, &gf, &gd))) return rc;
    const int r = gf.r;
    int TW = 32, TH = 16;
    auto smem_for = [&](int tw, int th) {
        size_t RW = (size_t)ss * (tw - 1) + 1 + 2 * r, RH = (size_t)ss * (th - 1) + 1 + 2 * r;
        return (RH * RW + RH * tw) * sizeof(float);
    };
    while (smem_for(TW, TH) > kMaxSmem && TH > 1) TH /= 2;
    while (smem_for(TW, TH) > kMaxSmem && TW > 1) TW /= 2;
    if (smem_for(TW, TH) > kMaxSmem) return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "pyramid kernel too wide for shared memory");
    size_t smem = smem_for(TW, TH);
    dim3 grid((ow + TW - 1) / TW, (oh + TH - 1) / TH, batch), block(32, 8);
    const double bytes = 4.0 * ((double)w * h + (double)ow * oh) * batch;
    if (precision == KLT_PRECISION_STRICT) {
        if ((rc = set_smem(ctx, pyr_down_kernel<true>))) return rc;
        KLT_LAUNCH(ctx, "pyr_down_strict", bytes,
                   (pyr_down_kernel<true><<<grid, block, smem, ctx->stream>>>(in, in_pitch, in_stride, w, h, out, out_pitch, out_stride, ow, oh, ss, TW, TH, gd)));
    } else {
        if ((rc = set_smem(ctx, pyr_down_kernel<false>))) return rc;
        KLT_LAUNCH(ctx, "pyr_down", bytes,
                   (pyr_down_kernel<false><<<grid, block, smem, ctx->stream>>>(in, in_pitch, in_stride, w, h, out, out_pitch, out_stride, ow, oh, ss, TW, TH, gf)));
    }
    return KLT_OK;
}
