// Traffic-mix ceilings (diagnostics; klt_probe_traffic_mix in klt_b200.h).
//
// The pyramid kernels are bound by HBM, and the reference figure bench.py divides by is a 1:1 copy (MEASURED_PEAKS.json).  What
// HBM3e sustains depends on the read/write mix, though (measured on this part, 4 GiB transfers under ncu: pure read 6.9 TB/s,
// pure write 6.7, 1:1 copy 6.15, 4:1 reads 6.2, 1:4 writes 5.6) and on the launch size (a 0.7 GB launch ramps up and drains).
// These kernels move a given mix with the simplest possible code -- grid-stride, linear, 128/256-bit accesses, nothing computed
// -- so that a kernel's achieved bandwidth can be stated against what its own mix and size allow on the box it runs on.
#include "klt_common.cuh"

namespace {

// unit = 8 "pixels": RQ 8-byte reads from `in`, W0 32-byte writes to `out0`, W1 8-byte writes to `out1` per unit
//   smooth0 (1 B in, 4 B out per pixel):            RQ = 1, W0 = 1, W1 = 0, unit = 8 level-0 pixels
//   level01 (1 B in, 4 B + 1 B out per pixel):      RQ = 1, W0 = 1, W1 = 1
//   down2   (16 B in, 4 B out per output pixel):    reads 4 x 32 B, writes 32 B: unit = 8 output pixels, kind handled below
__global__ void __launch_bounds__(256) mix_u8_kernel(const uint2 *__restrict__ in, float *__restrict__ out0, float2 *__restrict__ out1,
                                                     size_t units, int with_out1) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < units; i += (size_t)gridDim.x * blockDim.x) {
        const uint2 w = __ldg(in + i);
        const float a = (float)(w.x & 255u), b = (float)((w.x >> 8) & 255u), c = (float)((w.x >> 16) & 255u), d = (float)(w.x >> 24);
        const float e = (float)(w.y & 255u), f = (float)((w.y >> 8) & 255u), g = (float)((w.y >> 16) & 255u), h = (float)(w.y >> 24);
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out0 + 8 * i), "f"(a), "f"(b), "f"(c), "f"(d),
                     "f"(e), "f"(f), "f"(g), "f"(h)
                     : "memory");
        if (with_out1) out1[i] = make_float2(a + e, b + f);
    }
}
__global__ void __launch_bounds__(256) mix_f32_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, size_t units, int reads) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < units; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < reads; k++) {
            const float4 v = __ldg(in + (size_t)reads * i + k);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out[i] = acc;
    }
}

}  // namespace

extern "C" int klt_probe_traffic_mix(klt_ctx *ctx, int kind, double total_bytes, int reps, double *ms_per_rep, double *bytes_moved) {
    if (!ctx || !ms_per_rep || reps < 1 || total_bytes < 1e6 || total_bytes > 16e9) return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    KLT_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t units, in_bytes, out0_bytes, out1_bytes = 0;
    switch (kind) {
        case KLT_MIX_SMOOTH0: units = (size_t)(total_bytes / 40.0); in_bytes = units * 8; out0_bytes = units * 32; break;
        case KLT_MIX_LEVEL01: units = (size_t)(total_bytes / 48.0); in_bytes = units * 8; out0_bytes = units * 32; out1_bytes = units * 8; break;
        case KLT_MIX_DOWN2: units = (size_t)(total_bytes / 80.0); in_bytes = units * 64; out0_bytes = units * 16; break;
        case KLT_MIX_COPY: units = (size_t)(total_bytes / 32.0); in_bytes = units * 16; out0_bytes = units * 16; break;
        default: return klt_fail(ctx, KLT_ERR_INVALID, "unknown traffic mix %d", kind);
    }
    void *in = nullptr, *out0 = nullptr, *out1 = nullptr;
    cudaError_t e = cudaMalloc(&in, in_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&out0, out0_bytes);
    if (e == cudaSuccess && out1_bytes) e = cudaMalloc(&out1, out1_bytes);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaMemsetAsync(in, 1, in_bytes, ctx->stream);
    float ms = 0.f;
    if (e == cudaSuccess) {
        const int grid = ctx->num_sms * 16;
        for (int r = -2; r < reps; r++) {                // two warm-up launches, then `reps` launches between the events
            if (r == 0) cudaEventRecord(e0, ctx->stream);
            if (kind == KLT_MIX_SMOOTH0 || kind == KLT_MIX_LEVEL01)
                mix_u8_kernel<<<grid, 256, 0, ctx->stream>>>((const uint2 *)in, (float *)out0, (float2 *)out1, units, kind == KLT_MIX_LEVEL01);
            else
                mix_f32_kernel<<<grid, 256, 0, ctx->stream>>>((const float4 *)in, (float4 *)out0, units, kind == KLT_MIX_DOWN2 ? 4 : 1);
        }
        cudaEventRecord(e1, ctx->stream);
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(in); cudaFree(out0); cudaFree(out1);
    if (e != cudaSuccess) return klt_fail(ctx, KLT_ERR_CUDA, "traffic-mix probe failed: %s", cudaGetErrorString(e));
    *ms_per_rep = ms / reps;
    if (bytes_moved) *bytes_moved = (double)(in_bytes + out0_bytes + out1_bytes);
    return KLT_OK;
}
