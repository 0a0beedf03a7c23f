// FFMA vs FFMA2 issue / pipe throughput on sm_100a (why the streaming kernels moved to packed fp32 arithmetic).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu && tools/ffma2_probe
// Three kernels, each with 8 independent dependency chains per thread, 512 threads per CTA, 2 CTAs per SM:
//   scalar : 16 FFMA per iteration                      -> FMA/s when every FMA costs one issue slot
//   packed : 8 FFMA2 per iteration (same 16 FMAs)       -> does a packed instruction cost one slot or two?
//   mixed  : 8 FFMA2 + 8 IADD3/LOP3 per iteration       -> are the slots FFMA2 frees usable by other pipes?
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(512, 2) probe(float *out, int iters, float c) {
    float s[16];
    u64 p[8];
    unsigned int k[8];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; i++) { p[i] = pk(s[2 * i], s[2 * i + 1]); k[i] = threadIdx.x + i; }
    const u64 cc = pk(c, c);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %1, %0, %2;" : "+f"(s[i]) : "f"(c), "f"(0.5f));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(cc, p[i], cc);
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(k[i]) : "r"(it));
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) acc += s[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); acc += a + b + (float)k[i]; }
    if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
static void run(const char *name, int sms) {
    float *out;
    cudaMalloc(&out, 4);
    const int iters = 1 << 15;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<sms * 2, 512>>>(out, 64, 0.999f);
    cudaEventRecord(e0);
    probe<MODE><<<sms * 2, 512>>>(out, iters, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = 16.0 * iters * 512.0 * 2 * sms;
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"tfma_per_s\": %.2f}\n", name, ms, fma / ms * 1e-9);
    cudaFree(out);
}

int main() {
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    run<0>("scalar FFMA x16", pr.multiProcessorCount);
    run<1>("packed FFMA2 x8", pr.multiProcessorCount);
    run<2>("packed FFMA2 x8 + 8 integer ops", pr.multiProcessorCount);
    return 0;
}
