// Ceilings for the traffic MIXES of the streaming kernels, moved by the simplest possible kernels (grid-stride, linear,
// nothing computed): what can this B200 sustain for 1 B read + 4 B written per pixel (stream_smooth0) and for 16 B read + 4 B
// written per output pixel (stream_down2), against a 1:1 copy, a pure write and a pure read?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mix_probe tools/mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void copy_k(const float4 *in, float4 *out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void write_k(float4 *out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void read_k(const float4 *in, float *sink, size_t n) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = __ldg(in + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 12345.678f) *sink = acc;
}
// smooth0's mix: one u8 quad in, one float4 out
__global__ void mix14_k(const unsigned int *in, float4 *out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned int w = __ldg(in + i);
        out[i] = make_float4((float)(w & 255), (float)((w >> 8) & 255), (float)((w >> 16) & 255), (float)(w >> 24));
    }
}
// down2's mix: four float4 in, one float4 out
__global__ void mix41_k(const float4 *in, float4 *out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(in + 4 * i), b = __ldg(in + 4 * i + 1), c = __ldg(in + 4 * i + 2), d = __ldg(in + 4 * i + 3);
        out[i] = make_float4(a.x + b.x + c.x + d.x, a.y + b.y + c.y + d.y, a.z + b.z + c.z + d.z, a.w + b.w + c.w + d.w);
    }
}
int main() {
    const size_t G = (size_t)1 << 30;
    float4 *a, *b; float *sink;
    if (cudaMalloc(&a, 4 * G) != cudaSuccess || cudaMalloc(&b, 4 * G) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(a, 1, 4 * G); cudaMemset(b, 0, 4 * G);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int grid = 148 * 16;
    for (int rep = 0; rep < 3; rep++) {
#define T(name, bytes, launch) cudaEventRecord(e0); launch; cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); \
        if (rep == 2) printf("{\"probe\": \"%s\", \"ms\": %.4f, \"algorithmic_gbps\": %.1f}\n", name, ms, (double)(bytes) / ms / 1e6);
        T("copy 4 GiB + 4 GiB", 8.0 * G, (copy_k<<<grid, 256>>>(a, b, 4 * G / 16)))
        T("write 4 GiB", 4.0 * G, (write_k<<<grid, 256>>>(b, 4 * G / 16)))
        T("read 4 GiB", 4.0 * G, (read_k<<<grid, 256>>>(a, sink, 4 * G / 16)))
        T("smooth0 mix: 1 GiB u8 in, 4 GiB f32 out", 5.0 * G, (mix14_k<<<grid, 256>>>((const unsigned int *)a, b, G / 4)))
        T("down2 mix: 4 GiB in, 1 GiB out", 5.0 * G, (mix41_k<<<grid, 256>>>(a, b, G / 16)))
        // the sizes of one 64-frame launch (664 MB / 830 MB), as bench.py's kernels see them
        T("smooth0 mix, 133 Mpx (one 64 x 1080p launch)", 5.0 * 132710400, (mix14_k<<<grid, 256>>>((const unsigned int *)a, b, 132710400 / 4)))
        T("down2 mix, 33 M output px (one 64-frame level-1 launch)", 20.0 * 33177600, (mix41_k<<<grid, 256>>>(a, b, 33177600 / 4)))
    }
    return cudaGetLastError() != cudaSuccess;
}
