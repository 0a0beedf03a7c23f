// Bandwidth probe for the traffic mix of the level-0 kernel: how fast can B200 move 1 byte in + 12 bytes out per pixel
// when nothing is computed?  (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bw_probe bw_probe.cu)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void copy_k(const float4 *in, float4 *out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void write_k(float4 *out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void read_k(const float4 *in, float *sink, size_t n) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = in[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 12345.678f) *sink = acc;
}
// same geometry as stream_level0_kernel: warp = 120-column strip (lanes 1..30 write), marches down rows_per_seg rows,
// reads one u8 quad per lane-row and writes three float4 per lane-row into three planes
template <int WPC, bool ALIGNED = false>
__global__ void __launch_bounds__(WPC * 32, 16 / WPC)
mix_k(const unsigned char *frames, float *o0, float *o1, float *o2, int W, int H, int rows, int n_strips, size_t fstride, size_t ostride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * WPC + warp;
    if (strip >= n_strips) return;
    const int ys = blockIdx.y * rows, ye = min(H, ys + rows);
    const int c = ALIGNED ? strip * 128 + 4 * lane : strip * 120 + 4 * (lane - 1);
    if (!ALIGNED && (lane < 1 || lane > 30)) return;
    if (c >= W) return;
    const unsigned char *src = frames + blockIdx.z * fstride + c;
    const size_t ob = blockIdx.z * ostride + c;
    // loads run 8 rows ahead of the stores (register ring), so the loop is not bound by one load latency per row
    unsigned int q[8];
#pragma unroll
    for (int k = 0; k < 8; k++) q[k] = __ldg(reinterpret_cast<const unsigned int *>(src + (size_t)min(ys + k, ye - 1) * W));
    for (int y = ys; y < ye; y += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (y + k >= ye) break;
            const unsigned int w = q[k];
            q[k] = __ldg(reinterpret_cast<const unsigned int *>(src + (size_t)min(y + k + 8, ye - 1) * W));
            const float4 v = make_float4((float)(w & 255), (float)((w >> 8) & 255), (float)((w >> 16) & 255), (float)(w >> 24));
            *reinterpret_cast<float4 *>(o0 + ob + (size_t)(y + k) * W) = v;
            *reinterpret_cast<float4 *>(o1 + ob + (size_t)(y + k) * W) = v;
            *reinterpret_cast<float4 *>(o2 + ob + (size_t)(y + k) * W) = v;
        }
    }
}

int main() {
    const size_t N = (size_t)1 << 28;   // 256 Mi float4 = 4 GiB per buffer
    const size_t bytes = ((size_t)1 << 30);
    float4 *a, *b; float *sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
    (void)N;
    const size_t n4 = bytes / 16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); copy_k<<<148 * 16, 256>>>(a, b, n4); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("copy  (1 GiB read + 1 GiB write): %.1f GB/s\n", 2.0 * bytes / ms / 1e6);
        cudaEventRecord(e0); write_k<<<148 * 16, 256>>>(b, n4); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("write (1 GiB):                    %.1f GB/s\n", 1.0 * bytes / ms / 1e6);
        cudaEventRecord(e0); read_k<<<148 * 16, 256>>>(a, sink, n4); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 2) printf("read  (1 GiB):                    %.1f GB/s\n", 1.0 * bytes / ms / 1e6);
    }
    // level-0 traffic mix: 32 frames 1920x1080
    const int W = 1920, H = 1080, B = 32;
    unsigned char *fr; float *o;
    CK(cudaMalloc(&fr, (size_t)B * W * H)); CK(cudaMalloc(&o, (size_t)3 * B * W * H * 4));
    const size_t plane = (size_t)B * W * H;
    for (int rows : {120, 60, 30}) {
        for (int wpc : {4, 8, 16}) {
            dim3 grid(16 / wpc, (H + rows - 1) / rows, B);
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (wpc == 4) mix_k<4><<<grid, 128>>>(fr, o, o + plane, o + 2 * plane, W, H, rows, 16, (size_t)W * H, (size_t)W * H);
                else if (wpc == 8) mix_k<8><<<grid, 256>>>(fr, o, o + plane, o + 2 * plane, W, H, rows, 16, (size_t)W * H, (size_t)W * H);
                else mix_k<16><<<grid, 512>>>(fr, o, o + plane, o + 2 * plane, W, H, rows, 16, (size_t)W * H, (size_t)W * H);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
                if (rep == 2) printf("level-0 mix (1 B in + 12 B out per px, %d frames, rows/seg %4d, %2d warps/CTA): %.3f ms  %.1f GB/s\n", B, rows, wpc, ms, 13.0 * B * W * H / ms / 1e6);
            }
        }
    }
    for (int rows : {120, 60}) {
        dim3 grid(4, (H + rows - 1) / rows, B);     // 15 strips of 128 columns = 1920
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            mix_k<4, true><<<grid, 128>>>(fr, o, o + plane, o + 2 * plane, W, H, rows, 15, (size_t)W * H, (size_t)W * H);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) printf("level-0 mix, 128-column strips (512 B aligned stores, all lanes), rows/seg %4d: %.3f ms  %.1f GB/s\n", rows, ms, 13.0 * B * W * H / ms / 1e6);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
