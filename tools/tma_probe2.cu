// sweep of 2-D tensor-map parameters: tma_probe2 <box_w> <box_h> <l2promo> <swizzle> <dtype: 0 f32, 1 u32, 2 u8> <pitch_elems>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k_2d(const __grid_constant__ CUtensorMap map, int x, int y, int bytes, unsigned *out) {
    extern __shared__ __align__(1024) unsigned sm[];
    __shared__ alignas(8) unsigned long long mbar;
    const unsigned mb = smem_u32(&mbar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(mb) : "memory");
    }
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mb), "r"(0) : "memory");
    for (int i = threadIdx.x; i < bytes / 4; i += 32) out[i] = sm[i];
}
int main(int argc, char **argv) {
    const int bw = atoi(argv[1]), bh = atoi(argv[2]), l2 = atoi(argv[3]), sw = atoi(argv[4]), dt = atoi(argv[5]), pitch = atoi(argv[6]);
    const int es = dt == 2 ? 1 : 4;
    const int W = pitch - 3, H = 64;
    std::vector<unsigned char> h((size_t)pitch * H * es);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)(i * 7 + i / 251);
    unsigned char *d; unsigned *out;
    CK(cudaMalloc(&d, h.size()));
    CK(cudaMalloc(&out, 65536));
    CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
    void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m;
    const cuuint64_t d2[2] = {(cuuint64_t)W, (cuuint64_t)H}, s2[1] = {(cuuint64_t)pitch * es};
    const cuuint32_t b2[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, e2[2] = {1, 1};
    const CUtensorMapDataType t = dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    CUresult r = ((Fn)fp)(&m, t, 2, d, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw, (CUtensorMapL2promotion)l2,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d l2 %d sw %d dt %d pitch %d: encode rc %d ", bw, bh, l2, sw, dt, pitch, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    const int bytes = bw * bh * es, x = argc > 7 ? atoi(argv[7]) : 8, y = 5;
    k_2d<<<1, 32, 32768>>>(m, x, y, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("-> %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned char> o(bytes);
    CK(cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost));
    int bad = 0;
    if (sw == 0)
        for (int rr = 0; rr < bh; rr++)
            for (int c = 0; c < bw * es; c++) bad += o[rr * bw * es + c] != h[((size_t)(y + rr) * pitch + x) * es + c];
    printf("-> ran, %d mismatching bytes\n", bad);
    return 0;
}
