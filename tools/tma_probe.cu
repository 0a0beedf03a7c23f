// Standalone probe of the TMA box loads the windowed tracker uses (3-D tensor map over [image][row][column], box smaller
// than the image, arbitrary start column).  usage: tma_probe <variant>   (each variant in its own process)
//   0: map as a direct __grid_constant__ parameter, one lane issues
//   1: maps inside a struct array indexed at run time, one lane issues
//   2: as 1, lanes 0 and 16 issue one box each on their own mbarrier
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cuda/barrier>
using barrier_t = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct alignas(64) Maps { CUtensorMap m[8]; };
constexpr int BW = 16, BH = 14;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void load_box(unsigned dst, const CUtensorMap *map, int x, int y, int z, unsigned mb) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BW * BH * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(mb) : "memory");
}
__device__ __forceinline__ void wait(unsigned mb, unsigned parity) {
    unsigned ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
}

__global__ void k_direct(const __grid_constant__ CUtensorMap map, int x, int y, int z, float *out) {
    extern __shared__ __align__(128) float sm[];
    const unsigned mb = smem_u32(sm + 2 * BW * BH);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) load_box(smem_u32(sm), &map, x, y, z, mb);
    wait(mb, 0);
    for (int i = threadIdx.x; i < BW * BH; i += 32) out[i] = sm[i];
}
__global__ void k_global(const CUtensorMap *map, int x, int y, int z, float *out) {
    extern __shared__ __align__(128) float sm[];
    const unsigned mb = smem_u32(sm + 2 * BW * BH);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) load_box(smem_u32(sm), map, x, y, z, mb);
    wait(mb, 0);
    for (int i = threadIdx.x; i < BW * BH; i += 32) out[i] = sm[i];
}
__global__ void k_2d(const __grid_constant__ CUtensorMap map, int x, int y, float *out) {
    extern __shared__ __align__(128) float sm[];
    const unsigned mb = smem_u32(sm + 2 * BW * BH);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BW * BH * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(sm)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(mb) : "memory");
    }
    wait(mb, 0);
    for (int i = threadIdx.x; i < BW * BH; i += 32) out[i] = sm[i];
}
__global__ void k_libcu(const __grid_constant__ CUtensorMap map, int x, int y, float *out) {
    __shared__ alignas(128) float buf[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier_t bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier_t::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&buf, &map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BW * BH; i += 32) out[i] = (&buf[0][0])[i];
}
__global__ void k_bulk1d(const float *src, float *out) {
    __shared__ alignas(128) float buf[BH * BW];
    __shared__ alignas(8) unsigned long long mbar;
    const unsigned mb = smem_u32(&mbar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(BW * BH * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(buf)), "l"(src), "r"(BW * BH * 4), "r"(mb) : "memory");
    }
    wait(mb, 0);
    for (int i = threadIdx.x; i < BW * BH; i += 32) out[i] = buf[i];
}
__global__ void k_struct(const __grid_constant__ Maps M, int level, int x, int y, int z, int two, float *out) {
    extern __shared__ __align__(128) float sm[];
    const int lane = threadIdx.x, h = lane >> 4, q = lane & 15;
    const unsigned mb = smem_u32(sm + 2 * BW * BH + 2 * h);
    if (q == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (q == 0 && (h == 0 || two)) load_box(smem_u32(sm + h * BW * BH), &M.m[level], x + 3 * h, y + 5 * h, z, mb);
    if (h == 0 || two) wait(mb, 0);
    __syncwarp();
    for (int i = lane; i < (two ? 2 : 1) * BW * BH; i += 32) out[i] = sm[i];
}

int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 480, H = 270, PITCH = 480, B = 2;
    const size_t plane = (size_t)PITCH * H + 64;
    std::vector<float> h(plane * B);
    for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100003);
    float *d, *out;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMalloc(&out, 2 * BW * BH * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    if (!fp || qr != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Maps M;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)PITCH * 4, (cuuint64_t)plane * 4};
    const cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    for (int l = 0; l < 3; l++) {
        CUresult r = ((Fn)fp)(&M.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    }
    const int x = 101, y = 37, z = 1;
    const size_t smem = (2 * BW * BH + 8) * 4;
    if (variant == 5 || variant == 6) {
        CUtensorMap m2;
        const cuuint64_t d2[2] = {(cuuint64_t)W, (cuuint64_t)H}, s2[1] = {(cuuint64_t)PITCH * 4};
        const cuuint32_t b2[2] = {BW, BH}, e2[2] = {1, 1};
        CUresult r = ((Fn)fp)(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d + plane, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const unsigned *w = (const unsigned *)&m2;
        printf("encode rc %d, map:", (int)r);
        for (int i = 0; i < 32; i++) printf(" %08x", w[i]);
        printf("\n");
        if (variant == 5) k_libcu<<<1, 32>>>(m2, x, y, out);
        else {
            k_bulk1d<<<1, 32>>>(d, out);
            CK(cudaDeviceSynchronize());
            std::vector<float> o(BW * BH);
            CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int i = 0; i < BW * BH; i++) bad += o[i] != h[i];
            printf("variant 6 (1-D bulk copy): %s\n", bad ? "WRONG" : "ok");
            return bad != 0;
        }
    } else if (variant == 3) {
        CUtensorMap *dm;
        CK(cudaMalloc(&dm, sizeof(CUtensorMap)));
        CK(cudaMemcpy(dm, &M.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        k_global<<<1, 32, smem>>>(dm, x, y, z, out);
    } else if (variant == 4) {
        CUtensorMap m2;
        const cuuint64_t d2[2] = {(cuuint64_t)W, (cuuint64_t)H}, s2[1] = {(cuuint64_t)PITCH * 4};
        const cuuint32_t b2[2] = {BW, BH}, e2[2] = {1, 1};
        CUresult r = ((Fn)fp)(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d + plane, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode2 failed %d\n", (int)r); return 1; }
        k_2d<<<1, 32, smem>>>(m2, x, y, out);
    } else if (variant == 0) k_direct<<<1, 32, smem>>>(M.m[0], x, y, z, out);
    else k_struct<<<1, 32, smem>>>(M, 2, x, y, z, variant == 2, out);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(2 * BW * BH);
    CK(cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int hh = 0; hh < (variant == 2 ? 2 : 1); hh++)
        for (int r = 0; r < BH; r++)
            for (int c = 0; c < BW; c++)
                bad += o[hh * BW * BH + r * BW + c] != h[(size_t)z * plane + (size_t)(y + 5 * hh + r) * PITCH + x + 3 * hh + c];
    printf("variant %d: %s (%d mismatches)\n", variant, bad ? "WRONG" : "ok", bad);
    return bad != 0;
}
