// Probe: which TMA reduction / load forms work on this GPU (sm_100a) for a Float32 tensor with 24 x 11 x 1 boxes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_tensor_red(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int mode)
{
    extern __shared__ __align__(128) float s[];
    __shared__ __align__(8) unsigned long long bar;
    for (int i = threadIdx.x; i < 24 * 11; i += 32) s[i] = 1.f + 0.001f * i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    if (mode == 0) {            // tensor reduce add
        if (threadIdx.x == 0) {
            asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(sa) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (mode == 1) {     // tensor store (no reduction)
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(sa) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (mode == 2) {     // tensor load with mbarrier, then write the tile back with plain stores for checking
        const unsigned ba = (unsigned)__cvta_generic_to_shared(&bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(1056) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(sa), "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(ba) : "memory");
        }
        asm volatile("{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W1;\n\t}" ::"r"(ba), "r"(0) : "memory");
    }
}

__global__ void k_bulk_red(float *g)
{
    extern __shared__ __align__(128) float s[];
    for (int i = threadIdx.x; i < 24; i += 32) s[i] = 2.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(g), "r"(sa), "r"(96) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void k_read_tile(float *out)      // mode 2 helper: nothing (the load test only checks completion)
{
    out[0] = 1.f;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int Nx = 64, Ny = 32, Nz = 16;
    float *g;
    CK(cudaMalloc(&g, (size_t)2 * Nx * Ny * Nz * 4));
    CK(cudaMemset(g, 0, (size_t)2 * Nx * Ny * Nz * 4));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    const cuuint64_t dims[3] = {2 * Nx, Ny, Nz};
    const cuuint64_t strides[2] = {2 * Nx * 4, (cuuint64_t)2 * Nx * 4 * Ny};
    const cuuint32_t box[3] = {24, 11, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    std::vector<float> h((size_t)2 * Nx * Ny * Nz);
    const char *names[3] = {"tensor reduce add f32", "tensor store", "tensor load"};
    for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(g, 0, h.size() * 4));
        k_tensor_red<<<1, 32, 2048>>>(tm, 8, 3, 2, mode);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%s: %s", names[mode], cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("\n"); return 1; }
        if (mode < 2) {
            // twice for the reduction: expect 2x
            if (mode == 0) { k_tensor_red<<<1, 32, 2048>>>(tm, 8, 3, 2, mode); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(h.data(), g, h.size() * 4, cudaMemcpyDeviceToHost));
            double sum = 0; for (float v : h) sum += v;
            const float at = h[((size_t)2 * Ny + 3) * 2 * Nx + 8];
            printf("  sum %.3f  first element %.4f (expect %s)\n", sum, at, mode == 0 ? "2.0" : "1.0");
        } else printf("\n");
    }
    // 1-D bulk reduction first (independent of the tensor map)
    CK(cudaMemset(g, 0, h.size() * 4));
    k_bulk_red<<<1, 32, 2048>>>(g + 8);
    printf("bulk (1-D) reduce add f32: %s", cudaGetErrorString(cudaDeviceSynchronize()));
    CK(cudaMemcpy(h.data(), g, h.size() * 4, cudaMemcpyDeviceToHost));
    { double sum = 0; for (float v : h) sum += v; printf("  sum %.3f (expect 48)\n", sum); }
    // boxes sticking out of the tensor: which directions are legal for reductions / loads?
    const int cases[6][3] = {{104, 3, 2}, {8, 25, 2}, {8, 3, 15}, {-8, 3, 2}, {8, -3, 2}, {120, 28, 2}};
    for (int mode = 0; mode <= 2; mode += 2)
        for (int i = 0; i < 6; ++i) {
            cudaMemset(g, 0, h.size() * 4);
            k_tensor_red<<<1, 32, 2048>>>(tm, cases[i][0], cases[i][1], cases[i][2], mode);
            cudaError_t e = cudaDeviceSynchronize();
            printf("%s at (%d, %d, %d): %s", names[mode], cases[i][0], cases[i][1], cases[i][2], cudaGetErrorString(e));
            if (e != cudaSuccess) { printf("\n"); return 1; }
            cudaMemcpy(h.data(), g, h.size() * 4, cudaMemcpyDeviceToHost);
            double sum = 0; for (float v : h) sum += v;
            printf("  sum %.3f\n", sum);
        }
    return 0;
}
