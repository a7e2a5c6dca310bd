// Probe: does a TMA tiled load accept a start coordinate that is NOT 16-byte aligned in the innermost dimension?
// rank-1 tensor map over a float array, box = 32 floats (128 B), SWIZZLE_128B and SWIZZLE_NONE, start coordinates 0..7,
// 361, 362. Prints the first mismatch (or ok) per case; a fault shows up as a CUDA error.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_unaligned_probe tools/tma_unaligned_probe.cu && tools/tma_unaligned_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int coord, float* out, int rank) {
    __shared__ __align__(1024) float buf[8 * 32];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(8 * 128) : "memory");
        // eight boxes into eight consecutive 128-byte rows (row k gets the box starting at coord + k * 64)
        for (int k = 0; k < 8; ++k)
            if (rank == 1)
                asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];" ::"r"(
                                 smem_u32(buf + k * 32)),
                             "l"(&map), "r"(smem_u32(&bar)), "r"(coord + k * 64)
                             : "memory");
            else  // 2-D map (row length 368): the same element through (column, row)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                                 smem_u32(buf + k * 32)),
                             "l"(&map), "r"(smem_u32(&bar)), "r"((coord + k * 64) % 368), "r"((coord + k * 64) / 368)
                             : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int N = 1 << 16;
    std::vector<float> h(N);
    for (int i = 0; i < N; ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, N * 4);
    cudaMalloc(&o, 8 * 32 * 4);
    cudaMemcpy(d, h.data(), N * 4, cudaMemcpyHostToDevice);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    for (int cfg = 0; cfg < 4; ++cfg) {
        const int sw = cfg & 1, rank = cfg < 2 ? 2 : 1;
        CUtensorMap map;
        cuuint64_t gdim[2] = {rank == 2 ? 368ull : (cuuint64_t)N, (cuuint64_t)(N / 368)};
        cuuint64_t gstr[1] = {368 * 4};
        cuuint32_t box[2] = {32, 1};
        cuuint32_t estr[2] = {1, 1};
        printf("rank %d ", rank);
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("swizzle %d: encode rc %d\n", sw, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int coords[] = {0, 4, 1, 2, 3, 5, 7, 361, 362};
        for (int ci = 0; ci < 9; ++ci) {
            cudaMemset(o, 0xff, 8 * 32 * 4);
            probe<<<1, 64>>>(map, coords[ci], o, rank);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("  coord %d: CUDA error %s\n", coords[ci], cudaGetErrorString(e));
                return 1;  // (the context is gone)
            }
            float res[8 * 32];
            cudaMemcpy(res, o, sizeof(res), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int k = 0; k < 8 && !bad; ++k)
                for (int c = 0; c < 32; ++c) {
                    // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
                    const int chunk = c >> 2, within = c & 3;
                    const int pc = sw ? (((chunk ^ (k & 7)) << 2) | within) : c;
                    const float want = (float)(coords[ci] + k * 64 + c);
                    if (res[k * 32 + pc] != want) {
                        printf("  coord %d: MISMATCH row %d col %d got %.0f want %.0f\n", coords[ci], k, c, res[k * 32 + pc], want);
                        bad = 1;
                        break;
                    }
                }
            if (!bad) printf("  coord %d: ok\n", coords[ci]);
        }
    }
    return 0;
}
