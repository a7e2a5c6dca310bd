// Microbenchmark: how fast can (NA x TW) tiles of a (B*A*NA, HW) fp32 matrix be streamed into shared memory with
// TMA, as a function of the tile width TW, boxes per tile, stages and CTAs/SM? Consumers only wait + release.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_stream_bench tools/tma_stream_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../mmdet-yolov4_b200/csrc/yolopp_device.cuh"
using namespace ypp;

struct Cfg { int NA, HW, planes, TW, nbox, stages, hint, swz; int tpp, tiles; uint32_t box_bytes, sub_bytes, stage_bytes; };

__global__ void __launch_bounds__(64) stream_kernel(const __grid_constant__ CUtensorMap map, Cfg c, float* sink) {
    extern __shared__ unsigned char dyn[];
    unsigned char* base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    uint64_t* full = (uint64_t*)base; uint64_t* empty = full + 16; unsigned char* st = base + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int s = 0; s < c.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1);} fence_mbar_init(); }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            uint64_t pol = l2_policy_evict_first();
            int it = 0;
            for (int t = blockIdx.x; t < c.tiles; t += gridDim.x, ++it) {
                int s = it % c.stages; uint32_t ph = (it / c.stages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                int plane = t / c.tpp, ht = t - plane * c.tpp;
                unsigned char* dst = st + (size_t)s * c.stage_bytes;
                int hw0 = ht * c.TW * c.nbox;
                int nb = 0; for (int q = 0; q < c.nbox; ++q) if (hw0 + q * c.TW < c.HW) ++nb;
                mbar_arrive_expect_tx(&full[s], c.box_bytes * nb);
                for (int q = 0; q < nb; ++q) {
                    if (c.hint) tma_load_2d_hint(dst + q * c.sub_bytes, &map, hw0 + q * c.TW, plane * c.NA, &full[s], pol);
                    else tma_load_2d(dst + q * c.sub_bytes, &map, hw0 + q * c.TW, plane * c.NA, &full[s]);
                }
            }
        }
    } else {
        int it = 0; float acc = 0.f;
        for (int t = blockIdx.x; t < c.tiles; t += gridDim.x, ++it) {
            int s = it % c.stages; uint32_t ph = (it / c.stages) & 1;
            mbar_wait(&full[s], ph);
            acc += *(const float*)(st + (size_t)s * c.stage_bytes + lane * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 12345.678f) sink[0] = acc;
    }
}

// plain vectorised streaming read of the same bytes (reference ceiling)
__global__ void ldg_kernel(const float4* __restrict__ p, size_t n4, float* sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (; i + 3 * step < n4; i += 4 * step) {
        float4 a = __ldcs(p + i), b = __ldcs(p + i + step), c = __ldcs(p + i + 2 * step), d = __ldcs(p + i + 3 * step);
        acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n4; i += step) acc += __ldcs(p + i).x;
    if (acc == 12345.678f) sink[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int NA = 85, HW = 5776, planes = 64 * 3;
    size_t n = (size_t)planes * NA * HW;
    float* d; cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4);
    float* sink; cudaMalloc(&sink, 4);
    void* sym; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double bytes = (double)n * 4;
    // reference: LDG stream
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        ldg_kernel<<<sms * 8, 512>>>((const float4*)d, n / 4, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("ldg float4 stream: %.1f us  %.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    struct T { int TW, nbox, stages, cps, hint, swz; };
    std::vector<T> tests = {
        {60, 1, 4, 2, 0, 0}, {64, 1, 4, 2, 0, 0}, {32, 2, 4, 2, 1, 1}, {32, 2, 4, 2, 0, 1}, {32, 1, 8, 2, 0, 1},
        {128, 1, 4, 1, 0, 0}, {128, 1, 2, 2, 0, 0}, {256, 1, 2, 1, 0, 0}, {32, 4, 4, 1, 0, 1}, {32, 8, 2, 1, 0, 1},
        {64, 1, 8, 1, 0, 0}, {64, 1, 4, 2, 1, 0}, {128, 1, 4, 1, 1, 0}, {256, 1, 2, 1, 1, 0},
    };
    for (auto& t : tests) {
        Cfg c; c.NA = NA; c.HW = HW; c.planes = planes; c.TW = t.TW; c.nbox = t.nbox; c.stages = t.stages; c.hint = t.hint; c.swz = t.swz;
        c.tpp = (HW + t.TW * t.nbox - 1) / (t.TW * t.nbox); c.tiles = c.tpp * planes;
        c.box_bytes = NA * t.TW * 4; c.sub_bytes = (c.box_bytes + 1023) & ~1023u; c.stage_bytes = c.sub_bytes * t.nbox;
        size_t smem = 2048 + (size_t)c.stage_bytes * t.stages;
        if (smem > 227 * 1024 / t.cps) { printf("TW=%d nbox=%d stages=%d cps=%d: smem %zu too big\n", t.TW, t.nbox, t.stages, t.cps, smem); continue; }
        CUtensorMap map;
        cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)planes * NA}; cuuint64_t gstr[1] = {(cuuint64_t)HW * 4};
        cuuint32_t box[2] = {(cuuint32_t)t.TW, (cuuint32_t)NA}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         t.swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            stream_kernel<<<sms * t.cps, 64, smem>>>(map, c, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("TW=%3d nbox=%d stages=%d ctas/sm=%d hint=%d swz=%d smem=%6zu: %.1f us  %.0f GB/s  (%s)\n", t.TW, t.nbox, t.stages, t.cps, t.hint, t.swz, smem,
               best * 1e3, bytes / best / 1e6, cudaGetErrorString(err));
    }
    return 0;
}
