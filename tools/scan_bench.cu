// Microbenchmark: how fast can ONE CTA per image scan a 320 KB slice (80k u32) that sits in L2?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/scan_bench tools/scan_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int U>
__global__ void scan_kernel(const uint4* __restrict__ mat, int groups_per_img, uint32_t lo, uint32_t span, int* out, long long* cyc) {
    const uint4* m = mat + (size_t)blockIdx.x * groups_per_img;
    int cnt = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < 2; ++rep)
    for (int b0 = 0; b0 < groups_per_img; b0 += blockDim.x * U) {
        uint4 v[U];
#pragma unroll
        for (int q = 0; q < U; ++q) { int gi = b0 + q * blockDim.x + threadIdx.x; v[q] = gi < groups_per_img ? __ldcg(m + gi) : make_uint4(~0u, ~0u, ~0u, ~0u); }
#pragma unroll
        for (int q = 0; q < U; ++q) { cnt += (v[q].x - lo <= span) + (v[q].y - lo <= span) + (v[q].z - lo <= span) + (v[q].w - lo <= span); }
    }
    long long t1 = clock64();
    if (cnt == 123456789) out[0] = cnt;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// gather variant: a warp reads one 320-byte row (20 lanes x 16 B) from a list of row indices held in shared memory
__global__ void gather_kernel(const uint4* __restrict__ mat, int groups_per_img, const int* __restrict__ rowlist, int nrows, uint32_t lo, uint32_t span, int* out, long long* cyc) {
    __shared__ unsigned long long rows[1024];
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) rows[i] = rowlist[blockIdx.x * 1024 + i];
    __syncthreads();
    const uint32_t* m = (const uint32_t*)(mat + (size_t)blockIdx.x * groups_per_img);
    int cnt = 0;
    long long t0 = clock64();
    const int ng = nrows * 32;
    for (int b0 = 0; b0 < ng; b0 += blockDim.x * 4) {
        uint4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int g = b0 + q * blockDim.x + threadIdx.x;
            v[q] = make_uint4(~0u, ~0u, ~0u, ~0u);
            if (g < ng) { int i = g >> 5, j = g & 31; if (j < 20) v[q] = __ldcg((const uint4*)(m + (uint32_t)rows[i] * 80u + 4 * j)); }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { cnt += (v[q].x - lo <= span) + (v[q].y - lo <= span) + (v[q].z - lo <= span) + (v[q].w - lo <= span); }
    }
    long long t1 = clock64();
    if (cnt == 123456789) out[0] = cnt;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    const int B = 64, slots = 80000; size_t n = (size_t)B * slots;
    uint32_t* d; cudaMalloc(&d, n * 4); cudaMemset(d, 0x3c, n * 4);
    int* out; cudaMalloc(&out, 4); long long* cyc; cudaMalloc(&cyc, 8 * B); long long h[64];
    for (int threads : {256, 512, 1024}) {
        for (int pass = 0; pass < 2; ++pass) {
            scan_kernel<4><<<B, threads>>>((uint4*)d, slots / 4, 0x3f000000u, 0x00800000u, out, cyc); cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 8 * B, cudaMemcpyDeviceToHost);
            if (pass) printf("U=4 threads=%4d: %lld cycles for 2 scans of 320KB (%s)\n", threads, h[0], cudaGetErrorString(cudaGetLastError()));
            scan_kernel<8><<<B, threads>>>((uint4*)d, slots / 4, 0x3f000000u, 0x00800000u, out, cyc); cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, 8 * B, cudaMemcpyDeviceToHost);
            if (pass) printf("U=8 threads=%4d: %lld cycles\n", threads, h[0]);
        }
    }
    // gather of 441 random rows per image
    int* hl = new int[64 * 1024]; srand(1);
    for (int b = 0; b < 64; ++b) for (int i = 0; i < 1024; ++i) hl[b * 1024 + i] = rand() % 1000;
    int* dl; cudaMalloc(&dl, 64 * 1024 * 4); cudaMemcpy(dl, hl, 64 * 1024 * 4, cudaMemcpyHostToDevice);
    for (int pass = 0; pass < 3; ++pass) {
        gather_kernel<<<B, 512>>>((uint4*)d, slots / 4, dl, 441, 0x3f000000u, 0x00800000u, out, cyc); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8 * B, cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1LL << 60; for (int b = 0; b < B; ++b) { if (h[b] > mx) mx = h[b]; if (h[b] < mn) mn = h[b]; }
        printf("gather 441 rows, 512 threads: min %lld max %lld cycles (%s)\n", mn, mx, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
