// yolopp_mish.cuh — Mish activation, forward and backward, for sm_100a.
//
// Replaces the fork's only in-tree CUDA kernel, mmdet/ops/mish_cuda/src/kernel/mish_cuda.cu:26-71 (scalar, one
// element per thread per iteration, default stream), math from mmdet/ops/mish_cuda/src/mish.h:17-29:
//     fwd:  y = x * tanh(sp),  sp = x < 20 ? log1p(exp(x)) : x
//     bwd:  dx = dy * (x * (1 - tanh(sp)^2) * (1 - exp(-sp)) + tanh(sp))
// Half / BFloat16 compute in float (mish.h:33-50).
//
// Here: with e = exp(x) and n = e * (e + 2) = (1 + e)^2 - 1,
//     tanh(log1p(e)) = n / (n + 2)                                   (one exp, one division — no log1p, no tanh)
//     1 - tanh^2     = 4 (n + 1) / (n + 2)^2 = 4 (1 + e)^2 / (n + 2)^2
//     1 - exp(-sp)   = e / (1 + e)
//     => dx = dy * (4 x e (1 + e) / (n + 2)^2 + n / (n + 2))
// x >= 20 (the reference's softplus threshold) needs no branch: the exponential's argument is clamped there and
// the closed form rounds to the reference's value; x <= -87 underflows to 0 like the reference. 128-bit loads and stores (4 x f32 / 8 x f16 / 8 x bf16 per thread per iteration), grid = a multiple
// of the SM count, on the caller's stream. HBM bound: 8 B/element forward, 12 B/element backward in f32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ypp {

// x is clamped at the reference's threshold for the exponential only: for x >= 20, n = e(e+2) ~ 2e17, n / (n + 2)
// rounds to 1 and y = x — the reference's own branch (mish.h:18: sp = x, tanh(x) == 1 in fp32) without a branch.
__device__ __forceinline__ float mish_fwd_f(float x) {
    const float e = __expf(fminf(x, 20.0f));
    const float n = e * (e + 2.0f);
    return x * __fdividef(n, n + 2.0f);
}

__device__ __forceinline__ float mish_bwd_f(float dy, float x) {
    // x >= 20: e(1+e)/(n+2)^2 ~ 1/e^2 -> 0 and n/(n+2) -> 1: grad = 1, as the reference's branch gives (mish.h:23-28)
    const float e = __expf(fminf(x, 20.0f));
    const float n = e * (e + 2.0f);
    const float inv = __fdividef(1.0f, n + 2.0f);
    const float tsp = n * inv;
    const float g = 4.0f * x * (e * inv) * ((1.0f + e) * inv) + tsp;
    return dy * g;
}

template <typename T>
struct MishVec;
template <>
struct MishVec<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
        f[0] = __uint_as_float(v.x);
        f[1] = __uint_as_float(v.y);
        f[2] = __uint_as_float(v.z);
        f[3] = __uint_as_float(v.w);
    }
    __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    __device__ static __forceinline__ float load1(const float* p) { return *p; }
    __device__ static __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <>
struct MishVec<__half> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ static __forceinline__ float load1(const __half* p) { return __half2float(*p); }
    __device__ static __forceinline__ void store1(__half* p, float v) { *p = __float2half_rn(v); }
};
template <>
struct MishVec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ static __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// streaming loads / stores: every byte is touched once
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
#ifdef MISH_PLAIN
    return *p;
#else
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
#ifdef MISH_PLAIN
    *p = v;
#else
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#endif
}

constexpr int MISH_THREADS = 256;
#ifndef MISH_UNROLL_N
#define MISH_UNROLL_N 4
#endif
constexpr int MISH_UNROLL = MISH_UNROLL_N;  // independent 128-bit loads in flight per thread

template <typename T>
__global__ void __launch_bounds__(MISH_THREADS) mish_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, long long n) {
    constexpr int N = MishVec<T>::N;
    const long long nvec = n / N;
    const long long step = (long long)gridDim.x * MISH_THREADS;
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    uint4* out4 = reinterpret_cast<uint4*>(out);
    for (long long i0 = (long long)blockIdx.x * MISH_THREADS + threadIdx.x; i0 < nvec; i0 += step * MISH_UNROLL) {
        uint4 v[MISH_UNROLL];
#pragma unroll
        for (int u = 0; u < MISH_UNROLL; ++u) {
            const long long i = i0 + u * step;
            if (i < nvec) v[u] = ld_stream(in4 + i);
        }
#pragma unroll
        for (int u = 0; u < MISH_UNROLL; ++u) {
            const long long i = i0 + u * step;
            if (i < nvec) {
                float f[8];
                MishVec<T>::unpack(v[u], f);
#pragma unroll
                for (int k = 0; k < N; ++k) f[k] = mish_fwd_f(f[k]);
                st_stream(out4 + i, MishVec<T>::pack(f));
            }
        }
    }
    // scalar tail (n not a multiple of the vector width)
    const long long t = nvec * N + (long long)blockIdx.x * MISH_THREADS + threadIdx.x;
    if (t < n) MishVec<T>::store1(out + t, mish_fwd_f(MishVec<T>::load1(in + t)));
}

template <typename T>
__global__ void __launch_bounds__(MISH_THREADS) mish_bwd_kernel(const T* __restrict__ grad_out, const T* __restrict__ in,
                                                                T* __restrict__ grad_in, long long n) {
    constexpr int N = MishVec<T>::N;
    const long long nvec = n / N;
    const long long step = (long long)gridDim.x * MISH_THREADS;
    const uint4* g4 = reinterpret_cast<const uint4*>(grad_out);
    const uint4* in4 = reinterpret_cast<const uint4*>(in);
    uint4* out4 = reinterpret_cast<uint4*>(grad_in);
    constexpr int UB = 2;
    for (long long i0 = (long long)blockIdx.x * MISH_THREADS + threadIdx.x; i0 < nvec; i0 += step * UB) {
        uint4 vg[UB], vx[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const long long i = i0 + u * step;
            if (i < nvec) {
                vg[u] = ld_stream(g4 + i);
                vx[u] = ld_stream(in4 + i);
            }
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const long long i = i0 + u * step;
            if (i < nvec) {
                float g[8], x[8];
                MishVec<T>::unpack(vg[u], g);
                MishVec<T>::unpack(vx[u], x);
#pragma unroll
                for (int k = 0; k < N; ++k) g[k] = mish_bwd_f(g[k], x[k]);
                st_stream(out4 + i, MishVec<T>::pack(g));
            }
        }
    }
    const long long t = nvec * N + (long long)blockIdx.x * MISH_THREADS + threadIdx.x;
    if (t < n) MishVec<T>::store1(grad_in + t, mish_bwd_f(MishVec<T>::load1(grad_out + t), MishVec<T>::load1(in + t)));
}

}  // namespace ypp
