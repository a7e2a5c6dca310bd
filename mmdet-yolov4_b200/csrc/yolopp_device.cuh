// yolopp_device.cuh — device-side primitives shared by the yolopp kernels (sm_100a).
//
//  * canonical fp32 arithmetic: every reference operation is ONE IEEE rounding (__fadd_rn/__fmul_rn/... so
//    that nvcc can never contract a*b+c into an FMA), exp() is the canonical polynomial of DESIGN.md
//    §"Canonical arithmetic" (explicit __fmaf_rn; the same sequence as the CPU checker, so bit-reproducible).
//  * order-preserving float<->uint keys, the 64-bit candidate key, the exact IoU predicate of mmcv nms_cpu.
//  * block-level radix select / gather / bitonic sort over 64-bit unique keys.
//  * mbarrier / TMA (cp.async.bulk[.tensor]) wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ypp {

typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// canonical arithmetic
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// exp(x), <= 1.02 ulp, monotone; identical operation sequence on CPU (fmaf) and GPU (__fmaf_rn).
__device__ __forceinline__ float c_expf(float x) {
    float xc = x < -104.0f ? -104.0f : x;  // (NaN compares false twice and is restored at the end: branch-free)
    xc = xc > 89.0f ? 89.0f : xc;
    float t = __fmaf_rn(xc, 1.44269502f, 12582912.0f);
    float j = __fsub_rn(t, 12582912.0f);
    float r = __fmaf_rn(j, -0.693145752f, xc);
    r = __fmaf_rn(j, -1.42860677e-06f, r);
    float p = 1.9875691500e-4f;
    p = __fmaf_rn(p, r, 1.3981999507e-3f);
    p = __fmaf_rn(p, r, 8.3334519073e-3f);
    p = __fmaf_rn(p, r, 4.1665795894e-2f);
    p = __fmaf_rn(p, r, 1.6666665459e-1f);
    p = __fmaf_rn(p, r, 5.0000001201e-1f);
    float r2 = __fmul_rn(r, r);
    float e = __fmaf_rn(p, r2, r);
    e = __fadd_rn(e, 1.0f);
    int ji = __float2int_rz(j);
    int j1 = ji / 2;
    int j2 = ji - j1;
    float s1 = __uint_as_float((uint32_t)(j1 + 127) << 23);
    float s2 = __uint_as_float((uint32_t)(j2 + 127) << 23);
    float y = __fmul_rn(__fmul_rn(e, s1), s2);
    return (x == x) ? y : x;
}

// torch.sigmoid: 1 / (1 + exp(-x)). __frcp_rn is the correctly rounded reciprocal, i.e. bit-identical to the
// IEEE division 1.0f / d, at about half the instructions.
__device__ __forceinline__ float c_sigmoid(float x) {
    float e = c_expf(-x);
    return __frcp_rn(__fadd_rn(1.0f, e));
}

// four sigmoids behind ONE call: the decode kernel's class sweep would otherwise carry a dozen inlined copies of
// the polynomial in its consumer loop (instruction caches: L0 ~6 KB, L1.5 32 KB); four independent dependency
// chains are enough to keep the FMA pipe busy
__device__ __noinline__ float4 c_sigmoid4(float4 x) {
    return make_float4(c_sigmoid(x.x), c_sigmoid(x.y), c_sigmoid(x.z), c_sigmoid(x.w));
}

// ------------------------------------------------------------------------------------------------
// keys
// ------------------------------------------------------------------------------------------------
// order-preserving map float -> uint32 (a < b  <=>  ord(a) < ord(b))
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}

// candidate key: (~ord(score) << 32) | flat candidate index (row * C + class, the reference's own flat index,
// bbox_nms.py:47-49): ascending key order == (score desc, flat index asc).
__device__ __forceinline__ u64 make_key(float score, uint32_t flat) { return ((u64)(~f2ord(score)) << 32) | (u64)flat; }
__device__ __forceinline__ float key_score(u64 k) { return ord2f(~(uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_flat(u64 k) { return (uint32_t)k; }
constexpr uint32_t SCORE_NONE = 0xFFFFFFFFu;  // score-matrix entry of a (row, class) pair that is not a candidate

// ------------------------------------------------------------------------------------------------
// IoU predicate of mmcv nms_cpu:  inter / (area_i + area_j - inter) > thr   (IEEE fp32, division form)
// ------------------------------------------------------------------------------------------------
struct Box {
    float x1, y1, x2, y2, area;
};

__device__ __forceinline__ float box_area(float x1, float y1, float x2, float y2, float foff) {
    return fmul(fadd(fsub(x2, x1), foff), fadd(fsub(y2, y1), foff));
}

// `a` is the earlier (kept) box i, `b` the later box j, exactly as in the reference's inner loop.
__device__ __forceinline__ bool iou_gt(const Box& a, const Box& b, float thr, float foff) {
    float xx1 = a.x1 > b.x1 ? a.x1 : b.x1;
    float yy1 = a.y1 > b.y1 ? a.y1 : b.y1;
    float xx2 = a.x2 < b.x2 ? a.x2 : b.x2;
    float yy2 = a.y2 < b.y2 ? a.y2 : b.y2;
    float w = fadd(fsub(xx2, xx1), foff);
    w = w > 0.f ? w : 0.f;
    float h = fadd(fsub(yy2, yy1), foff);
    h = h > 0.f ? h : 0.f;
    float inter = fmul(w, h);
    float uni = fsub(fadd(a.area, b.area), inter);
    // exact shortcuts around the IEEE division (the division result is monotone in inter/uni, and the
    // guard band of 1e-6 relative is ~8 ulp wide, far more than the one rounding of thr*uni):
    if (uni > 0.f && thr >= 0.f) {
        float t = fmul(thr, uni);
        if (inter < fmul(t, 0.999999f)) return false;
        if (inter > fmul(t, 1.000001f) && t > 1e-30f) return true;
    }
    return fdiv(inter, uni) > thr;
}

// ------------------------------------------------------------------------------------------------
// block-level primitives over unique 64-bit keys
// ------------------------------------------------------------------------------------------------
// Shared scratch used by radix_select / gather (one per block).
struct SelectSmem {
    int hist[256];
    int digit;
    int need;
    int done;
    int count;
};

// Finds T such that exactly `m` eligible keys are <= T. A key is eligible when fetch(i, key) is true and
// (!has_lo || key > lo). Requires: eligible keys unique, m >= 1, m <= #eligible. All threads of the block
// must call this (blockDim.x multiple of 32).
template <class Fetch>
__device__ u64 radix_select(Fetch fetch, int n_slots, bool has_lo, u64 lo, int m, SelectSmem& S) {
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31;
    u64 prefix = 0, pmask = 0;
    int need = m;
#pragma unroll 1
    for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll 1
        for (int i = tid; i < 256; i += nth) S.hist[i] = 0;
        __syncthreads();
#pragma unroll 1
        for (int base = 0; base < n_slots; base += nth) {
            int i = base + tid;
            u64 key = 0;
            bool v = (i < n_slots) && fetch(i, key);
            v = v && (!has_lo || key > lo) && ((key & pmask) == prefix);
            unsigned d = v ? (unsigned)((key >> shift) & 0xFFull) : 256u;
            unsigned peers = __match_any_sync(0xffffffffu, d);
            if (v && lane == (__ffs(peers) - 1)) atomicAdd(&S.hist[d], __popc(peers));
        }
        __syncthreads();
        if (tid < 32) {
            int local[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                local[q] = S.hist[tid * 8 + q];
                s += local[q];
            }
            int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v2 = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += v2;
            }
            int excl = incl - s;
            if (excl < need && need <= incl) {
                int cum = excl;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (cum + local[q] >= need) {
                        S.digit = tid * 8 + q;
                        S.need = need - cum;
                        S.done = (local[q] == need - cum);
                        break;
                    }
                    cum += local[q];
                }
            }
        }
        __syncthreads();
        const int d = S.digit;
        need = S.need;
        const int done = S.done;
        prefix |= (u64)d << shift;
        pmask |= 0xFFull << shift;
        __syncthreads();
        if (done) return prefix | ((shift > 0) ? ((1ull << shift) - 1ull) : 0ull);
    }
    return prefix;
}

// Appends every eligible key <= T to out[0..cap) (unordered); returns the count (same in all threads).
template <class Fetch>
__device__ int gather_le(Fetch fetch, int n_slots, bool has_lo, u64 lo, u64 T, u64* out, int cap, SelectSmem& S) {
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31;
    if (tid == 0) S.count = 0;
    __syncthreads();
#pragma unroll 1
    for (int base = 0; base < n_slots; base += nth) {
        int i = base + tid;
        u64 key = 0;
        bool v = (i < n_slots) && fetch(i, key);
        v = v && (!has_lo || key > lo) && key <= T;
        unsigned bal = __ballot_sync(0xffffffffu, v);
        int slot0 = 0;
        if (lane == 0 && bal) slot0 = atomicAdd(&S.count, __popc(bal));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        if (v) {
            int slot = slot0 + __popc(bal & ((1u << lane) - 1u));
            if (slot < cap) out[slot] = key;
        }
    }
    __syncthreads();
    int c = S.count;
    __syncthreads();
    return c < cap ? c : cap;
}

// In-place ascending bitonic sort of s[0..p2), p2 a power of two. All threads of the block call it.
__device__ __forceinline__ void bitonic_sort(u64* s, int p2) {
    const int tid = threadIdx.x, nth = blockDim.x;
#pragma unroll 1
    for (int size = 2; size <= p2; size <<= 1) {
#pragma unroll 1
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
#pragma unroll 1
            for (int i = tid; i < (p2 >> 1); i += nth) {
                int pos = 2 * i - (i & (stride - 1));
                int j = pos + stride;
                bool up = ((pos & size) == 0);
                u64 a = s[pos], b = s[j];
                if ((a > b) == up) {
                    s[pos] = b;
                    s[j] = a;
                }
            }
        }
    }
    __syncthreads();
}

// Ascending bitonic sort of (key, 16-bit payload) pairs with the data in REGISTERS, one pair per thread: every
// exchange at distance < 32 is a warp shuffle, only the exchanges at distance >= 32 go through shared memory (two
// buffers used alternately: one block barrier per such step). The per-image kernels are bound by instruction issue
// (16 warps on one SM), so what counts is instructions per step: ~12 here against ~80 for a sort that walks shared
// memory. p2 = power of two, 32 <= p2 <= blockDim.x; key[0..p2) / pay[0..p2) in, sorted out; tk / tp = scratch of
// the same size; the caller has synchronised the block after filling key / pay; all threads call.
__device__ __noinline__ void bitonic_sort_kv(u64* __restrict__ key, unsigned short* __restrict__ pay, u64* __restrict__ tk,
                                             unsigned short* __restrict__ tp, int p2) {
    const int e = threadIdx.x, lane = e & 31;
    const bool active = e < p2;  // whole warps
    u64 k = active ? key[e] : ~0ull;
    unsigned v = active ? pay[e] : 0u;
    int flip = 0;
#pragma unroll 1
    for (int size = 2; size <= p2; size <<= 1) {
        const bool up = (e & size) == 0;
        int stride = size >> 1;
#pragma unroll 1
        for (; stride >= 32; stride >>= 1) {
            u64* bk = flip ? key : tk;
            unsigned short* bp = flip ? pay : tp;
            flip ^= 1;
            if (active) {
                bk[e] = k;
                bp[e] = (unsigned short)v;
            }
            __syncthreads();
            if (active) {
                const u64 pk = bk[e ^ stride];
                const unsigned pv = bp[e ^ stride];
                const bool take = (k < pk) != (((e & stride) == 0) == up);  // the lower position keeps the smaller key when ascending
                k = take ? pk : k;
                v = take ? pv : v;
            }
        }
        if (active) {
#pragma unroll 1
            for (; stride > 0; stride >>= 1) {
                const u64 pk = __shfl_xor_sync(0xffffffffu, k, stride);
                const unsigned pv = __shfl_xor_sync(0xffffffffu, v, stride);
                const bool take = (k < pk) != (((lane & stride) == 0) == up);
                k = take ? pk : k;
                v = take ? pv : v;
            }
        }
    }
    __syncthreads();  // (the readers of the last exchange buffer are done)
    if (active) {
        key[e] = k;
        pay[e] = (unsigned short)v;
    }
    __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
#pragma unroll 1
    while (p < v) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------------
// adaptive two-pass selection of the smallest keys (fast path of every top-k on the path)
// ------------------------------------------------------------------------------------------------
constexpr int TS_BINS = 2048;
struct TopSelSmem {
    int hist[TS_BINS];
    int wsum[32];
    int kb, above, bsize, count;
    u64 red[2][32];
    unsigned long long bar;  // mbarrier of the staging copies (select kernel fast path, NMS row staging)
    SelectSmem rs;  // fallback radix select
#ifdef YPP_PROFILE
    int prof_kernel, prof_call;
#endif
};
#ifdef YPP_PROFILE
// stage timestamps of select_sorted_prefix (profiling build only): [kernel][image][call][stage]
__device__ long long g_ssp[2][64][4][10];
__device__ long long g_ssp2[2][64][4][4];  // pass 1 as seen by thread 0: load + test | vote + slots | stash writes | end
#define YPP_SSP(i) do { if (threadIdx.x == 0 && S.prof_call < 4) g_ssp[S.prof_kernel][blockIdx.x + blockIdx.y][S.prof_call][i] = clock64(); } while (0)
#else
#define YPP_SSP(i) do { } while (0)
#endif

// block-wide min / max of 64-bit values (all threads call; result broadcast)
__device__ __forceinline__ void block_minmax(u64 vmin, u64 vmax, u64& omin, u64& omax, TopSelSmem& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        u64 a = __shfl_xor_sync(0xffffffffu, vmin, o), b = __shfl_xor_sync(0xffffffffu, vmax, o);
        vmin = a < vmin ? a : vmin;
        vmax = b > vmax ? b : vmax;
    }
    if (lane == 0) {
        S.red[0][warp] = vmin;
        S.red[1][warp] = vmax;
    }
    __syncthreads();
    if (warp == 0) {
        u64 a = lane < nw ? S.red[0][lane] : ~0ull, b = lane < nw ? S.red[1][lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            u64 a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o);
            a = a2 < a ? a2 : a;
            b = b2 > b ? b2 : b;
        }
        if (lane == 0) {
            S.red[0][0] = a;
            S.red[1][0] = b;
        }
    }
    __syncthreads();
    omin = S.red[0][0];
    omax = S.red[1][0];
    __syncthreads();
}

// Writes into out[0..cnt) — SORTED ascending — a prefix of the ascending order of the eligible keys (eligible:
// delivered by the source and lo_incl <= key <= hi_incl) that holds at least min(m, #eligible) keys; cnt <= cap.
//
// A Source streams its keys in groups through ONE vector load per group and a cheap prefilter, so that the hot
// loop is "load, subtract, compare" and 64-bit keys are only built for the (rare) survivors:
//     typename Source::Raw            raw group (e.g. uint4 of four score words)
//     Raw      load(int g)  const     vector load of group g
//     unsigned mask(Raw, int g) const bit v set: element v of the group MAY be eligible (must not miss any)
//     u64      key(Raw, int v, int g) const   64-bit key of element v
//     int      groups()     const
// Fast path (bucket sort): one histogram pass over the 11 bits below the common prefix of [lo_incl, hi_incl],
// an exclusive scan, one scatter pass (keys land grouped by bucket), and a ranking step inside each bucket.
// Falls back to the exact 8-bit radix select + bitonic sort when the pivot bucket does not fit in `cap`.
// Requires unique keys, cap a power of two >= m, blockDim.x in {256, 512, 1024}; `tmp` is scratch of `cap`
// keys; all threads call. With `prefilled` >= 0 the caller has already placed the (<= cap) eligible keys in
// out[0..prefilled) and the source is not touched (StashSource).
template <class Source>
__device__ __noinline__ int select_sorted_prefix(const Source& src_ref, u64 lo_incl, u64 hi_incl, int m, u64* __restrict__ out,
                                                 u64* __restrict__ tmp, int cap, TopSelSmem& S, int prefilled = -1) {
    typedef typename Source::Raw Raw;
    const Source src = src_ref;  // fields in registers: the caller's object lives in local memory
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nth >> 5;
    const int n_groups = src.groups();
    const u64 x = lo_incl ^ hi_incl;
    int shift = 0;
    if (x) {
        int hb = 63 - __clzll((long long)x);
        shift = hb > 10 ? hb - 10 : 0;
    }
    const u64 base = lo_incl >> shift;
    YPP_SSP(0);
#pragma unroll 1
    for (int i = tid; i < TS_BINS; i += nth) S.hist[i] = 0;
    if (tid == 0) S.count = prefilled < 0 ? 0 : prefilled;  // prefilled: the caller already put the eligible keys in `out`
    __syncthreads();
    YPP_SSP(1);
    constexpr int U = Source::U, V = Source::V;  // U independent vector loads in flight per thread
    static_assert(U * V <= 32 && (V & (V - 1)) == 0, "survivor bitmask");
    // Pass 1 — the only trip to global memory in the common case: stream the source, keep the eligible keys in
    // `out` (the stash). The eligibility test is exact and branch-free (Source::exact), survivors are rare: per
    // batch ONE warp vote, and only when some lane holds a survivor one warp scan + one shared-memory atomic
    // hand out the stash slots; a survivor's key is rebuilt from its (cached) element.
#ifdef YPP_PROFILE
    long long pt[4] = {0, 0, 0, 0}, pt_t = clock64();
#define YPP_ACC(i) do { long long t2 = clock64(); pt[i] += t2 - pt_t; pt_t = t2; } while (0)
#else
#define YPP_ACC(i) do { } while (0)
#endif
#pragma unroll 1
    for (int b0 = 0; b0 < n_groups && prefilled < 0; b0 += nth * U) {
        Raw raw[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const int gi = b0 + q * nth + tid;
            if (gi < n_groups) raw[q] = src.load(gi);
        }
        unsigned emask = 0u;  // bit q*V+v: element v of group q is eligible
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const int gi = b0 + q * nth + tid;
            if (gi < n_groups) emask |= src.exact(raw[q], gi, lo_incl, hi_incl) << (q * V);
        }
        const bool any = __any_sync(0xffffffffu, emask != 0u);
        YPP_ACC(0);
        if (any) {
            const int c = __popc(emask);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            int sp = 0;
            if (lane == 31) sp = atomicAdd(&S.count, incl);
            sp = __shfl_sync(0xffffffffu, sp, 31) + incl - c;
            YPP_ACC(1);
#pragma unroll 1
            while (emask) {
                const int pos = __ffs(emask) - 1;
                emask &= emask - 1;
                if (sp < cap) out[sp] = src.key_at(b0 + (pos / V) * nth + tid, pos % V);
                ++sp;
            }
            __syncwarp();
            YPP_ACC(2);
        }
    }
#ifdef YPP_PROFILE
    if (threadIdx.x == 0 && S.prof_call < 4) {
        long long* g = g_ssp2[S.prof_kernel][blockIdx.x + blockIdx.y][S.prof_call];
        g[0] = pt[0]; g[1] = pt[1]; g[2] = pt[2]; g[3] = clock64();
    }
#endif
    __syncthreads();
    YPP_SSP(2);
    const int n_stash = S.count;
    const bool stash_ok = n_stash <= cap;
    if (stash_ok) {
#pragma unroll 1
        for (int i = tid; i < n_stash; i += nth) atomicAdd(&S.hist[(int)((out[i] >> shift) - base)], 1);
    } else {
        // too many eligible keys for the stash (no tight bound was available): histogram straight from the source
#pragma unroll 1
        for (int b0 = 0; b0 < n_groups; b0 += nth * U) {
            Raw raw[U];
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int gi = b0 + q * nth + tid;
                if (gi < n_groups) raw[q] = src.load(gi);
            }
            unsigned emask = 0u;
#pragma unroll
            for (int q = 0; q < U; ++q) {
                const int gi = b0 + q * nth + tid;
                if (gi < n_groups) emask |= src.exact(raw[q], gi, lo_incl, hi_incl) << (q * V);
            }
#pragma unroll 1
            while (emask) {
                const int pos = __ffs(emask) - 1;
                emask &= emask - 1;
                const u64 k = src.key_at(b0 + (pos / V) * nth + tid, pos % V);
                atomicAdd(&S.hist[(int)((k >> shift) - base)], 1);
            }
        }
    }
    __syncthreads();
    YPP_SSP(3);
    // exclusive scan of the histogram (in place) + pivot bucket: per-thread partial sums -> warp scan -> block scan
    const int per = TS_BINS / nth;  // 2, 4 or 8
    int loc[8], s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        loc[q] = q < per ? S.hist[tid * per + q] : 0;
        s += loc[q];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) S.wsum[warp] = incl;
    if (tid == 0) S.kb = -1;
    __syncthreads();
    // offset of this warp / total: every warp scans the (<= 32) warp sums itself
    int woff, total;
    {
        const int wv = lane < nw ? S.wsum[lane] : 0;
        int wi = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        total = __shfl_sync(0xffffffffu, wi, 31);
        woff = __shfl_sync(0xffffffffu, wi - wv, warp);
    }
    if (m > total) m = total;
    if (total == 0) return 0;
    const int excl = woff + incl - s;
    {
        int cum = excl;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q < per) {
                S.hist[tid * per + q] = cum;  // exclusive offset of the bucket
                if (cum < m && m <= cum + loc[q]) {
                    S.kb = tid * per + q;
                    S.above = cum;
                    S.bsize = loc[q];
                }
                cum += loc[q];
            }
        }
    }
    __syncthreads();
    YPP_SSP(4);
    const int kb = S.kb, above = S.above, bsize = S.bsize;
    int cnt;
    if (above + bsize <= cap) {
        // scatter: hist[bucket] is the running cursor of the bucket, keys land grouped by bucket in tmp
        if (stash_ok) {
            // every survivor of the first pass sits in `out`: no second trip to global memory
#pragma unroll 1
            for (int i = tid; i < n_stash; i += nth) {
                const u64 k = out[i];
                const int bk = (int)((k >> shift) - base);
                if (bk <= kb) tmp[atomicAdd(&S.hist[bk], 1)] = k;
            }
            __syncthreads();  // `out` is rewritten by the ranking step below
        } else {
#pragma unroll 1
            for (int b0 = 0; b0 < n_groups; b0 += nth * U) {
                Raw raw[U];
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    const int gi = b0 + q * nth + tid;
                    if (gi < n_groups) raw[q] = src.load(gi);
                }
                unsigned emask = 0u;
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    const int gi = b0 + q * nth + tid;
                    if (gi < n_groups) emask |= src.exact(raw[q], gi, lo_incl, hi_incl) << (q * V);
                }
#pragma unroll 1
                while (emask) {
                    const int pos = __ffs(emask) - 1;
                    emask &= emask - 1;
                    const u64 k = src.key_at(b0 + (pos / V) * nth + tid, pos % V);
                    const int bk = (int)((k >> shift) - base);
                    if (bk <= kb) tmp[atomicAdd(&S.hist[bk], 1)] = k;
                }
            }
        }
        __syncthreads();
        YPP_SSP(5);
        cnt = above + bsize;
        // rank inside the bucket: after the scatter hist[bk] is the END of bucket bk, hist[bk-1] its start
#pragma unroll 1
        for (int pos = tid; pos < cnt; pos += nth) {
            const u64 key = tmp[pos];
            const int bk = (int)((key >> shift) - base);
            const int st = bk ? S.hist[bk - 1] : 0, en = S.hist[bk];
            int rnk = st;
#pragma unroll 1
            for (int q = st; q < en; ++q) rnk += (tmp[q] < key) ? 1 : 0;
            out[rnk] = key;
        }
        __syncthreads();
        } else {
        // slow, exact path: element-wise view of the source
        const bool has_lo = lo_incl > 0;
        auto f1 = [&](int i, u64& key) -> bool {
            const int gi = i / V, v = i % V;
            const Raw r = src.load(gi);
            if (!((src.exact(r, gi, lo_incl, hi_incl) >> v) & 1u)) return false;
            key = src.key_at(gi, v);
            return true;
        };
        u64 T = radix_select(f1, n_groups * Source::V, has_lo, lo_incl - 1, m, S.rs);
        cnt = gather_le(f1, n_groups * Source::V, has_lo, lo_incl - 1, T, out, cap, S.rs);
        const int p2 = next_pow2(cnt);
#pragma unroll 1
        for (int i = cnt + tid; i < p2; i += nth) out[i] = ~0ull;
        __syncthreads();
        bitonic_sort(out, p2);
    }
    YPP_SSP(6);
#ifdef YPP_PROFILE
    if (threadIdx.x == 0) {
        if (S.prof_call < 4) {
            g_ssp[S.prof_kernel][blockIdx.x + blockIdx.y][S.prof_call][7] = n_stash;
            g_ssp[S.prof_kernel][blockIdx.x + blockIdx.y][S.prof_call][8] = cnt;
            g_ssp[S.prof_kernel][blockIdx.x + blockIdx.y][S.prof_call][9] = n_groups;
        }
        ++S.prof_call;
    }
#endif
    return cnt;
}

// placeholder source of a prefilled stash
struct StashSource {
    typedef uint32_t Raw;
    static constexpr int V = 1, U = 1;
    __device__ __forceinline__ Raw load(int) const { return 0u; }
    __device__ __forceinline__ unsigned exact(const Raw&, int, u64, u64) const { return 0u; }
    __device__ __forceinline__ u64 key_at(int, int) const { return 0ull; }
    __device__ __forceinline__ int groups() const { return 0; }
};

// ------------------------------------------------------------------------------------------------
// mbarrier + TMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// flag word in shared memory read with acquire semantics at CTA scope (pairs with a fence + store on the writer's side)
__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 2-D tiled TMA load: box (c0 = inner coordinate, c1 = row) -> smem, completes on `bar`.
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// Same, with an L2 cache policy (createpolicy): the head tensors are read exactly once -> evict_first keeps the
// small, re-read intermediates (score matrix, rank map, boxes) resident in the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const void* tmap, int c0, int c1, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, "
        "%4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}
// 1-D bulk copy global -> smem (16-byte aligned, size multiple of 16), completes on `bar`.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// 4-byte asynchronous copy global -> shared (LDGSTS): many in flight per thread without holding registers
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// 16-byte variant (both addresses 16-byte aligned), bypassing L1
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

}  // namespace ypp
