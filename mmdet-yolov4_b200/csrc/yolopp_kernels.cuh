// yolopp_kernels.cuh — the kernels of the post-processing path (sm_100a). Included by yolopp_capi.cu.
//
//   K0 select_kernel      objectness top-k per (image, segment): adaptive 2-pass select + smem sort -> rank map
//   K1 decode_tma_kernel  TMA-streamed fused decode + score + threshold -> dense (row, class) score matrix
//      decode_ldg_kernel  same semantics with plain coalesced loads (HW % 4 != 0 levels, dense admission)
//   K2 nms_image_kernel   per image: candidates visited in global (score desc, flat index asc) order, fetched in
//                         sorted chunks by the adaptive select; greedy NMS against a kept list with the
//                         reference's regime rule (n < split_thr: one problem over class-offset boxes;
//                         else: classes independent); stops at max_per_img kept and writes the outputs.
#pragma once
#include "yolopp_device.cuh"

namespace ypp {

constexpr int MAXL = 8;
constexpr int MAXA = 8;
constexpr uint32_t RANK_INVALID = 0xFFFFFFFFu;

constexpr int TILE_T = 60;       // positions per TMA tile: 240-byte rows (16-B multiple), row pitch = 4*odd banks
constexpr int DEC_STAGES = 4;    // TMA pipeline depth == consumer warps: every consumer warp owns one stage
constexpr int DEC_CWARPS = 4;    // consumer warps per CTA (+1 producer warp), one whole tile per warp
constexpr int DEC_THREADS = 32 * (1 + DEC_CWARPS);

constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAX_K = 4096;  // largest nms_pre the select kernel sorts in shared memory

constexpr int NMS_THREADS = 1024;
constexpr int NMS_KCAP = 2048;   // sorted-chunk buffer (keys)
constexpr int NMS_CH = 1024;     // candidates staged (boxes) per chunk
constexpr int NMS_MAX_KEEP = 4096;

struct LevelDev {
    const float* ptr;
    int H, W, HW;
    int n_off;    // offset of this level in the concatenated anchor index n = n_off + (y*W+x)*A + a
    int m_off;    // offset in the plane-major, 4-element aligned index m = m_off + a*HW + hw
    int seg;      // top-k segment this level belongs to
    int use_tma;  // decoded by decode_tma_kernel (else decode_ldg_kernel)
    int tile0;    // first tile id of this level in its kernel's tile enumeration
    int tpp;      // tiles per plane
    float sx, sy; // anchor-generator strides
    float cstride;// coder stride
    float base[MAXA][4];
};

struct SegDev {
    int first_level, num_levels;
    int N;        // anchors in the segment
    int k;        // rows kept (k == N when no top-k runs)
    int row_off;  // first row of the segment
    int has_topk;
    int m_begin, m_end;  // plane-major index range of the segment
};

struct DevParams {
    int mode, B, L, A, C, NA, agnostic;  // C = effective classes (1 when class agnostic)
    int nsegs, ntopk;
    int topk_segs[MAXL];
    int N, M_pad, R;
    float score_thr, conf_thr, iou_thr, foff;
    int split_thr, nms_agnostic, m_eff, keep_cap, out_cap, rescale;
    int sel_kcap;  // key buffer (power of two) of the select kernel
    int tma_tiles, ldg_blocks;
    LevelDev lv[MAXL];
    SegDev seg[MAXL];
    // workspace
    u64* ckey;           // [B][M_pad]  (~ord(conf) << 32 | n), ~0 in the alignment padding
    uint32_t* rank;      // [B][M_pad]  row of each anchor (plane-major), RANK_INVALID when not admitted
    int* row_anchor;     // [B][R]
    float4* row_box;     // [B][R]
    uint32_t* mat;       // [B][R][C]   score bits of candidate (row, class), SCORE_NONE otherwise
    uint32_t* img_max;   // [B]  f2ord(max coordinate over candidate boxes)
    const float* scale;  // [B][4] or null
    // outputs
    float* o_dets;
    long long* o_labels;
    int* o_anchors;
    int* o_rows;
    int* o_count;
    int* o_ncand;
    int* o_status;
};

// ------------------------------------------------------------------------------------------------
// K0: objectness top-k
// ------------------------------------------------------------------------------------------------
// One CTA per (top-k segment, image). Keys are (~ord(conf) << 32 | n): ascending = (conf desc, anchor asc) —
// the canonical order of conf_pred.topk(nms_pre) (yolocsp_head.py:350-355 / yolo_head.py:281-302).
__global__ void __launch_bounds__(SEL_THREADS) select_kernel(const __grid_constant__ DevParams P) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    __shared__ TopSelSmem S;
    u64* sel = reinterpret_cast<u64*>(sel_smem);
    const int b = blockIdx.y;
    const SegDev& sg = P.seg[P.topk_segs[blockIdx.x]];
    const int tid = threadIdx.x;
    u64* ckey = P.ckey + (size_t)b * P.M_pad;
    uint32_t* rank = P.rank + (size_t)b * P.M_pad;

    // pass 0: objectness of every anchor of the segment -> composite key; rank map cleared; key range
    u64 kmin = ~0ull, kmax = 0ull;
    for (int li = 0; li < sg.num_levels; ++li) {
        const LevelDev& lv = P.lv[sg.first_level + li];
        for (int a = 0; a < P.A; ++a) {
            const float* plane = lv.ptr + ((size_t)(b * P.A + a) * P.NA + 4) * lv.HW;
            for (int hw = tid; hw < lv.HW; hw += SEL_THREADS) {
                float conf = c_sigmoid(__ldg(plane + hw));
                int m = lv.m_off + a * lv.HW + hw;
                u64 key = ((u64)(~f2ord(conf)) << 32) | (u64)(uint32_t)(lv.n_off + hw * P.A + a);
                ckey[m] = key;
                rank[m] = RANK_INVALID;
                kmin = key < kmin ? key : kmin;
                kmax = key > kmax ? key : kmax;
            }
        }
        // alignment padding between levels never matches a key
        const int pad0 = lv.m_off + P.A * lv.HW, pad1 = (pad0 + 3) & ~3;
        if (tid < pad1 - pad0) ckey[pad0 + tid] = ~0ull;
    }
    u64 gmin, gmax;
    block_minmax(kmin, kmax, gmin, gmax, S);  // also orders the global writes above (block barrier)

    const u64* kp = ckey + sg.m_begin;
    auto fetch = [=](int i, u64& key) -> bool {
        key = kp[i];
        return key != ~0ull;
    };
    const int cnt = select_sorted_prefix(fetch, sg.m_end - sg.m_begin, gmin, gmax, sg.k, sel, P.sel_kcap, S);
    const int k = cnt < sg.k ? cnt : sg.k;
    const int first = sg.first_level, nl = sg.num_levels, A = P.A;
    for (int i = tid; i < k; i += SEL_THREADS) {
        const int n = (int)(uint32_t)sel[i];
        int l = first;
        for (int q = nl - 1; q >= 0; --q)
            if (n >= P.lv[first + q].n_off) {
                l = first + q;
                break;
            }
        const LevelDev& lv = P.lv[l];
        const int loc = n - lv.n_off;
        const int hw = loc / A, a = loc - hw * A;
        rank[lv.m_off + a * lv.HW + hw] = (uint32_t)(sg.row_off + i);
        P.row_anchor[(size_t)b * P.R + sg.row_off + i] = n;
    }
}

// ------------------------------------------------------------------------------------------------
// K1: decode
// ------------------------------------------------------------------------------------------------
// Box of one anchor, in the reference's exact operation order.
//   CSP: yolocsp_head.py:274-275 + YOLOV4BBoxCoder.decode yolov4_bbox_coder.py:52-65
//   V3 : yolo_head.py:267-274    + YOLOBBoxCoder.decode  yolo_bbox_coder.py:74-87
// act[0..3]: CSP sigmoid(t0..t3); V3 sigmoid(t0), sigmoid(t1), exp(t2), exp(t3).
template <int MODE>
__device__ __forceinline__ float4 decode_box(const LevelDev& lv, int a, int x, int y, float a0, float a1, float a2,
                                             float a3) {
    // AnchorGenerator.single_level_grid_anchors: base + shift (anchor_generator.py:256-266)
    float shx = fmul((float)x, lv.sx), shy = fmul((float)y, lv.sy);  // exact: integers < 2^24
    float bx1 = fadd(lv.base[a][0], shx), by1 = fadd(lv.base[a][1], shy);
    float bx2 = fadd(lv.base[a][2], shx), by2 = fadd(lv.base[a][3], shy);
    float xc = fmul(fadd(bx1, bx2), 0.5f), yc = fmul(fadd(by1, by2), 0.5f);
    float w = fsub(bx2, bx1), h = fsub(by2, by1);
    float xcp, ycp, wp, hp;
    if (MODE == 0) {
        float px = fsub(fmul(a0, 2.0f), 1.0f), py = fsub(fmul(a1, 2.0f), 1.0f);
        float w2 = fmul(a2, 2.0f), h2 = fmul(a3, 2.0f);
        float pw = fmul(w2, w2), ph = fmul(h2, h2);
        xcp = fadd(fmul(px, lv.cstride), xc);
        ycp = fadd(fmul(py, lv.cstride), yc);
        wp = fmul(pw, w);
        hp = fmul(ph, h);
    } else {
        xcp = fadd(fmul(fsub(a0, 0.5f), lv.cstride), xc);
        ycp = fadd(fmul(fsub(a1, 0.5f), lv.cstride), yc);
        wp = fmul(a2, w);
        hp = fmul(a3, h);
    }
    float hw_ = fmul(wp, 0.5f), hh_ = fmul(hp, 0.5f);  // "/ 2" (exact)
    return make_float4(fsub(xcp, hw_), fsub(ycp, hh_), fadd(xcp, hw_), fadd(ycp, hh_));
}

__device__ __forceinline__ float4 rescale_box(float4 bx, const float* sc) {
    return make_float4(fdiv(bx.x, sc[0]), fdiv(bx.y, sc[1]), fdiv(bx.z, sc[2]), fdiv(bx.w, sc[3]));
}

__device__ __forceinline__ float box_max(float4 bx) {
    float m = bx.x > bx.y ? bx.x : bx.y;
    m = bx.z > m ? bx.z : m;
    m = bx.w > m ? bx.w : m;
    return m;
}

// row index of anchor (level, a, hw): from the rank map when the segment ran a top-k, arithmetic otherwise
__device__ __forceinline__ uint32_t row_of(const DevParams& P, const LevelDev& lv, int b, int a, int hw) {
    const SegDev& sg = P.seg[lv.seg];
    if (sg.has_topk) return P.rank[(size_t)b * P.M_pad + lv.m_off + a * lv.HW + hw];
    return (uint32_t)(sg.row_off + (lv.n_off - P.lv[sg.first_level].n_off) + hw * P.A + a);
}

struct TmapPack {
    CUtensorMap m[MAXL];
};

// Persistent, warp-specialised: warp 0 streams (NA x TILE_T) tiles of the raw head tensor into a 4-stage smem
// ring with TMA (+ the tile's rank-map row with a 1-D bulk copy); each of the 4 consumer warps owns one stage
// and processes whole tiles: admitted anchors one after the other, lanes over classes, scores written with
// coalesced stores into the (row, class) matrix. No atomics with a return value anywhere.
template <int MODE>
__global__ void __launch_bounds__(DEC_THREADS) decode_tma_kernel(const __grid_constant__ DevParams P,
                                                                  const __grid_constant__ TmapPack maps) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NA = P.NA;
    const uint32_t tile_bytes = (uint32_t)NA * TILE_T * 4u;
    const uint32_t rank_bytes = TILE_T * 4u;
    const uint32_t stage_bytes = (tile_bytes + rank_bytes + 127u) & ~127u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty = full + DEC_STAGES;
    unsigned char* stages = smem_raw + 128;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEC_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int total = P.tma_tiles;
    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
                const int s = it % DEC_STAGES;
                const uint32_t ph = (uint32_t)(it / DEC_STAGES) & 1u;
                mbar_wait(&empty[s], ph ^ 1u);
                int l = 0;
                for (int q = 0; q < P.L; ++q)
                    if (P.lv[q].use_tma && t >= P.lv[q].tile0) l = q;
                const LevelDev& lv = P.lv[l];
                const int loc = t - lv.tile0;
                const int plane = loc / lv.tpp;  // b*A + a
                const int ht = loc - plane * lv.tpp;
                const int bb = plane / P.A, a = plane - bb * P.A;
                unsigned char* dst = stages + (size_t)s * stage_bytes;
                const bool topk = P.seg[lv.seg].has_topk != 0;
                mbar_arrive_expect_tx(&full[s], tile_bytes + (topk ? rank_bytes : 0u));
                tma_load_2d(dst, &maps.m[l], ht * TILE_T, plane * NA, &full[s]);
                if (topk)
                    bulk_load_1d(dst + tile_bytes, P.rank + (size_t)bb * P.M_pad + lv.m_off + a * lv.HW + ht * TILE_T,
                                 rank_bytes, &full[s]);
            }
        }
        return;
    }
    // ---------------- consumers: warp cw takes iterations it == cw (mod DEC_CWARPS), stage == cw ----------------
    const int cw = warp - 1;
    for (int it = cw, t = blockIdx.x + cw * gridDim.x; t < total; it += DEC_CWARPS, t += DEC_CWARPS * gridDim.x) {
        const int s = it % DEC_STAGES;
        const uint32_t ph = (uint32_t)(it / DEC_STAGES) & 1u;
        int l = 0;
        for (int q = 0; q < P.L; ++q)
            if (P.lv[q].use_tma && t >= P.lv[q].tile0) l = q;
        const LevelDev& lv = P.lv[l];
        const int loc = t - lv.tile0;
        const int plane = loc / lv.tpp;
        const int ht = loc - plane * lv.tpp;
        const int b = plane / P.A, a = plane - b * P.A;
        const int hw0 = ht * TILE_T;
        const SegDev& sg = P.seg[lv.seg];

        mbar_wait(&full[s], ph);
        const float* tile = reinterpret_cast<const float*>(stages + (size_t)s * stage_bytes);
        const uint32_t* rk = reinterpret_cast<const uint32_t*>(stages + (size_t)s * stage_bytes + tile_bytes);

        // admitted positions of the tile as a 64-bit mask
        uint32_t r_lo = RANK_INVALID, r_hi = RANK_INVALID;
        {
            const int p0 = lane, p1 = lane + 32;
            const int rbase = sg.row_off + (lv.n_off - P.lv[sg.first_level].n_off) + a;
            if (hw0 + p0 < lv.HW) r_lo = sg.has_topk ? rk[p0] : (uint32_t)(rbase + (hw0 + p0) * P.A);
            if (p1 < TILE_T && hw0 + p1 < lv.HW) r_hi = sg.has_topk ? rk[p1] : (uint32_t)(rbase + (hw0 + p1) * P.A);
        }
        uint32_t m_lo = __ballot_sync(0xffffffffu, r_lo != RANK_INVALID);
        uint32_t m_hi = __ballot_sync(0xffffffffu, r_hi != RANK_INVALID);
        while (m_lo | m_hi) {
            int pos;
            if (m_lo) {
                pos = __ffs(m_lo) - 1;
                m_lo &= m_lo - 1;
            } else {
                pos = 32 + __ffs(m_hi) - 1;
                m_hi &= m_hi - 1;
            }
            const uint32_t r = __shfl_sync(0xffffffffu, pos < 32 ? r_lo : r_hi, pos & 31);
            uint32_t* mrow = P.mat + ((size_t)b * P.R + r) * P.C;

            // attributes 0..4: one lane each
            float act = 0.f;
            if (lane < 5) {
                float v = tile[lane * TILE_T + pos];
                act = (MODE == 0 || lane < 2 || lane == 4) ? c_sigmoid(v) : c_expf(v);
            }
            const float a0 = __shfl_sync(0xffffffffu, act, 0), a1 = __shfl_sync(0xffffffffu, act, 1);
            const float a2 = __shfl_sync(0xffffffffu, act, 2), a3 = __shfl_sync(0xffffffffu, act, 3);
            const float conf = __shfl_sync(0xffffffffu, act, 4);
            if (MODE == 1 && P.conf_thr > 0.f && !(conf >= P.conf_thr)) {  // yolo_head.py:365-376: row dropped
                for (int c = lane; c < P.C; c += 32) mrow[c] = SCORE_NONE;
                continue;
            }
            const int hw = hw0 + pos;
            const int y = hw / lv.W, x = hw - y * lv.W;
            float4 bx = decode_box<MODE>(lv, a, x, y, a0, a1, a2, a3);
            if (P.rescale) bx = rescale_box(bx, P.scale + 4 * b);
            if (lane == 0) {
                P.row_box[(size_t)b * P.R + r] = bx;
                if (!sg.has_topk) P.row_anchor[(size_t)b * P.R + r] = lv.n_off + hw * P.A + a;
            }
            bool any = false;
            if (P.agnostic) {
                // cls_pred = conf_pred[:, None]  (yolocsp_head.py:360): one class, score = objectness
                const bool pass = conf > P.score_thr;
                if (lane == 0) mrow[0] = pass ? __float_as_uint(conf) : SCORE_NONE;
                any = pass;
            } else {
                for (int c0 = 0; c0 < P.C; c0 += 32) {
                    const int c = c0 + lane;
                    bool pass = false;
                    if (c < P.C) {
                        const float sgm = c_sigmoid(tile[(5 + c) * TILE_T + pos]);
                        float score;
                        if (MODE == 0) {
                            score = fmul(sgm, conf);     // cls_pred *= conf_pred[:, None]   (yolocsp_head.py:358)
                            pass = score > P.score_thr;  // bbox_nms.py:54
                        } else {
                            pass = sgm > P.score_thr;  // threshold on the class score alone (bbox_nms.py:54) ...
                            score = fmul(sgm, conf);   // ... then scores * score_factors     (bbox_nms.py:57-62)
                        }
                        mrow[c] = pass ? __float_as_uint(score) : SCORE_NONE;
                    }
                    any |= __any_sync(0xffffffffu, pass);
                }
            }
            if (any && lane == 0) atomicMax(&P.img_max[b], f2ord(box_max(bx)));  // boxes.max() (mmcv batched_nms)
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
}

// Generic path: one thread per position, coalesced scalar loads, loop over classes. Used for levels whose
// plane stride is not 16-byte aligned (e.g. 19x19) and for dense admission (no top-k), where every anchor is
// computed anyway.
template <int MODE>
__global__ void __launch_bounds__(128) decode_ldg_kernel(const __grid_constant__ DevParams P) {
    const int t = blockIdx.x;
    int l = 0;
    for (int q = 0; q < P.L; ++q)
        if (!P.lv[q].use_tma && t >= P.lv[q].tile0) l = q;
    const LevelDev& lv = P.lv[l];
    const int loc = t - lv.tile0;
    const int plane = loc / lv.tpp;
    const int ht = loc - plane * lv.tpp;
    const int b = plane / P.A, a = plane - b * P.A;
    const int hw = ht * 128 + threadIdx.x;
    const SegDev& sg = P.seg[lv.seg];
    const float* slab = lv.ptr + (size_t)plane * P.NA * lv.HW;

    if (hw >= lv.HW) return;
    const uint32_t r = row_of(P, lv, b, a, hw);
    if (r == RANK_INVALID) return;
    uint32_t* mrow = P.mat + ((size_t)b * P.R + r) * P.C;
    const float t0 = __ldg(slab + 0 * (size_t)lv.HW + hw), t1 = __ldg(slab + 1 * (size_t)lv.HW + hw);
    const float t2 = __ldg(slab + 2 * (size_t)lv.HW + hw), t3 = __ldg(slab + 3 * (size_t)lv.HW + hw);
    const float conf = c_sigmoid(__ldg(slab + 4 * (size_t)lv.HW + hw));
    if (MODE == 1 && P.conf_thr > 0.f && !(conf >= P.conf_thr)) {
        for (int c = 0; c < P.C; ++c) mrow[c] = SCORE_NONE;
        return;
    }
    const float a0 = c_sigmoid(t0), a1 = c_sigmoid(t1);
    const float a2 = (MODE == 0) ? c_sigmoid(t2) : c_expf(t2);
    const float a3 = (MODE == 0) ? c_sigmoid(t3) : c_expf(t3);
    const int y = hw / lv.W, x = hw - y * lv.W;
    float4 bx = decode_box<MODE>(lv, a, x, y, a0, a1, a2, a3);
    if (P.rescale) bx = rescale_box(bx, P.scale + 4 * b);
    P.row_box[(size_t)b * P.R + r] = bx;
    if (!sg.has_topk) P.row_anchor[(size_t)b * P.R + r] = lv.n_off + hw * P.A + a;
    bool any = false;
    if (P.agnostic) {
        any = conf > P.score_thr;
        mrow[0] = any ? __float_as_uint(conf) : SCORE_NONE;
    } else {
        const float* cls = slab + 5 * (size_t)lv.HW + hw;
#pragma unroll 8
        for (int c = 0; c < P.C; ++c) {
            const float sgm = c_sigmoid(__ldg(cls + (size_t)c * lv.HW));
            float score;
            bool pass;
            if (MODE == 0) {
                score = fmul(sgm, conf);
                pass = score > P.score_thr;
            } else {
                pass = sgm > P.score_thr;
                score = fmul(sgm, conf);
            }
            mrow[c] = pass ? __float_as_uint(score) : SCORE_NONE;
            any |= pass;
        }
    }
    if (any) atomicMax(&P.img_max[b], f2ord(box_max(bx)));
}

// ------------------------------------------------------------------------------------------------
// K2: per-image NMS over the merged, score-ordered candidate stream
// ------------------------------------------------------------------------------------------------
// Suppression rule of mmcv batched_nms, both regimes:
//   n < split_thr : ONE greedy pass over all candidates; boxes carry the class offset idx * (max + 1) unless
//                   nms class_agnostic, so classes normally cannot touch — but the test is purely geometric,
//                   exactly like the reference (cross-class suppression possible when offset ranges overlap).
//   n >= split_thr: classes are processed independently (a kept box only suppresses boxes of its own class),
//                   still on the offset boxes.
// In both regimes the result order is (score desc, flat index asc) and only the first max_num are returned,
// so candidates are visited in that global order and the pass stops once `cap` boxes are kept.
__global__ void __launch_bounds__(NMS_THREADS) nms_image_kernel(const __grid_constant__ DevParams P) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    __shared__ TopSelSmem S;
    __shared__ unsigned s_sup;
    __shared__ int s_nk, s_cnt;
    const int cap = P.keep_cap;
    u64* keys = reinterpret_cast<u64*>(nms_smem);                 // [NMS_KCAP]
    u64* kkey = keys + NMS_KCAP;                                   // [cap]
    float* cx1 = reinterpret_cast<float*>(kkey + cap);             // [NMS_CH] x 5
    float* cy1 = cx1 + NMS_CH;
    float* cx2 = cy1 + NMS_CH;
    float* cy2 = cx2 + NMS_CH;
    float* car = cy2 + NMS_CH;
    int* ccl = reinterpret_cast<int*>(car + NMS_CH);               // [NMS_CH]
    float* kx1 = reinterpret_cast<float*>(ccl + NMS_CH);           // [cap] x 5
    float* ky1 = kx1 + cap;
    float* kx2 = ky1 + cap;
    float* ky2 = kx2 + cap;
    float* kar = ky2 + cap;
    int* kcl = reinterpret_cast<int*>(kar + cap);                  // [cap]

    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = NMS_THREADS / 32;
    const int C = P.C;
    const int slots = P.R * C;
    const uint32_t* mat = P.mat + (size_t)b * slots;
    const float4* row_box = P.row_box + (size_t)b * P.R;

    // pass 0: count the candidates and find the key range
    u64 kmin = ~0ull, kmax = 0ull;
    int cntl = 0;
    for (int i = tid; i < slots; i += NMS_THREADS) {
        const uint32_t sb = mat[i];
        if (sb != SCORE_NONE) {
            const u64 key = make_key(__uint_as_float(sb), (uint32_t)i);
            kmin = key < kmin ? key : kmin;
            kmax = key > kmax ? key : kmax;
            ++cntl;
        }
    }
    if (tid == 0) {
        s_cnt = 0;
        s_nk = 0;
        s_sup = 0u;
    }
    u64 gmin, gmax;
    block_minmax(kmin, kmax, gmin, gmax, S);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cntl += __shfl_xor_sync(0xffffffffu, cntl, o);
    if (lane == 0 && cntl) atomicAdd(&s_cnt, cntl);
    __syncthreads();
    const int ntot = s_cnt;
    if (tid == 0 && P.o_ncand) P.o_ncand[b] = ntot;
    if (ntot == 0) {
        if (tid == 0) P.o_count[b] = 0;
        return;
    }
    const bool per_class = !(ntot < P.split_thr);  // regime of mmcv batched_nms
    const bool use_off = !P.nms_agnostic;
    const float mp1 = fadd(ord2f(P.img_max[b]), 1.0f);  // max_coordinate + 1
    const float thr = P.iou_thr, foff = P.foff;

    auto fetch = [=](int i, u64& key) -> bool {
        const uint32_t sb = mat[i];
        key = make_key(__uint_as_float(sb), (uint32_t)i);
        return sb != SCORE_NONE;
    };

    int processed = 0;
    u64 lo = gmin;
    while (processed < ntot && s_nk < cap) {
        const int want = min(NMS_CH, ntot - processed);
        int got = select_sorted_prefix(fetch, slots, lo, gmax, want, keys, NMS_KCAP, S);
        const int m = got < NMS_CH ? got : NMS_CH;  // boxes staged this round (a prefix of the sorted order)
        if (m == 0) break;
        for (int i = tid; i < m; i += NMS_THREADS) {
            const uint32_t flat = key_flat(keys[i]);
            const int r = (int)(flat / (uint32_t)C), c = (int)(flat - (uint32_t)r * (uint32_t)C);
            const float4 bx = row_box[r];
            float x1 = bx.x, y1 = bx.y, x2 = bx.z, y2 = bx.w;
            if (use_off) {  // boxes + idxs.to(boxes) * (max_coordinate + 1)
                const float off = fmul((float)c, mp1);
                x1 = fadd(x1, off);
                y1 = fadd(y1, off);
                x2 = fadd(x2, off);
                y2 = fadd(y2, off);
            }
            cx1[i] = x1;
            cy1[i] = y1;
            cx2[i] = x2;
            cy2[i] = y2;
            car[i] = box_area(x1, y1, x2, y2, foff);
            ccl[i] = c;
        }
        __syncthreads();
        for (int s0 = 0; s0 < m; s0 += 32) {
            const int nk = s_nk;
            if (nk >= cap) break;
            const int j = s0 + lane;
            const bool valid = j < m;
            Box bj;
            bj.x1 = valid ? cx1[j] : 0.f;
            bj.y1 = valid ? cy1[j] : 0.f;
            bj.x2 = valid ? cx2[j] : 0.f;
            bj.y2 = valid ? cy2[j] : 0.f;
            bj.area = valid ? car[j] : 0.f;
            const int cj = valid ? ccl[j] : -1;
            // phase A: against the kept list, kept boxes strided over the warps
            bool sup = false;
            for (int k = warp; k < nk; k += NW) {
                Box bk;
                bk.x1 = kx1[k];
                bk.y1 = ky1[k];
                bk.x2 = kx2[k];
                bk.y2 = ky2[k];
                bk.area = kar[k];
                const bool same = !per_class || (kcl[k] == cj);
                if (valid && !sup && same && iou_gt(bk, bj, thr, foff)) sup = true;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, sup);
            if (lane == 0 && bal) atomicOr(&s_sup, bal);
            __syncthreads();
            // phase B: inside the 32-box group, sequential over i, parallel over j > i (warp 0)
            if (warp == 0) {
                bool alive = valid && !((s_sup >> lane) & 1u);
                const int room = cap - nk;
                int kept_here = 0;
                for (int i = 0; i < 32; ++i) {
                    const bool ai = __shfl_sync(0xffffffffu, alive, i) != 0;
                    if (!ai) continue;
                    if (++kept_here > room) break;  // later boxes cannot enter the first `cap` kept
                    Box bi;
                    bi.x1 = __shfl_sync(0xffffffffu, bj.x1, i);
                    bi.y1 = __shfl_sync(0xffffffffu, bj.y1, i);
                    bi.x2 = __shfl_sync(0xffffffffu, bj.x2, i);
                    bi.y2 = __shfl_sync(0xffffffffu, bj.y2, i);
                    bi.area = __shfl_sync(0xffffffffu, bj.area, i);
                    const int ci = __shfl_sync(0xffffffffu, cj, i);
                    const bool same = !per_class || (ci == cj);
                    if (lane > i && alive && same && iou_gt(bi, bj, thr, foff)) alive = false;
                }
                const unsigned km = __ballot_sync(0xffffffffu, alive);
                const int rnk = __popc(km & ((1u << lane) - 1u));
                if (alive && rnk < room) {
                    const int kidx = nk + rnk;
                    kx1[kidx] = bj.x1;
                    ky1[kidx] = bj.y1;
                    kx2[kidx] = bj.x2;
                    ky2[kidx] = bj.y2;
                    kar[kidx] = bj.area;
                    kcl[kidx] = cj;
                    kkey[kidx] = keys[j];
                }
                if (lane == 0) {
                    s_nk = nk + min(__popc(km), room);
                    s_sup = 0u;
                }
            }
            __syncthreads();
        }
        processed += m;
        lo = keys[m - 1] + 1ull;
        __syncthreads();
    }
    // outputs: dets = (boxes[keep], scores[keep]), labels[keep]  (bbox_nms.py:84-93)
    int nk = s_nk;
    if (nk > P.out_cap) {
        nk = P.out_cap;
        if (tid == 0) atomicMax(P.o_status, 3);  // YOLOPP_E_OVERFLOW
    }
    for (int i = tid; i < nk; i += NMS_THREADS) {
        const u64 key = kkey[i];
        const uint32_t flat = key_flat(key);
        const int r = (int)(flat / (uint32_t)C), c = (int)(flat - (uint32_t)r * (uint32_t)C);
        const float4 bx = row_box[r];
        float* d = P.o_dets + ((size_t)b * P.out_cap + i) * 5;
        d[0] = bx.x;
        d[1] = bx.y;
        d[2] = bx.z;
        d[3] = bx.w;
        d[4] = key_score(key);
        P.o_labels[(size_t)b * P.out_cap + i] = (long long)c;
        if (P.o_anchors) P.o_anchors[(size_t)b * P.out_cap + i] = P.row_anchor[(size_t)b * P.R + r];
        if (P.o_rows) P.o_rows[(size_t)b * P.out_cap + i] = r;
    }
    if (tid == 0) P.o_count[b] = nk;
}

// ------------------------------------------------------------------------------------------------
// small standalone kernels
// ------------------------------------------------------------------------------------------------
__global__ void unary_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int op) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) out[i] = op == 0 ? c_sigmoid(in[i]) : c_expf(in[i]);
}

// YOLOV4BBoxCoder.decode / YOLOBBoxCoder.decode on explicit (anchor, pred) pairs
__global__ void coder_decode_kernel(int mode, const float4* __restrict__ anchors, const float4* __restrict__ pred,
                                    float stride, long long n, float4* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        float4 a = anchors[i], p = pred[i];
        float xc = fmul(fadd(a.x, a.z), 0.5f), yc = fmul(fadd(a.y, a.w), 0.5f);
        float w = fsub(a.z, a.x), h = fsub(a.w, a.y);
        float xcp, ycp, wp, hp;
        if (mode == 0) {
            xcp = fadd(fmul(p.x, stride), xc);
            ycp = fadd(fmul(p.y, stride), yc);
            wp = fmul(p.z, w);
            hp = fmul(p.w, h);
        } else {
            xcp = fadd(fmul(fsub(p.x, 0.5f), stride), xc);
            ycp = fadd(fmul(fsub(p.y, 0.5f), stride), yc);
            wp = fmul(c_expf(p.z), w);
            hp = fmul(c_expf(p.w), h);
        }
        float hw_ = fmul(wp, 0.5f), hh_ = fmul(hp, 0.5f);
        out[i] = make_float4(fsub(xcp, hw_), fsub(ycp, hh_), fadd(xcp, hw_), fadd(ycp, hh_));
    }
}

__device__ __forceinline__ u64 splitmix64(u64 x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

struct Synth3 {
    float mean[3];  // box (attr 0..3), objectness (attr 4), class (attr 5..)
    float std[3];
};

// bit-reproducible synthetic head tensor (see yolopp.h)
__global__ void synth_kernel(float* __restrict__ out, long long n, int na, int hw, Synth3 st, u64 seed) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        u64 x = splitmix64(seed + (u64)(i + 1) * 0x9E3779B97F4A7C15ULL);
        int s = (int)(x & 0xFFFF) + (int)((x >> 16) & 0xFFFF) + (int)((x >> 32) & 0xFFFF) + (int)(x >> 48);
        float z = fmul((float)(s - 131070), 2.64290273e-05f);
        int k = (int)((i / hw) % na);
        int g = k < 4 ? 0 : (k == 4 ? 1 : 2);
        out[i] = fadd(st.mean[g], fmul(st.std[g], z));
    }
}

}  // namespace ypp
