// yolopp_kernels.cuh — the kernels of the post-processing path (sm_100a). Included by yolopp_capi.cu.
//
//   K0 select_kernel      objectness top-k per (image, segment) on the raw logits: staged in shared memory, sampled
//                         histogram cut, sort of the survivors -> rank map (exact fallback over 64-bit keys)
//   K1 decode_tma_kernel  persistent, TMA-streamed fused decode + score + threshold -> dense (row, class) score matrix
//      decode_dense / decode_rows / decode_ldg_kernel   the same semantics for densely admitted levels, channels-last
//                         tensors, and levels with more than 256 attributes
//   K2 nms_image_kernel   per image: the candidates that can reach the first max_per_img are fetched as a prefix of the
//                         global (score desc, flat index asc) order (row-maximum bound + bulk-copied matrix rows);
//                         greedy NMS with the reference's regime rule — n >= split_thr: classes independent, resolved
//                         in parallel (pair masks per class), only the kept boxes sorted; n < split_thr: one problem
//                         over class-offset boxes, walked in sorted groups of 64; stops at max_per_img kept, writes
//                         the outputs and their grouping by label.
#pragma once
#include "yolopp_device.cuh"

namespace ypp {

constexpr int MAXL = 8;
constexpr int MAXA = 8;
constexpr uint32_t RANK_INVALID = 0xFFFFFFFFu;

// Positions per TMA tile: 64 (two 32-wide boxes with 128-byte rows, SWIZZLE_128B; 4 stages) or 32 (one box; 8 stages),
// chosen per plan (template parameter of the decode kernel). The ring is in-order: a tile with many admitted anchors
// holds its stage through several register batches and everything behind it waits; half-size tiles in twice the
// stages keep the same bytes in flight but halve what a slow tile blocks (608^2 b64: decode 100.2 -> 94.7 us; YOLOv3
// 640^2: 284 -> 237 us). Where hardly any tile holds an admitted anchor (1280^2: 1 % of the positions) the per-tile
// costs dominate instead and 64-position tiles stay ahead (663 vs 712 us).
__host__ __device__ constexpr int dec_stages_of(int tile_t) { return tile_t == 64 ? 4 : 8; }
constexpr int TILE_SUB = 32;     // positions per box
// positions per GATHER tile (levels without a tensor map): nothing is streamed for them, a tile is just the unit of work
// one consumer warp takes — and keeps for several DRAM round trips per admitted anchor. They sit at the head of the tile
// sequence, so their number per CTA decides how many consumer warps are tied up while the ring fills: 8 per CTA
// (32 positions) tie up every warp for ~9 us; measured decode 95.7 / 94.9 / 96.4 / 117 us at 32 / 64 / 128 / 256.
#ifndef YPP_GATHER_T
#define YPP_GATHER_T 64
#endif
constexpr int GATHER_T = YPP_GATHER_T;
constexpr int DEC_PWARPS = 2;    // producer warps per CTA: warp p issues the tiles of iterations it == p (mod 2)
#ifndef YPP_CWARPS
#define YPP_CWARPS 8
#endif
#ifndef YPP_SLEEP
#define YPP_SLEEP 64
#endif
constexpr int DEC_CWARPS = YPP_CWARPS;  // consumer warps per CTA, one whole tile per warp
constexpr int DEC_BATCH = 4;     // admitted anchors whose logits a consumer pulls into registers at once
constexpr int DEC_ROUNDS = 3;    // class sweeps held in registers (C <= 96); wider heads read the tile in place
constexpr int DEC_THREADS = 32 * (DEC_PWARPS + DEC_CWARPS);

constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAX_K = 4096;  // largest nms_pre the select kernel sorts in shared memory

constexpr int NMS_THREADS = 512;
constexpr int NMS_KCAP = 2048;   // sorted-chunk buffer (keys)
constexpr int NMS_CH = 1024;     // candidates staged (boxes) per chunk
constexpr int NMS_G = 64;        // candidates resolved per round
constexpr int NMS_MAX_KEEP = 4096;

struct LevelDev {
    const float* ptr;
    int H, W, HW;
    int n_off;    // offset of this level in the concatenated anchor index n = n_off + (y*W+x)*A + a
    int m_off;    // offset in the plane-major, 4-element aligned index m = m_off + a*HW + hw
    int seg;      // top-k segment this level belongs to
    int use_tma;  // persistent decode_tma_kernel: 1 = TMA tiles (plane stride 16-byte aligned), 3 = quad-row TMA tiles
                  // (unaligned plane stride: the tensor is viewed as rows of FOUR planes, whose stride always is a
                  // multiple of 16 bytes), 2 = gather tiles (no tensor map possible); 0: decode_ldg / decode_dense kernel
    int dense;    // use_tma == 0 only: 1 = most anchors are admitted -> decode_dense_kernel (thread per position)
    int sel_bulk; // top-k staging: the level's objectness planes are 16-byte aligned on both sides -> one bulk copy per plane
    int qrows;    // use_tma == 3: rows of the quad-row view = floor(B * A * NA / 4)
    int tile0;    // first tile id of this level in its kernel's tile enumeration
    int tpp;      // tiles per plane
    int tile_pos; // positions per tile on this level: the kernel's tile size, or GATHER_T on a gathered level
    float sx, sy; // anchor-generator strides
    float cstride;// coder stride
    float inv_w;  // 1 / W (position -> (x, y) without an integer division)
    float base[MAXA][4];
};

struct SegDev {
    int first_level, num_levels;
    int N;        // anchors in the segment
    int k;        // rows kept (k == N when no top-k runs)
    int row_off;  // first row of the segment
    int has_topk;
    int m_begin, m_end;  // plane-major index range of the segment
    unsigned sel_bulk_bytes;  // top-k staging: bytes the bulk copies of one image bring in (levels with sel_bulk)
};

#ifdef YPP_PROFILE
// per-CTA phase timestamps of the per-image kernels (profiling build only): [kernel][image][phase]
__device__ long long g_phase[2][256][16];
#define YPP_PHASE(k, blk, i) do { if (threadIdx.x == 0 && (blk) < 256) g_phase[k][blk][i] = clock64(); } while (0)
// finer stamps inside the NMS kernel's first chunk: [image][stamp]
__device__ long long g_sub[256][32];  // (stamps 0..16 in use)
#define YPP_SUB(i) do { if (threadIdx.x == 0 && blockIdx.x < 256) g_sub[blockIdx.x][i] = clock64(); } while (0)
#define YPP_SUBY(i) do { if (threadIdx.x == 0 && blockIdx.y < 256) g_sub[blockIdx.y][i] = clock64(); } while (0)
#define YPP_SUBV(i, v) do { if (threadIdx.x == 0 && blockIdx.x < 256) g_sub[blockIdx.x][i] = (long long)(v); } while (0)
#else
#define YPP_PHASE(k, blk, i) do { } while (0)
#define YPP_SUB(i) do { } while (0)
#define YPP_SUBY(i) do { } while (0)
#define YPP_SUBV(i, v) do { } while (0)
#endif

struct DevParams {
    int mode, B, L, A, C, NA, agnostic;  // C = effective classes (1 when class agnostic)
    int nsegs, ntopk;
    int topk_segs[MAXL];
    int N, M_pad, R;
    float score_thr, conf_thr, iou_thr, foff;
    float nms_score_thr;  // mmcv NMSop score_threshold (0: off)
    int split_thr, nms_agnostic, m_eff, keep_cap, out_cap, rescale;
    int sel_kcap;  // key buffer (power of two) of the select kernel
    int sel_stage; // 32-bit slots of the select kernel's logit staging buffer (0: exact path only)
    int sel_stride;// 1: whole segments are staged; 2^j: every 2^j-th logit is staged (sample), the rest is streamed
    int nhwc;      // level tensors are channels-last memory (B, H, W, A*NA): row-driven decode, no rank map
    int nms_rowkeys_off;  // byte offset of the sorted row-best keys inside the NMS kernel's dynamic smem
    int nms_stage_off, nms_stage_rows;  // staging buffer of the candidate scan: nms_stage_rows score-matrix rows
    int tma_tiles, ldg_blocks, dense_tiles;
    int dec_quad;            // some level is streamed as quad-row tiles (stage geometry)
    int tile_t;              // positions per tile of the persistent decode kernel (64 or 32)
    unsigned* tile_ctr;      // workspace: next unclaimed position of the decode kernel's tile sequence (set by select_kernel)
    unsigned dec_first;      // positions [0, dec_first) of the tile sequence are dealt round-robin, the rest is claimed
    LevelDev lv[MAXL];
    SegDev seg[MAXL];
    // workspace
    u64* ckey;           // [B][M_pad]  (~ord(conf) << 32 | n), ~0 in the alignment padding
    uint32_t* rank;      // [B][M_pad]  row of each anchor (plane-major), RANK_INVALID when not admitted
    int* row_anchor;     // [B][R]
    float4* row_box;     // [B][R]
    uint32_t* mat;       // [B][R][C]   score bits of candidate (row, class), SCORE_NONE otherwise
    uint4* row_stat;     // [B][R]  per row: (max ord(score), max ~ord(score), #candidates, 0) — plain stores, the
                         //         per-image reduction happens in the NMS kernel (no same-address atomics)
    const float* scale;  // [B][4] or null
    // standalone NMS entry points (yolopp_batched_nms / yolopp_multiclass_nms): B = 1
    int generic;               // 1: candidates are the n input boxes themselves (C = 1), scores may be any float
    int boxes_per_class;       // multi_bboxes is (n, 4C): the box of candidate (row, class) is boxes[row*C + class]
    int num_labels;            // generic: labels are in [0, num_labels)
    const float* g_scores;     // generic: [n]
    const long long* g_labels; // generic: [n] or null (one class)
    long long* o_keep;         // kept candidate (flat) indices, in output order; may be null
    unsigned char* nms_kept;   // kept list in global memory, [B][nms_kept_stride] bytes (keep_cap > NMS_MAX_KEEP), else null
    long long nms_kept_stride;
    // outputs
    float* o_dets;
    long long* o_labels;
    int* o_anchors;
    int* o_rows;
    int* o_count;
    int* o_ncand;
    int* o_status;
    float* o_cls_dets;      // [B][out_cap][5] detections grouped by label (bbox2result), may be null
    int* o_cls_offsets;     // [B][C+1] first row of each label's group in o_cls_dets
};

// ---- key sources of the bucket select (see select_sorted_prefix) --------------------------------------------
// objectness keys of a top-k segment, two per 16-byte load (the ~0 alignment padding is above every `hi`)
struct CkeySource {
    typedef ulonglong2 Raw;
    static constexpr int V = 2, U = 4;
    const ulonglong2* p;
    int n;
    __device__ __forceinline__ Raw load(int g) const { return p[g]; }
    __device__ __forceinline__ unsigned exact(const Raw& r, int, u64 lo, u64 hi) const {
        return ((r.x >= lo && r.x <= hi) ? 1u : 0u) | ((r.y >= lo && r.y <= hi) ? 2u : 0u);
    }
    __device__ __forceinline__ u64 key_at(int g, int v) const { return reinterpret_cast<const u64*>(p)[2 * g + v]; }
    __device__ __forceinline__ int groups() const { return n; }
};
// every `stride`-th objectness key (pivot sample)
struct CkeySampleSource {
    typedef u64 Raw;
    static constexpr int V = 1, U = 4;
    const u64* p;
    int n, stride;
    __device__ __forceinline__ Raw load(int g) const { return p[(size_t)g * stride]; }
    __device__ __forceinline__ unsigned exact(const Raw& r, int, u64 lo, u64 hi) const { return (r >= lo && r <= hi) ? 1u : 0u; }
    __device__ __forceinline__ u64 key_at(int g, int) const { return p[(size_t)g * stride]; }
    __device__ __forceinline__ int groups() const { return n; }
};
// Window test shared by the two score-matrix sources. Scores are non-negative floats (bit patterns order like the
// values), key = ((0x7FFFFFFF - bits) << 32) | flat, so lo <= key <= hi is a window on the raw word plus a flat
// index test on its two end values; SCORE_NONE = 0xFFFFFFFF is above every window. Branch-free: the hot loop of
// the candidate scan is "load, subtract, compare".
struct ScoreWindow {
    uint32_t worst, span, best;  // raw-word window [worst, worst + span], best = worst + span
    uint32_t lo_flat, hi_flat;   // flat bounds that apply at word == best / word == worst
    bool ends;                   // some flat bound is not trivial
    __device__ __forceinline__ void set(u64 lo, u64 hi) {
        best = (~(uint32_t)(lo >> 32)) & 0x7FFFFFFFu;
        worst = (~(uint32_t)(hi >> 32)) & 0x7FFFFFFFu;
        span = best - worst;
        lo_flat = (uint32_t)lo;
        hi_flat = (uint32_t)hi;
        ends = lo_flat != 0u || hi_flat != 0xFFFFFFFFu;
    }
    __device__ __forceinline__ unsigned test4(const uint4& w, uint32_t flat0) const {
        unsigned mk = (w.x - worst <= span ? 1u : 0u) | (w.y - worst <= span ? 2u : 0u) | (w.z - worst <= span ? 4u : 0u) |
                      (w.w - worst <= span ? 8u : 0u);
        if (ends) {
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const bool out = (ws[v] == best && flat0 + v < lo_flat) || (ws[v] == worst && flat0 + v > hi_flat);
                mk &= out ? ~(1u << v) : ~0u;
            }
        }
        return mk;
    }
};
__device__ __forceinline__ u64 score_key(uint32_t bits, uint32_t flat) {
    return ((u64)((~bits) & 0x7FFFFFFFu) << 32) | (u64)flat;  // == make_key for a non-negative score
}
// score matrix of one image, four entries per 16-byte load. `windowed` = scores are known non-negative (the
// pipeline's own matrix); the standalone NMS entries accept any float and build the 64-bit key per element.
struct MatSource {
    typedef uint4 Raw;
    static constexpr int V = 4, U = 4;
    const uint32_t* m;
    int slots;
    bool vec4, windowed;
    ScoreWindow win;
    __device__ __forceinline__ Raw load(int g) const {
        if (vec4) return reinterpret_cast<const uint4*>(m)[g];
        uint4 r;
        r.x = (g * 4 + 0 < slots) ? m[g * 4 + 0] : SCORE_NONE;
        r.y = (g * 4 + 1 < slots) ? m[g * 4 + 1] : SCORE_NONE;
        r.z = (g * 4 + 2 < slots) ? m[g * 4 + 2] : SCORE_NONE;
        r.w = (g * 4 + 3 < slots) ? m[g * 4 + 3] : SCORE_NONE;
        return r;
    }
    __device__ __forceinline__ unsigned exact(const Raw& r, int g, u64 lo, u64 hi) const {
        if (windowed) return win.test4(r, (uint32_t)(g * 4));
        const uint32_t ws[4] = {r.x, r.y, r.z, r.w};
        unsigned mk = 0u;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (ws[v] != SCORE_NONE) {
                const u64 k = make_key(__uint_as_float(ws[v]), (uint32_t)(g * 4 + v));
                mk |= (k >= lo && k <= hi) ? (1u << v) : 0u;
            }
        }
        return mk;
    }
    __device__ __forceinline__ u64 key_at(int g, int v) const {
        const uint32_t flat = (uint32_t)(g * 4 + v);
        return make_key(__uint_as_float(m[flat]), flat);
    }
    __device__ __forceinline__ int groups() const { return (slots + 3) / 4; }
};
// score-matrix entries of a LIST of rows (the rows whose best score can still matter), four per 16-byte load.
struct RowListSource {
    typedef uint4 Raw;
    static constexpr int V = 4, U = 4;
    const uint32_t* m;   // score matrix of the image
    const u64* rows;     // low word = row index
    int nrows, C;
    int nsub;            // 32-group (= 128-class) slabs per row: group g -> slab g >> 5, chunk g & 31
    bool vec4;           // C % 4 == 0: rows are 16-byte aligned
    ScoreWindow win;
    // flat index of the group's first entry; a warp's 32 consecutive groups are the 32 chunks of ONE slab of one
    // row (no division on the load path); chunks past the row's end report C (nothing to load)
    __device__ __forceinline__ uint32_t flat0(int g, int& cls0) const {
        const int slab = g >> 5, jl = g & 31;
        const int i = nsub == 1 ? slab : slab / nsub;
        const int j = nsub == 1 ? jl : (slab - i * nsub) * 32 + jl;
        cls0 = 4 * j;
        return (uint32_t)rows[i] * (uint32_t)C + (uint32_t)(4 * j);
    }
    __device__ __forceinline__ Raw load(int g) const {
        int c0;
        const uint32_t f0 = flat0(g, c0);
        Raw r = make_uint4(SCORE_NONE, SCORE_NONE, SCORE_NONE, SCORE_NONE);
        if (c0 < C) {
            const uint32_t* p = m + f0;
            if (vec4) {
                r = *reinterpret_cast<const uint4*>(p);
            } else {
                r.x = p[0];
                r.y = (c0 + 1 < C) ? p[1] : SCORE_NONE;
                r.z = (c0 + 2 < C) ? p[2] : SCORE_NONE;
                r.w = (c0 + 3 < C) ? p[3] : SCORE_NONE;
            }
        }
        return r;
    }
    __device__ __forceinline__ unsigned exact(const Raw& r, int g, u64, u64) const {
        int c0;
        const uint32_t f0 = win.ends ? flat0(g, c0) : 0u;
        return win.test4(r, f0);
    }
    __device__ __forceinline__ u64 key_at(int g, int v) const {
        int c0;
        const uint32_t flat = flat0(g, c0) + (uint32_t)v;
        return score_key(m[flat], flat);
    }
    __device__ __forceinline__ int groups() const { return nrows * nsub * 32; }
};
// ------------------------------------------------------------------------------------------------
// K0: objectness top-k
// ------------------------------------------------------------------------------------------------
// (level, anchor, position) of plane-major slot m of a segment; false for the alignment padding between levels
__device__ __forceinline__ bool seg_slot(const DevParams& P, const SegDev& sg, int m, int& l, int& a, int& hw) {
#pragma unroll 1
    for (int li = sg.num_levels - 1; li >= 0; --li) {
        const LevelDev& lv = P.lv[sg.first_level + li];
        if (m >= lv.m_off) {
            int e = m - lv.m_off;
            if (e >= P.A * lv.HW) return false;
            int aa = 0;
#pragma unroll 1
            while (e >= lv.HW) {
                e -= lv.HW;
                ++aa;
            }
            l = sg.first_level + li;
            a = aa;
            hw = e;
            return true;
        }
    }
    return false;
}

// address of the objectness logit of (level, image, anchor, position): NCHW (B, A*NA, H, W) or, with P.nhwc,
// the channels-last memory of the same logical tensor (B, H, W, A*NA)
__device__ __forceinline__ const float* obj_addr(const DevParams& P, const LevelDev& lv, int b, int a, int hw) {
    if (P.nhwc) return lv.ptr + (((size_t)b * lv.HW + hw) * P.A + a) * P.NA + 4;
    return lv.ptr + ((size_t)(b * P.A + a) * P.NA + 4) * lv.HW + hw;
}

// rank map + row -> anchor table from sorted top-k keys: sel[0..k) are the rows row0 .. row0+k-1 of the segment
__device__ __forceinline__ void select_write(const DevParams& P, const SegDev& sg, int b, const u64* sel, int k, int row0 = 0) {
    uint32_t* rank = P.rank + (size_t)b * P.M_pad;
    const int first = sg.first_level, nl = sg.num_levels, A = P.A;
#pragma unroll 1
    for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
        const int n = (int)(uint32_t)sel[i];
        int l = first;
#pragma unroll 1
        for (int q = nl - 1; q >= 0; --q)
            if (n >= P.lv[first + q].n_off) {
                l = first + q;
                break;
            }
        const LevelDev& lv = P.lv[l];
        const int loc = n - lv.n_off;
        const int hw = loc / A, a = loc - hw * A;
        if (!P.nhwc) rank[lv.m_off + a * lv.HW + hw] = (uint32_t)(sg.row_off + row0 + i);  // (the NHWC decode is row driven)
        P.row_anchor[(size_t)b * P.R + sg.row_off + row0 + i] = n;
    }
}

// staging mode of a segment with M plane-major slots: 1 = staged whole, 2^j = 1-in-2^j sample + streamed pass
__device__ __forceinline__ int sel_stride_of(const DevParams& P, int M) {
    return (M <= P.sel_stage && M <= 32768) ? 1 : P.sel_stride;
}

// Fast path of the objectness top-k: work on the RAW logits. sigmoid is monotone, so the k best confidences
// belong to the k best logits — except for ties: distinct logits can round to the same confidence, and the
// canonical order breaks confidence ties by anchor index. Hence:
//   1. STAGED (the segment fits the staging buffer, P.sel_stride == 1): the segment's objectness logits are copied
//      to shared memory (bulk copies per plane / 4-byte async copies: one DRAM round trip for the whole segment)
//      and mapped to order-preserving integers; a histogram of a 1-in-8 sample picks a cut that keeps ~1.3 k
//      logits (the "stash").
//      STREAMED (larger segments, e.g. 100 800 anchors at 1280^2; P.sel_stride = 2^j): only every sel_stride-th
//      logit is staged; the histogram of that sample picks the cut (k + 4 sqrt(k * stride) expected survivors),
//      then ONE pass over the segment in global memory stashes every logit above the cut.
//   2. only the stash goes through sigmoid and becomes (~ord(conf) << 32 | anchor) keys, which are sorted;
//   3. the result is exact iff every logit outside the stash is strictly worse than the k-th key: checked with
//      a 16-ulp guard band on the confidence of the cut itself (covers ties and any last-ulp non-monotonicity
//      of the polynomial). Otherwise — saturated / mass-tied objectness — the caller runs the exact path over
//      all confidences.
// Returns true when the top-k has been written.
__device__ __noinline__ bool select_fast(const DevParams& P, const SegDev& sg, int b, u64* sel, u64* tmp, uint32_t* ox, int kcap,
                                         TopSelSmem& S) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    static_assert(SEL_THREADS * 2 == TS_BINS, "two histogram bins per thread in the pivot scan");
    const int M = sg.m_end - sg.m_begin;
    const int stride = sel_stride_of(P, M);
    const bool staged = stride == 1;
    int ns = staged ? M : 0;  // staged slots (streamed: counted while the sample is issued)
    uint32_t* rank = P.rank + (size_t)b * P.M_pad + sg.m_begin;
#pragma unroll 1
    for (int i = tid; i < TS_BINS; i += SEL_THREADS) S.hist[i] = 0;
    if (tid == 0) S.count = 0;
    uint64_t* bar = reinterpret_cast<uint64_t*>(&S.bar);
    YPP_SUBY(20);
    if (staged) {
        // stage the logits (slot order = plane-major order of the segment): one bulk copy per objectness plane
        // where planes are 16-byte aligned (issued by warp 0, one plane per lane), 4-byte async copies otherwise
        // (which levels qualify and how many bytes that makes is worked out on the host: the start of this kernel is
        // a serial stretch of cold code, every instruction taken out of it counts)
        if (tid == 0) {
            mbar_init(bar, 1);
            fence_mbar_init();
            mbar_arrive_expect_tx(bar, sg.sel_bulk_bytes);
        }
        YPP_SUBY(21);
        __syncthreads();
        YPP_SUBY(22);
        if (warp == 0) {
            // bulk copies: one lane per (level, anchor) plane, all issued at once
#pragma unroll 1
            for (int q = lane; q < sg.num_levels * P.A; q += 32) {
                int li = 0, a = q;
#pragma unroll 1
                while (a >= P.A) {
                    a -= P.A;
                    ++li;
                }
                const LevelDev& lv = P.lv[sg.first_level + li];
                if (lv.sel_bulk)
                    bulk_load_1d(ox + (lv.m_off - sg.m_begin) + a * lv.HW,
                                 lv.ptr + ((size_t)(b * P.A + a) * P.NA + 4) * lv.HW, (uint32_t)lv.HW * 4u, bar);
            }
        }
#pragma unroll 1
        for (int li = 0; li < sg.num_levels; ++li) {
            const LevelDev& lv = P.lv[sg.first_level + li];
            const int AHW = P.A * lv.HW, s0 = lv.m_off - sg.m_begin;
            if (!lv.sel_bulk) {
#pragma unroll 1
                for (int e = tid; e < AHW; e += SEL_THREADS) {
                    int a = 0, hw = e;
#pragma unroll 1
                    while (hw >= lv.HW) {
                        hw -= lv.HW;
                        ++a;
                    }
                    cp_async4(ox + s0 + e, obj_addr(P, lv, b, a, hw));
                }
            }
            const int pad0 = s0 + AHW, pad1 = li + 1 < sg.num_levels ? P.lv[sg.first_level + li + 1].m_off - sg.m_begin : M;
            if (tid < pad1 - pad0) ox[pad0 + tid] = 0xFFFFFFFFu;  // alignment padding: becomes ord 0, never eligible
            YPP_SUBY(23 + li);
        }
        YPP_PHASE(0, b, 5);
        cp_async_wait_all();
        mbar_wait(bar, 0u);
    } else {
        // sample 1 in `stride` logits of every plane: four consecutive logits every 4 * stride (16-byte copies) where
        // the plane is 16-byte aligned, single logits every `stride` otherwise. `ns` = samples staged.
        ns = 0;
#pragma unroll 1
        for (int li = 0; li < sg.num_levels; ++li) {
            const LevelDev& lv = P.lv[sg.first_level + li];
            const bool vec = !P.nhwc && ((lv.HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(lv.ptr) & 15) == 0);
#pragma unroll 1
            for (int a = 0; a < P.A; ++a) {
                if (vec) {
                    const float* plane = obj_addr(P, lv, b, a, 0);
                    const int n4 = lv.HW / (4 * stride);
#pragma unroll 1
                    for (int i = tid; i < n4; i += SEL_THREADS) cp_async16(ox + ns + 4 * i, plane + (size_t)4 * stride * i);
                    ns += 4 * n4;
                } else {
                    const int n1 = lv.HW / stride, n1p = (n1 + 3) & ~3;  // (running offset stays 16-byte aligned)
#pragma unroll 1
                    for (int i = tid; i < n1; i += SEL_THREADS) cp_async4(ox + ns + i, obj_addr(P, lv, b, a, i * stride));
                    if (tid < n1p - n1) ox[ns + n1 + tid] = 0xFFFFFFFFu;  // padding: becomes ord 0, never counted
                    ns += n1p;
                }
            }
        }
        YPP_PHASE(0, b, 5);
        cp_async_wait_all();
    }
    __syncthreads();
    YPP_PHASE(0, b, 6);
    // The staging buffer keeps the RAW logits (0xFFFFFFFF = alignment padding -> ord 0, never eligible); the order-
    // preserving integers are formed on the fly. Range and histogram come from the same sample (staged: 1 in 8 slots;
    // streamed: everything that was staged) — logits outside the sample's range are simply above / below every bin.
    auto ord_at = [&](int m) -> uint32_t {
        const uint32_t raw = ox[m];
        return raw == 0xFFFFFFFFu ? 0u : f2ord(__uint_as_float(raw));
    };
    const int hstep = staged ? 8 : 1;
    uint32_t omin = 0xFFFFFFFFu, omax = 0u;
#pragma unroll 1
    for (int m = tid * hstep; m < ns; m += SEL_THREADS * hstep) {
        const uint32_t o = ord_at(m);
        omin = (o && o < omin) ? o : omin;
        omax = o > omax ? o : omax;
    }
    omin = __reduce_min_sync(0xffffffffu, omin);
    omax = __reduce_max_sync(0xffffffffu, omax);
    YPP_PHASE(0, b, 7);
    uint32_t* red = reinterpret_cast<uint32_t*>(&S.red[0][0]);
    if (lane == 0) {
        red[warp] = omin;
        red[32 + warp] = omax;
    }
    __syncthreads();
    omin = red[lane];
    omax = red[32 + lane];
    omin = __reduce_min_sync(0xffffffffu, omin);
    omax = __reduce_max_sync(0xffffffffu, omax);
    YPP_PHASE(0, b, 1);
    if (omax == 0u) return false;  // nothing staged (cannot happen for a non-empty segment)
    const uint32_t range = omax - omin;
    const int shift = range >= (uint32_t)TS_BINS ? (32 - __clz(range)) - 11 : 0;
    // sampled histogram of (omax - o): bin 0 holds the best logits (staged: 1-in-8 of the slots; streamed: the
    // whole sample)
#pragma unroll 1
    for (int m = tid * hstep; m < ns; m += SEL_THREADS * hstep) {
        const uint32_t o = ord_at(m);
        if (o) atomicAdd(&S.hist[(omax - o) >> shift], 1);
    }
    __syncthreads();
    int ks;  // rank of the cut among the histogram's samples
    if (staged) {
        ks = (sg.k * 13 + 79) / 80 + 24;  // 1.3 k / 8 + slack
    } else {
        const float want = (float)sg.k + 4.0f * sqrtf((float)sg.k * (float)stride);
        ks = (int)(want / (float)stride) + 3;
    }
    {
        const int h0 = S.hist[2 * tid], h1 = S.hist[2 * tid + 1], s = h0 + h1;
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) S.wsum[warp] = incl;
        if (tid == 0) S.kb = TS_BINS - 1;
        __syncthreads();
        if (warp == 0) {
            int w = S.wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            S.wsum[lane] = wi - w;
        }
        __syncthreads();
        const int excl = S.wsum[warp] + incl - s;
        if (excl < ks && ks <= excl + h0) S.kb = 2 * tid;
        else if (excl + h0 < ks && ks <= excl + s) S.kb = 2 * tid + 1;
        __syncthreads();
    }
    const int pb = S.kb;
    const uint32_t cut = (uint32_t)((((u64)pb + 1ull) << shift) - 1ull);  // keep omax - o <= cut
    const uint32_t cut_ord = cut >= omax ? 1u : omax - cut;               // i.e. o >= cut_ord (ord 0 = padding)
    // stash: slots of the surviving logits (unordered)
    uint32_t* slots = reinterpret_cast<uint32_t*>(tmp);
    if (staged) {
        static_assert(SEL_THREADS * 32 >= 32768, "one survivor bit per staged slot of a thread");
        unsigned em = 0u;
#pragma unroll 1
        for (int m = tid, j = 0; m < M; m += SEL_THREADS, ++j) {
            const uint32_t o = ord_at(m);
            em |= ((o && o >= cut_ord) ? 1u : 0u) << j;
            if (!P.nhwc) rank[m] = RANK_INVALID;  // (the rank map is cleared on the way)
        }
        const int c = __popc(em);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int sp = 0;
        if (lane == 31 && incl) sp = atomicAdd(&S.count, incl);
        sp = __shfl_sync(0xffffffffu, sp, 31) + incl - c;
#pragma unroll 1
        while (em) {
            const int j = __ffs(em) - 1;
            em &= em - 1;
            if (sp < 2 * kcap) slots[sp] = (uint32_t)(tid + j * SEL_THREADS);
            ++sp;
        }
    } else {
        // one pass over the whole segment in global memory: 16-byte loads (4 x 4 logits in flight per thread) where
        // the planes are 16-byte aligned, 8 scalar loads in flight otherwise. Survivors are rare (~1.4 k of the
        // segment): each takes its stash slot with one shared-memory atomic.
#pragma unroll 1
        for (int li = 0; li < sg.num_levels; ++li) {
            const LevelDev& lv = P.lv[sg.first_level + li];
            const int s0 = lv.m_off - sg.m_begin;
            const bool vec = !P.nhwc && ((lv.HW & 3) == 0) && ((reinterpret_cast<uintptr_t>(lv.ptr) & 15) == 0);
#pragma unroll 1
            for (int a = 0; a < P.A; ++a) {
                if (vec) {
                    constexpr int U = 4;
                    const float4* plane4 = reinterpret_cast<const float4*>(obj_addr(P, lv, b, a, 0));
                    uint4* rank4 = reinterpret_cast<uint4*>(rank + s0 + a * lv.HW);  // (m_begin, m_off, HW: multiples of 4)
                    const int n4 = lv.HW >> 2;
#pragma unroll 1
                    for (int i0 = 0; i0 < n4; i0 += SEL_THREADS * U) {
                        float4 v[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int i = i0 + u * SEL_THREADS + tid;
                            v[u] = i < n4 ? __ldg(plane4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int i = i0 + u * SEL_THREADS + tid;
                            if (i < n4) {
                                rank4[i] = make_uint4(RANK_INVALID, RANK_INVALID, RANK_INVALID, RANK_INVALID);
                                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    if (f2ord(e[q]) >= cut_ord) {
                                        const int sp = atomicAdd(&S.count, 1);
                                        if (sp < 2 * kcap) slots[sp] = (uint32_t)(s0 + a * lv.HW + 4 * i + q);
                                    }
                            }
                        }
                    }
                } else {
                    constexpr int U = 8;
#pragma unroll 1
                    for (int h0 = 0; h0 < lv.HW; h0 += SEL_THREADS * U) {
                        float v[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int hw = h0 + u * SEL_THREADS + tid;
                            v[u] = hw < lv.HW ? __ldg(obj_addr(P, lv, b, a, hw)) : 0.f;
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int hw = h0 + u * SEL_THREADS + tid;
                            if (hw < lv.HW) {
                                const int m = s0 + a * lv.HW + hw;
                                if (!P.nhwc) rank[m] = RANK_INVALID;
                                if (f2ord(v[u]) >= cut_ord) {
                                    const int sp = atomicAdd(&S.count, 1);
                                    if (sp < 2 * kcap) slots[sp] = (uint32_t)m;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    const int n_stash = S.count;
    YPP_PHASE(0, b, 2);
    if (n_stash < sg.k || n_stash > kcap) return false;
    // keys of the stash + their exact range
    uint32_t hmin = 0xFFFFFFFFu, hmax = 0u;
#pragma unroll 1
    for (int i = tid; i < n_stash; i += SEL_THREADS) {
        const int m = (int)slots[i];
        int l = 0, a = 0, hw = 0;
        seg_slot(P, sg, m + sg.m_begin, l, a, hw);
        const uint32_t o = staged ? ord_at(m) : f2ord(__ldg(obj_addr(P, P.lv[l], b, a, hw)));
        const float conf = c_sigmoid(ord2f(o));
        const uint32_t h = ~f2ord(conf);
        sel[i] = ((u64)h << 32) | (u64)(uint32_t)(P.lv[l].n_off + hw * P.A + a);
        hmin = h < hmin ? h : hmin;
        hmax = h > hmax ? h : hmax;
    }
    hmin = __reduce_min_sync(0xffffffffu, hmin);
    hmax = __reduce_max_sync(0xffffffffu, hmax);
    if (lane == 0) {
        red[warp] = hmin;
        red[32 + warp] = hmax;
    }
    __syncthreads();
    hmin = __reduce_min_sync(0xffffffffu, red[lane]);
    hmax = __reduce_max_sync(0xffffffffu, red[32 + lane]);
    __syncthreads();  // red[] and the slots in `tmp` are reused below
    YPP_PHASE(0, b, 8);
    StashSource none;
    const int cnt = select_sorted_prefix(none, (u64)hmin << 32, ((u64)hmax << 32) | 0xFFFFFFFFull, sg.k, sel, tmp, kcap, S, n_stash);
    YPP_PHASE(0, b, 3);
    if (cnt < sg.k) return false;
    // every logit outside the stash is below the cut: its confidence is at most conf(cut) (+ rounding noise)
    const uint32_t conf_cut = f2ord(c_sigmoid(ord2f(cut_ord)));
    const uint32_t conf_k = ~(uint32_t)(sel[sg.k - 1] >> 32);
    const bool excluded = n_stash < sg.N;  // (the range is the sample's: logits below it are outside the stash whatever the cut)
    if (excluded && !(conf_k >= conf_cut + 16u)) return false;
    select_write(P, sg, b, sel, sg.k);
    return true;
}

// One CTA per (top-k segment, image). Keys are (~ord(conf) << 32 | n): ascending = (conf desc, anchor asc) —
// the canonical order of conf_pred.topk(nms_pre) (yolocsp_head.py:350-355 / yolo_head.py:281-302).
__global__ void __launch_bounds__(SEL_THREADS, 1) select_kernel(const __grid_constant__ DevParams P) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    __shared__ TopSelSmem S;
    u64* sel = reinterpret_cast<u64*>(sel_smem);
    const int b = blockIdx.y;
    const SegDev& sg = P.seg[P.topk_segs[blockIdx.x]];
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        *P.tile_ctr = P.dec_first;  // the decode kernel's tile scheduler
        if (P.o_status) *P.o_status = 0;  // first kernel of the call: the data-dependent status starts clean
    }
    u64* ckey = P.ckey + (size_t)b * P.M_pad;
    uint32_t* rank = P.rank + (size_t)b * P.M_pad;

    YPP_PHASE(0, b, 0);
#ifdef YPP_PROFILE
    if (threadIdx.x == 0) { S.prof_kernel = 0; S.prof_call = 0; }
    __syncthreads();
#endif
    const int seg_m = sg.m_end - sg.m_begin, seg_stride = sel_stride_of(P, seg_m);
    if (P.sel_stage > 0 && sg.k <= SEL_MAX_K && (seg_m + seg_stride - 1) / seg_stride <= P.sel_stage) {
        // fast path on the raw logits; falls through to the exact path when it cannot prove its result
        if (select_fast(P, sg, b, sel, sel + P.sel_kcap, reinterpret_cast<uint32_t*>(sel + 2 * P.sel_kcap), P.sel_kcap, S)) {
            YPP_PHASE(0, b, 4);
            return;
        }
        __syncthreads();
    }
    // pass 0: objectness of every anchor of the segment -> composite key; rank map cleared; key range.
    // Loads are issued in batches of 8 per thread so that one DRAM round trip covers 8 anchors.
    u64 kmin = ~0ull, kmax = 0ull;
#pragma unroll 1
    for (int li = 0; li < sg.num_levels; ++li) {
        const LevelDev& lv = P.lv[sg.first_level + li];
        constexpr int U = 8;
#pragma unroll 1
        for (int a = 0; a < P.A; ++a) {
            const int m0 = lv.m_off + a * lv.HW;
#pragma unroll 1
            for (int h0 = 0; h0 < lv.HW; h0 += SEL_THREADS * U) {
                float v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int hw = h0 + u * SEL_THREADS + tid;
                    v[u] = hw < lv.HW ? __ldg(obj_addr(P, lv, b, a, hw)) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int hw = h0 + u * SEL_THREADS + tid;
                    if (hw < lv.HW) {
                        const float conf = c_sigmoid(v[u]);
                        const u64 key = ((u64)(~f2ord(conf)) << 32) | (u64)(uint32_t)(lv.n_off + hw * P.A + a);
                        ckey[m0 + hw] = key;
                        if (!P.nhwc) rank[m0 + hw] = RANK_INVALID;
                        kmin = key < kmin ? key : kmin;
                        kmax = key > kmax ? key : kmax;
                    }
                }
            }
        }
        // alignment padding between levels never matches a key
        const int pad0 = lv.m_off + P.A * lv.HW, pad1 = (pad0 + 3) & ~3;
        if (tid < pad1 - pad0) ckey[pad0 + tid] = ~0ull;
    }
    u64 gmin, gmax;
    block_minmax(kmin, kmax, gmin, gmax, S);  // also orders the global writes above (block barrier)
    YPP_PHASE(0, b, 1);

    // keys are fetched two at a time (16-byte loads); m_begin is 4-aligned, an odd tail reads one padding key (~0)
    CkeySource src;
    src.p = reinterpret_cast<const ulonglong2*>(ckey + sg.m_begin);
    src.n = (sg.m_end - sg.m_begin + 1) / 2;
    if (sg.k > SEL_MAX_K) {
        // nms_pre beyond the shared-memory sort capacity: the sorted prefix is produced in chunks, each one a
        // select over the keys above the previous chunk's last key (keys are unique)
        const int chunk = P.sel_kcap / 2;
        u64 lo = gmin;
        int done = 0;
#pragma unroll 1
        while (done < sg.k) {
            const int want = min(chunk, sg.k - done);
            const int cnt = select_sorted_prefix(src, lo, gmax, want, sel, sel + P.sel_kcap, P.sel_kcap, S);
            const int take = min(cnt, sg.k - done);
            if (take <= 0) break;
            select_write(P, sg, b, sel, take, done);
            done += take;
            const u64 last = sel[take - 1];
            __syncthreads();  // sel is rewritten by the next chunk
            if (cnt < want || last == ~0ull) break;
            lo = last + 1ull;
        }
        YPP_PHASE(0, b, 4);
        return;
    }
    u64 hi_sel = gmax;
    // Pivot from a 1-in-16 sample: the objectness distribution is extremely skewed (most anchors share a few
    // histogram buckets), so the exact select only looks at keys up to ~1.5x the expected k-th key. If the
    // sample pivot turns out too tight (cnt < k) the select is redone over the whole range.
    {
        CkeySampleSource ss;
        ss.p = ckey + sg.m_begin;
        ss.stride = 16;
        ss.n = (sg.m_end - sg.m_begin) / 16;
        const int ks = (sg.k * 3 + 31) / 32 + 16;  // 1.5 * k / 16 + slack
        if (ss.n >= 4 * ks && ks <= P.sel_kcap) {
            const int gs = select_sorted_prefix(ss, gmin, gmax, ks, sel, sel + P.sel_kcap, P.sel_kcap, S);
            if (gs >= ks) hi_sel = sel[ks - 1];
            __syncthreads();
        }
    }
    YPP_PHASE(0, b, 2);
    int cnt = select_sorted_prefix(src, gmin, hi_sel, sg.k, sel, sel + P.sel_kcap, P.sel_kcap, S);
    YPP_PHASE(0, b, 3);
    if (cnt < sg.k && hi_sel != gmax) {
        __syncthreads();
        cnt = select_sorted_prefix(src, gmin, gmax, sg.k, sel, sel + P.sel_kcap, P.sel_kcap, S);
    }
    select_write(P, sg, b, sel, cnt < sg.k ? cnt : sg.k);
    YPP_PHASE(0, b, 4);
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

// (x, y) of position hw on a W-wide map: float estimate of hw / W, corrected to the exact quotient
__device__ __forceinline__ void pos_xy(const LevelDev& lv, int hw, int& x, int& y) {
    int q = __float2int_rz(__fmul_rn((float)hw, lv.inv_w));
    int rem = hw - q * lv.W;
    if (rem < 0) {
        --q;
        rem += lv.W;
    } else if (rem >= lv.W) {
        ++q;
        rem -= lv.W;
    }
    x = rem;
    y = q;
}

// One admitted anchor, processed by a whole warp: lanes 0..4 activate the box / objectness logits (`attr_v` is
// the raw logit of attribute `lane` there), then the lanes sweep the classes: `cls(u)` returns the raw logit of
// class u*32 + lane (registers, shared-memory tile or a global gather). Writes row r of the score matrix, the
// decoded box, and the per-image statistics.
template <int MODE, class ClsLoad>
__device__ __forceinline__ void process_anchor(const DevParams& P, const LevelDev& lv, const SegDev& sg, int b, int a,
                                               int hw, uint32_t r, int lane, float attr_v, ClsLoad cls) {
    uint32_t* mrow = P.mat + ((size_t)b * P.R + r) * P.C;
    float act = 0.f;
    if (lane < 5) act = (MODE == 0 || lane < 2 || lane == 4) ? c_sigmoid(attr_v) : c_expf(attr_v);
    const float a0 = __shfl_sync(0xffffffffu, act, 0), a1 = __shfl_sync(0xffffffffu, act, 1);
    const float a2 = __shfl_sync(0xffffffffu, act, 2), a3 = __shfl_sync(0xffffffffu, act, 3);
    const float conf = __shfl_sync(0xffffffffu, act, 4);
    if (MODE == 1 && P.conf_thr > 0.f && !(conf >= P.conf_thr)) {  // yolo_head.py:365-376: row dropped
        for (int c = lane; c < P.C; c += 32) mrow[c] = SCORE_NONE;
        if (lane == 0) P.row_stat[(size_t)b * P.R + r] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    int x, y;
    pos_xy(lv, hw, x, y);
    float4 bx = decode_box<MODE>(lv, a, x, y, a0, a1, a2, a3);
    if (P.rescale) bx = rescale_box(bx, P.scale + 4 * b);
    if (lane == 0) {
        P.row_box[(size_t)b * P.R + r] = bx;
        if (!sg.has_topk) P.row_anchor[(size_t)b * P.R + r] = lv.n_off + hw * P.A + a;
    }
    uint32_t best = 0u, worst = 0u;  // max of ord(score) / ~ord(score) over this lane's candidates
    int npass = 0;
    if (P.agnostic) {
        // cls_pred = conf_pred[:, None]  (yolocsp_head.py:360): one class, score = objectness
        const bool pass = conf > P.score_thr;
        if (lane == 0) {
            mrow[0] = pass ? __float_as_uint(conf) : SCORE_NONE;
            if (pass) {
                best = f2ord(conf);
                worst = ~best;
                npass = 1;
            }
        }
    } else {
        for (int c0 = 0, u = 0; c0 < P.C; c0 += 32, ++u) {
            const int c = c0 + lane;
            if (c < P.C) {
                const float sgm = c_sigmoid(cls(u));
                float score;
                bool pass;
                if (MODE == 0) {
                    score = fmul(sgm, conf);     // cls_pred *= conf_pred[:, None]   (yolocsp_head.py:358)
                    pass = score > P.score_thr;  // bbox_nms.py:54
                } else {
                    pass = sgm > P.score_thr;  // threshold on the class score alone (bbox_nms.py:54) ...
                    score = fmul(sgm, conf);   // ... then scores * score_factors     (bbox_nms.py:57-62)
                }
                mrow[c] = pass ? __float_as_uint(score) : SCORE_NONE;
                if (pass) {
                    const uint32_t o = f2ord(score);
                    best = o > best ? o : best;
                    worst = ~o > worst ? ~o : worst;
                    ++npass;
                }
            }
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    worst = __reduce_max_sync(0xffffffffu, worst);
    npass = __reduce_add_sync(0xffffffffu, npass);
    if (lane == 0) P.row_stat[(size_t)b * P.R + r] = make_uint4(best, worst, (uint32_t)npass, 0u);
}

// Up to DEC_BATCH admitted anchors of one tile whose logits already sit in registers. Lane group g = lane / 8
// owns anchor slot g for the box / objectness part (attribute k = lane % 8 < 5): ONE activation pass and ONE box
// decode serve all four anchors; the class sweeps then run per anchor with all 32 lanes.
//   av      this lane's raw logit of attribute k of anchor slot g
//   tv[q][u] raw logit of class u*32 + lane of anchor slot q
template <int MODE, int OFF>
__device__ __forceinline__ void process_batch(const DevParams& P, const LevelDev& lv, const SegDev& sg, int b, int a,
                                              int hw0, int nb_all, const int (&pos_all)[DEC_BATCH],
                                              const uint32_t (&rr_all)[DEC_BATCH], float av,
                                              const float (&tv_all)[DEC_BATCH][DEC_ROUNDS], int lane) {
    // this pass serves anchor slots OFF .. OFF+3
    const int nb = nb_all - OFF;
    const int pos[4] = {pos_all[OFF], pos_all[OFF + 1], pos_all[OFF + 2], pos_all[OFF + 3]};
    const uint32_t rr[4] = {rr_all[OFF], rr_all[OFF + 1], rr_all[OFF + 2], rr_all[OFF + 3]};
    const int g = lane >> 3, k = lane & 7;
    float act = 0.f;
    if (k < 5 && g < nb) act = (MODE == 0 || k < 2 || k == 4) ? c_sigmoid(av) : c_expf(av);
    const int gl = lane & ~7;
    const float a0 = __shfl_sync(0xffffffffu, act, gl), a1 = __shfl_sync(0xffffffffu, act, gl + 1);
    const float a2 = __shfl_sync(0xffffffffu, act, gl + 2), a3 = __shfl_sync(0xffffffffu, act, gl + 3);
    const float conf_g = __shfl_sync(0xffffffffu, act, gl + 4);
    const int pg = g == 0 ? pos[0] : (g == 1 ? pos[1] : (g == 2 ? pos[2] : pos[3]));
    const uint32_t rg = g == 0 ? rr[0] : (g == 1 ? rr[1] : (g == 2 ? rr[2] : rr[3]));
    const bool drop_g = (MODE == 1) && P.conf_thr > 0.f && !(conf_g >= P.conf_thr);  // yolo_head.py:365-376
    if (g < nb && !drop_g && k == 0) {
        const int hw = hw0 + pg;
        int x, y;
        pos_xy(lv, hw, x, y);
        float4 bx = decode_box<MODE>(lv, a, x, y, a0, a1, a2, a3);
        if (P.rescale) bx = rescale_box(bx, P.scale + 4 * b);
        P.row_box[(size_t)b * P.R + rg] = bx;
        if (!sg.has_topk) P.row_anchor[(size_t)b * P.R + rg] = lv.n_off + hw * P.A + a;
    }
    // class activations of all four slots, four at a time behind one call (c_sigmoid4): four independent
    // dependency chains keep the FMA pipe busy and the consumer loop stays small. Empty slots compute on zeros.
    float sgm[4][DEC_ROUNDS];
    if (!P.agnostic) {
#pragma unroll
        for (int u = 0; u < DEC_ROUNDS; ++u) {
            const float4 r = c_sigmoid4(make_float4(tv_all[OFF][u], tv_all[OFF + 1][u], tv_all[OFF + 2][u], tv_all[OFF + 3][u]));
            sgm[0][u] = r.x;
            sgm[1][u] = r.y;
            sgm[2][u] = r.z;
            sgm[3][u] = r.w;
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (q < nb) {
            const float conf = __shfl_sync(0xffffffffu, conf_g, q * 8);
            const bool drop = __shfl_sync(0xffffffffu, (int)drop_g, q * 8) != 0;
            uint32_t* mrow = P.mat + ((size_t)b * P.R + rr[q]) * P.C;
            uint32_t best = 0u, worst = 0u;
            int npass = 0;
            if (drop) {
                for (int c = lane; c < P.C; c += 32) mrow[c] = SCORE_NONE;
            } else if (P.agnostic) {
                // cls_pred = conf_pred[:, None]  (yolocsp_head.py:360): one class, score = objectness
                const bool pass = conf > P.score_thr;
                if (lane == 0) {
                    mrow[0] = pass ? __float_as_uint(conf) : SCORE_NONE;
                    if (pass) {
                        best = f2ord(conf);
                        worst = ~best;
                        npass = 1;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < DEC_ROUNDS; ++u) {
                    const int c = u * 32 + lane;
                    const float sv = sgm[q][u];
                    float score;
                    bool pass;
                    if (MODE == 0) {
                        score = fmul(sv, conf);      // cls_pred *= conf_pred[:, None]   (yolocsp_head.py:358)
                        pass = score > P.score_thr;  // bbox_nms.py:54
                    } else {
                        pass = sv > P.score_thr;   // threshold on the class score alone (bbox_nms.py:54) ...
                        score = fmul(sv, conf);    // ... then scores * score_factors     (bbox_nms.py:57-62)
                    }
                    pass = pass && (c < P.C);
                    if (c < P.C) mrow[c] = pass ? __float_as_uint(score) : SCORE_NONE;
                    const uint32_t o = f2ord(score);
                    best = (pass && o > best) ? o : best;
                    worst = (pass && ~o > worst) ? ~o : worst;
                    npass += pass ? 1 : 0;
                }
            }
            best = __reduce_max_sync(0xffffffffu, best);
            worst = __reduce_max_sync(0xffffffffu, worst);
            npass = __reduce_add_sync(0xffffffffu, npass);
            if (lane == 0) P.row_stat[(size_t)b * P.R + rr[q]] = make_uint4(best, worst, (uint32_t)npass, 0u);
        }
    }
}

// Shared-memory geometry of one pipeline stage (all offsets multiples of 1024 so that the 128-byte TMA swizzle
// pattern is a function of the offset alone).
struct StageGeom {
    uint32_t sub_bytes;    // one (NA x 32) box, padded to 1024
    uint32_t quad_rows;    // quad-row tiles: rows of four planes a box spans
    uint32_t quad_box;     // one (quad_rows x QUAD_W) box, padded to 1024
    uint32_t rank_off;     // rank row: 64 entries + 4 (aligned superset when the row is not 16-byte aligned)
    uint32_t desc_off;     // tile descriptor written by the producer
    uint32_t stage_bytes;
};
constexpr uint32_t RANK_ROW_BYTES = (64 + 4) * 4u;  // (64-position tiles; quad-row tiles fetch the 16-byte aligned superset)
// A TMA box must START on a 16-byte boundary of the innermost dimension (an unaligned start coordinate raises an
// illegal-instruction fault on sm_100), so a quad-row box fetches the aligned superset of its 64 positions: 68
// floats per row, not swizzled (the 128-byte swizzle would cap the row at 32 floats).
constexpr int QUAD_W = 64 + 4;  // (quad-row tiles exist with 64-position tiles only)
constexpr uint32_t QUAD_ROW_BYTES = QUAD_W * 4u;
__host__ __device__ inline StageGeom stage_geom(int NA, int quad, int tile_t) {
    StageGeom g;
    g.sub_bytes = ((uint32_t)NA * 128u + 1023u) & ~1023u;
    // NA consecutive planes starting at plane p0 live in rows p0/4 .. (p0 + NA - 1)/4 of the four-plane view
    g.quad_rows = (((uint32_t)NA + 2u) >> 2) + 1u;
    g.quad_box = (g.quad_rows * QUAD_ROW_BYTES + 1023u) & ~1023u;
    uint32_t data = (uint32_t)(tile_t / TILE_SUB) * g.sub_bytes;
    if (quad && 4u * g.quad_box > data) data = 4u * g.quad_box;
    g.rank_off = data;
    g.desc_off = g.rank_off + ((RANK_ROW_BYTES + 31u) & ~31u);
    g.stage_bytes = (g.desc_off + 32u + 1023u) & ~1023u;
#ifndef YPP_NO_PACK
    // the rank row and the descriptor fit into the padding behind the last box (NA * 128 bytes are written, the box
    // area is rounded up to 1024): no extra kilobyte per stage
    const uint32_t used = data - g.sub_bytes + (((uint32_t)NA * 128u + 31u) & ~31u);
    if (!quad && used + ((RANK_ROW_BYTES + 31u) & ~31u) + 32u <= data) {
        g.rank_off = used;
        g.desc_off = g.rank_off + ((RANK_ROW_BYTES + 31u) & ~31u);
        g.stage_bytes = data;
    }
#endif
    return g;
}
// logit of attribute k at position p of the tile (SWIZZLE_128B: 16-byte chunk index XOR (row mod 8))
__device__ __forceinline__ uint32_t tile_off(uint32_t sub_bytes, int k, int p) {
    const int col = p & 31;
    return (uint32_t)(p >> 5) * sub_bytes + (uint32_t)k * 128u +
           ((uint32_t)(((col >> 2) ^ (k & 7)) << 4) | (uint32_t)((col & 3) << 2));
}
__device__ __forceinline__ float tile_at(const unsigned char* stage, uint32_t sub_bytes, int k, int p) {
    return *reinterpret_cast<const float*>(stage + tile_off(sub_bytes, k, p));
}
// The same for a quad-row tile: the level is viewed as rows of four planes (row stride 16 * HW bytes); the tile
// arrives as 4 boxes, box j = plane-in-row j, rows r0 .. r0 + quad_rows - 1, QUAD_W floats from the 16-byte aligned
// position at or below j*HW + hw0 (the first `sh` floats of a row are the alignment slack). Attribute k of the slab
// is plane p0 + k, i.e. row (pl0 + k) / 4 - relative to r0 - and plane-in-row (pl0 + k) % 4, pl0 = p0 % 4.
__device__ __forceinline__ float quad_at(const unsigned char* stage, uint32_t quad_box, int pl0, int hwn, int hw0, int k, int p) {
    const int pk = pl0 + k, j = pk & 3, r = pk >> 2;
    const int sh = (j * hwn + hw0) & 3;
    const uint32_t off = (uint32_t)j * quad_box + (uint32_t)r * QUAD_ROW_BYTES + (uint32_t)(sh + p) * 4u;
    return *reinterpret_cast<const float*>(stage + off);
}

#ifdef YPP_QUAD
// One quad-row tile, consumed in place by one warp: logits are read from the four boxes, one admitted anchor at a
// time (4.8 % of the bytes at 608^2). `z` = descriptor word: bit 0 top-k, bits 4..5 misalignment of the rank row,
// bits 8..9 first plane of the slab within its row of four.
template <int MODE>
__device__ __noinline__ void consume_quad_tile(const DevParams& P, const unsigned char* stage, uint32_t quad_box, const uint32_t* rk,
                                               int lvl, int b, int a, int hw0, int HWn, int z, int lane) {
    const int shift = (z >> 4) & 3, pl0 = (z >> 8) & 3;
    const LevelDev& lv = P.lv[lvl];
    const SegDev& sg = P.seg[lv.seg];
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        const int ps0 = h * 32 + lane;
        uint32_t r = RANK_INVALID;
        if (hw0 + ps0 < HWn) r = rk[shift + ps0];
        unsigned adm = __ballot_sync(0xffffffffu, r != RANK_INVALID);
        while (adm) {
            const int src = __ffs(adm) - 1;
            adm &= adm - 1;
            const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
            const int ps = h * 32 + src;
            const float av = lane < 5 ? quad_at(stage, quad_box, pl0, HWn, hw0, lane, ps) : 0.f;
            process_anchor<MODE>(P, lv, sg, b, a, hw0 + ps, rr, lane, av, [&](int u) -> float {
                return quad_at(stage, quad_box, pl0, HWn, hw0, 5 + u * 32 + lane, ps);
            });
        }
    }
}
#endif

// Persistent, warp-specialised. Two producer warps: a warp's 32 lanes work out the coordinates of its next 32 tiles
// in parallel (the level lookup and the integer divisions are the long latency chain of a tile), then lane 0 streams
// each (NA x TILE_T) tile of the raw head tensor into the shared-memory ring — 8 stages of one 128B-swizzled TMA box
// (TILE_T = 32) or 4 stages of two (TILE_T = 64), no L2 eviction hint — plus the tile's rank-map row (1-D bulk copy)
// and a 32-byte tile descriptor. 8 consumer warps claim whole tiles in order: a consumer pulls the logits of up to
// four admitted anchors into registers, hands the stage back to the producer before the last batch's math, and only
// then does the math (lanes over classes), so a stage is held for a few hundred cycles unless the tile is heavy.
#ifdef YPP_PROFILE
// per-tile timestamps (profiling build only): [tile][0..5] = issue start, issued, landed, released, done, smid
__device__ long long g_prof[(1 << 16) * 8];
// per CTA: [0] clock64 and [1] globaltimer (ns) at kernel entry, [2]/[3] the same when the CTA's last consumer warp
// leaves, [4] smid — lets the per-SM clock64 stamps be placed on one time axis
__device__ long long g_prof_cta[1024 * 8];
__device__ __forceinline__ long long ypp_globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define YPP_STAMP(t, i) do { if (lane == 0 && (t) < (1 << 16)) g_prof[(size_t)(t) * 8 + (i)] = clock64(); } while (0)
#else
#define YPP_STAMP(t, i) do { } while (0)
#endif

// Handing a stage back is a write-after-read hazard ACROSS PROXIES: the consumer read the stage through the generic
// proxy (ld.shared), the refill writes it through the async proxy (TMA). The mbarrier orders the two threads but
// not the two proxies, so every consumer lane issues a proxy fence before the warp's arrive, and the producer
// issues one after it has acquired the stage. Without them the ring gave run-to-run differences in the first
// tiles of a launch (3 % of the runs at batch 64, 35 % at batch 128; tools/race_probe.py); with either fence 0 of
// 400 runs differed, at no measurable cost.
__device__ __forceinline__ void stage_release_fence() { fence_proxy_async(); }

constexpr int DEC_KIND_STOP = 7;  // stage header: the producer has run out of tiles (tile kinds are 1, 2, 3)

template <int MODE, int TILE_T>
__global__ void __launch_bounds__(DEC_THREADS, 2) decode_tma_kernel(const __grid_constant__ DevParams P,
                                                                  const __grid_constant__ TmapPack maps) {
    constexpr int DEC_STAGES = dec_stages_of(TILE_T);
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte aligned base (swizzle) — the launch reserves 1 KB of slack
    unsigned char* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int NA = P.NA;
    const StageGeom g = stage_geom(NA, P.dec_quad, TILE_T);
    const uint32_t box_bytes = (uint32_t)NA * 128u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);  // [DEC_STAGES]: the tile streamed into stage s has landed
    uint64_t* empty = full + DEC_STAGES;                      // [DEC_STAGES]: stage s may be refilled
    int* next_it = reinterpret_cast<int*>(empty + DEC_STAGES);  // next ring iteration a consumer may claim
    int* done = next_it + 1;                                    // [DEC_PWARPS]: producer p has posted its STOP
    unsigned char* stages = smem_raw + 1024;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef YPP_PROFILE
    if (threadIdx.x == 0 && blockIdx.x < 1024) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_prof_cta[blockIdx.x * 8 + 0] = clock64();
        g_prof_cta[blockIdx.x * 8 + 1] = ypp_globaltimer();
        g_prof_cta[blockIdx.x * 8 + 4] = smid;
    }
#endif
    if (threadIdx.x == 0) {
        *next_it = 0;
        for (int pw = 0; pw < DEC_PWARPS; ++pw) done[pw] = 0;
        for (int s = 0; s < DEC_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            reinterpret_cast<int4*>(stages + (size_t)s * g.stage_bytes + g.desc_off)->w = -1;  // "no tile yet"
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int total = P.tma_tiles;
    if (warp < DEC_PWARPS) {
        // ---------------- producers: warp p fills the ring iterations it == p (mod DEC_PWARPS), i.e. stages p, p+2 ----
        // SMs do not all stream at the same rate: with a purely static round-robin of the tiles the CTAs finished
        // between 78 and 102 us (tools/prof_timeline.py) and the last 20 us ran far below the HBM rate. Hybrid
        // schedule: the first 3/4 of the tile sequence is dealt round-robin exactly as before (descriptors of 32
        // tiles computed across the lanes, no atomics, neighbouring tiles in flight at the same time), the last
        // quarter is claimed from ONE device-wide counter in small guided batches (remaining / (4 x producer warps),
        // at most 8, at least 1; the next batch is claimed before the current one is issued so that the atomic's
        // latency hides behind the waits for empty stages), so fast CTAs keep going until the sequence is empty and
        // all CTAs run dry within a tile or two of each other.
        static_assert(DEC_STAGES % DEC_PWARPS == 0, "a stage must belong to one producer warp");
        const unsigned nprod = gridDim.x * DEC_PWARPS;
        const unsigned S = P.dec_first;  // positions [0, S) are dealt round-robin, [S, total) come from the counter
        auto batch_size = [&](unsigned seen) -> unsigned {
            const unsigned rem = seen < (unsigned)total ? (unsigned)total - seen : 0u;
            const unsigned n = rem / (4u * nprod);
            return n < 1u ? 1u : (n > 8u ? 8u : n);
        };
        unsigned ks = 0;      // static phase: tiles of this warp dealt so far
        bool dyn = false;     // dynamic phase entered
        unsigned n = 0, base = 0;
        int k = 0;  // tiles this warp has put into the ring
        while (true) {
            unsigned q_lane = 0xFFFFFFFFu, n_next = 0u, base_next = 0u;
            bool to_dyn = dyn;
            if (!dyn) {
                // round-robin over the CTAs (neighbouring tiles are in flight at the same time in different CTAs):
                // this warp's ks-th tile is position blockIdx.x + (ks * DEC_PWARPS + warp) * gridDim.x
                const unsigned q0 = blockIdx.x + (ks * DEC_PWARPS + (unsigned)warp) * gridDim.x;
                if (q0 < S) {
                    const unsigned left = (S - q0 + nprod - 1u) / nprod;
                    n = left < 32u ? left : 32u;
                    q_lane = q0 + (unsigned)lane * nprod;
                    ks += n;
                    if (n == left) {  // last static batch: claim the first dynamic one now
                        n_next = batch_size(S);
                        // (everything dealt statically: nothing to claim, the STOP follows without an atomic's latency)
                        base_next = (unsigned)total;
                        if (lane == 0 && S < (unsigned)total) base_next = atomicAdd(P.tile_ctr, n_next);
                        to_dyn = true;
                    }
                } else {
                    dyn = to_dyn = true;
                    n = batch_size(S);
                    if (lane == 0) base = atomicAdd(P.tile_ctr, n);
                }
            }
            if (dyn) {
                base = __shfl_sync(0xffffffffu, base, 0);
                // the following batch, claimed now (lane 0 keeps the result to itself until the next round)
                n_next = batch_size(base + n);
                if (lane == 0 && base < (unsigned)total) base_next = atomicAdd(P.tile_ctr, n_next);
                q_lane = base + (unsigned)lane;
            }
            // lane j < n: coordinates of its sequence position
            int d_t = -1, d_l = 0, d_plane = 0, d_hw0 = 0, d_b = 0, d_a = 0, d_hw = 0, d_rbase = 0, d_topk = 0, d_kind = 0;
            if ((unsigned)lane < n && q_lane < (unsigned)total) {
                const int t = (int)q_lane;
                int best0 = -1;  // the level with the largest first-tile id <= t (gather levels are enumerated first)
                for (int i = 0; i < P.L; ++i)
                    if (P.lv[i].use_tma && t >= P.lv[i].tile0 && P.lv[i].tile0 > best0) {
                        best0 = P.lv[i].tile0;
                        d_l = i;
                    }
                const LevelDev& lv = P.lv[d_l];
                const SegDev& sg = P.seg[lv.seg];
                const int loc = t - lv.tile0;
                d_t = t;
                d_plane = loc / lv.tpp;  // b*A + a
                d_hw0 = (loc - d_plane * lv.tpp) * lv.tile_pos;
                d_b = d_plane / P.A;
                d_a = d_plane - d_b * P.A;
                d_hw = lv.HW;
                d_topk = sg.has_topk;
                d_kind = lv.use_tma;
                // quad-row view: the slab's last plane must lie inside the view (only the tensor's very last slab can
                // fall outside, when B * A * NA is not a multiple of four) — otherwise the tile is gathered
                if (d_kind == 3 && ((d_plane * NA + NA - 1) >> 2) >= lv.qrows) d_kind = 2;
                d_rbase = sg.row_off + (lv.n_off - P.lv[sg.first_level].n_off) + d_a;  // row of position 0 (no top-k)
            }
            for (unsigned j = 0; j < n; ++j, ++k) {
                const int t = __shfl_sync(0xffffffffu, d_t, j);
                const int l = __shfl_sync(0xffffffffu, d_l, j), plane = __shfl_sync(0xffffffffu, d_plane, j);
                const int hw0 = __shfl_sync(0xffffffffu, d_hw0, j), bb = __shfl_sync(0xffffffffu, d_b, j);
                const int a = __shfl_sync(0xffffffffu, d_a, j);
                const int hwn = __shfl_sync(0xffffffffu, d_hw, j), rbase = __shfl_sync(0xffffffffu, d_rbase, j);
                const int tk = __shfl_sync(0xffffffffu, d_topk, j), kind = __shfl_sync(0xffffffffu, d_kind, j);
                if (lane == 0) {
                    const int it = k * DEC_PWARPS + warp;
                    const int s = it % DEC_STAGES;
                    const uint32_t ph = (uint32_t)(it / DEC_STAGES) & 1u;
                    unsigned char* dst = stages + (size_t)s * g.stage_bytes;
                    uint64_t* fb = &full[s];
                    YPP_STAMP(t < 0 ? (1 << 16) : t, 0);
                    mbar_wait(&empty[s], ph ^ 1u);
                    if (t < 0) {
                        // out of tiles: tell the consumers (the header lives in the stage, so the stage must be free)
                        *reinterpret_cast<int4*>(dst + g.desc_off) = make_int4(DEC_KIND_STOP << 16, 0, 0, it);
                        mbar_arrive(fb);
                    } else {
                        fence_proxy_async();  // see stage_release_fence(): generic reads of the stage -> TMA writes
                        YPP_STAMP(t, 1);
                        // descriptor first (it carries the iteration number the consumer matches and everything the
                        // consumer needs before it can release the stage), then arm the barrier
                        // quad-row tiles: first plane of the slab in its row of four, misalignment of the rank row
                        const int q_pl = plane * NA, q_m = P.lv[l].m_off + a * hwn + hw0, q_shift = q_m & 3;
                        const int tkz = kind == 3 ? (tk | (q_shift << 4) | ((q_pl & 3) << 8)) : tk;
                        *reinterpret_cast<int4*>(dst + g.desc_off + 16) = make_int4(hwn, rbase, tkz, t);
                        __threadfence_block();  // release: the iteration word below is what the consumers poll (acquire)
                        *reinterpret_cast<int4*>(dst + g.desc_off) = make_int4(l | (a << 8) | (kind << 16), bb, hw0, it);
                        if (kind == 2) {
                            mbar_arrive(fb);  // gather tile: nothing to stream, the consumer reads global memory itself
#ifdef YPP_QUAD
                        } else if (kind == 3) {
                            // quad-row tile: one box per plane-in-row; neither the boxes nor the rank row start 16-byte
                            // aligned: the aligned supersets are fetched and the consumer skips the slack
                            const int r0 = q_pl >> 2;
                            mbar_arrive_expect_tx(fb, g.quad_rows * QUAD_ROW_BYTES * 4u + (tk ? RANK_ROW_BYTES : 0u));
#pragma unroll
                            for (int qj = 0; qj < 4; ++qj)
                                tma_load_2d(dst + (size_t)qj * g.quad_box, &maps.m[l], (qj * hwn + hw0) & ~3, r0, fb);
                            if (tk)
                                bulk_load_1d(dst + g.rank_off, P.rank + (size_t)bb * P.M_pad + (q_m - q_shift), RANK_ROW_BYTES, fb);
#endif
                        } else {
                            const bool topk = tk != 0;
                            const bool two = TILE_T > TILE_SUB && hw0 + TILE_SUB < hwn;  // the second box is not entirely out of bounds
                            mbar_arrive_expect_tx(fb, box_bytes * (two ? 2u : 1u) + (topk ? TILE_T * 4u : 0u));
                            // No L2 eviction hint on purpose: plane rows are in general not 128-byte aligned, so
                            // neighbouring tiles share the lines at their common edge; with evict_first the second
                            // tile re-fetched them from DRAM (measured: 635 MB read per launch instead of 512 MB).
                            tma_load_2d(dst, &maps.m[l], hw0, plane * NA, fb);
                            if (two) tma_load_2d(dst + g.sub_bytes, &maps.m[l], hw0 + TILE_SUB, plane * NA, fb);
                            if (topk)
                                bulk_load_1d(dst + g.rank_off,
                                             P.rank + (size_t)bb * P.M_pad + P.lv[l].m_off + a * hwn + hw0, TILE_T * 4u, fb);
                        }
                    }
                }
                if (t < 0) return;  // STOP posted (uniform: t comes from a shuffle)
            }
            if (to_dyn) {
                dyn = true;
                n = n_next;
                base = base_next;
            }
        }
    }
    // ---------------- consumers: whichever warp is free claims the CTA's next ring iteration, in order ------------
    // Tiles carry 0..8 admitted anchors, so a static round-robin would let one heavy tile block its stage (and
    // with it the in-order producers) while the other warps idle. Several warps therefore wait on the same
    // stage barrier for DIFFERENT fills, and a parity wait can only tell a barrier's current phase from the one
    // before it. The tile descriptor disambiguates: the producer writes the iteration number into the stage
    // header before arming the barrier, which it can only do after the previous fill of that stage was consumed;
    // a consumer first waits until the header shows ITS iteration (the barrier is then in that fill's phase or
    // later), and only then does the parity wait. A producer that has run out of tiles posts a STOP header; the
    // consumer that meets it raises the producer's `done` flag, which also frees the consumers already waiting
    // for later iterations of that producer.
    const bool in_regs = P.C <= 32 * DEC_ROUNDS;
    while (true) {
        int it = 0;
        if (lane == 0) it = atomicAdd(next_it, 1);
        it = __shfl_sync(0xffffffffu, it, 0);
        const int s = it % DEC_STAGES;
        const unsigned char* stage = stages + (size_t)s * g.stage_bytes;
        {
            // (acquire loads: what the producer wrote before the iteration word / the flag is visible afterwards, and the
            // parity wait that follows cannot be served from an older view of the barrier)
            const int* dit = reinterpret_cast<const int*>(stage + g.desc_off) + 3;
            bool gone = false;
            while (ld_acquire_cta_shared(dit) != it) {
                if (ld_acquire_cta_shared(&done[it % DEC_PWARPS])) {
                    // the STOP was posted after every real fill of this producer was armed: look once more
                    gone = ld_acquire_cta_shared(dit) != it;
                    break;
                }
                __nanosleep(YPP_SLEEP);
            }
            if (gone) {
                bool all = true;
#pragma unroll
                for (int pw = 0; pw < DEC_PWARPS; ++pw) all = all && ld_acquire_cta_shared(&done[pw]) != 0;
                if (all) break;
                continue;
            }
        }
        mbar_wait(&full[s], (uint32_t)(it / DEC_STAGES) & 1u);
        const int4 desc = *reinterpret_cast<const int4*>(stage + g.desc_off);
        const int kind = desc.x >> 16;
        if (kind == DEC_KIND_STOP) {
            if (lane == 0) {
                __threadfence_block();
                *reinterpret_cast<volatile int*>(&done[it % DEC_PWARPS]) = 1;
            }
            __syncwarp();
            bool all = true;
#pragma unroll
            for (int pw = 0; pw < DEC_PWARPS; ++pw) all = all && ld_acquire_cta_shared(&done[pw]) != 0;
            if (all) break;
            continue;
        }
        const int4 desc2 = *reinterpret_cast<const int4*>(stage + g.desc_off + 16);
        const int b = desc.y, a = (desc.x >> 8) & 0xFF, hw0 = desc.z;
        const int lvl = desc.x & 0xFF, HWn = desc2.x, rbase = desc2.y;
        const bool topk = (desc2.z & 1) != 0;
        const int tile_id = desc2.w;
        (void)tile_id;
        YPP_STAMP(tile_id, 2);
        if (kind == 2) {
            // gather tile (plane stride not 16-byte aligned, e.g. 19x19): the stage is not used
            stage_release_fence();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            const LevelDev& lv = P.lv[lvl];
            const SegDev& sg = P.seg[lv.seg];
            const float* slab = lv.ptr + (size_t)(b * P.A + a) * NA * lv.HW;
            const size_t HW = (size_t)lv.HW;
#pragma unroll 1
            for (int h = 0; h < lv.tile_pos / 32; ++h) {
                const int hw = hw0 + h * 32 + lane;
                uint32_t r = RANK_INVALID;
                if (hw < lv.HW) r = row_of(P, lv, b, a, hw);
                unsigned adm = __ballot_sync(0xffffffffu, r != RANK_INVALID);
                while (adm) {
                    const int src = __ffs(adm) - 1;
                    adm &= adm - 1;
                    const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
                    const int hwp = hw0 + h * 32 + src;
                    const float av = lane < 5 ? __ldg(slab + (size_t)lane * HW + hwp) : 0.f;
                    process_anchor<MODE>(P, lv, sg, b, a, hwp, rr, lane, av, [&](int u) -> float {
                        return __ldg(slab + (size_t)(5 + u * 32 + lane) * HW + hwp);
                    });
                }
            }
            continue;
        }
        const uint32_t* rk = reinterpret_cast<const uint32_t*>(stage + g.rank_off);
#ifdef YPP_QUAD
        if (kind == 3) {
            // quad-row tile (plane stride not 16-byte aligned, e.g. 19x19): behind a call, so that the hot loop of the
            // ordinary tiles keeps its instruction footprint (with this block inlined every tile got 10 % slower)
            consume_quad_tile<MODE>(P, stage, g.quad_box, rk, lvl, b, a, hw0, HWn, desc2.z, lane);
            stage_release_fence();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            YPP_STAMP(tile_id, 3);
            YPP_STAMP(tile_id, 4);
#ifdef YPP_PROFILE
            if (lane == 0 && tile_id < (1 << 16)) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                g_prof[(size_t)tile_id * 8 + 5] = (long long)smid | ((long long)blockIdx.x << 16);
            }
#endif
            continue;
        }
#endif  // YPP_QUAD

        // admitted positions of the tile as a 64-bit mask (everything needed comes from the stage header: no
        // parameter-space lookups before the stage is released)
        uint32_t r_lo = RANK_INVALID, r_hi = RANK_INVALID;
        {
            const int p0 = lane, p1 = lane + 32;
            if (hw0 + p0 < HWn) r_lo = topk ? rk[p0] : (uint32_t)(rbase + (hw0 + p0) * P.A);
            if (TILE_T > 32 && hw0 + p1 < HWn) r_hi = topk ? rk[p1] : (uint32_t)(rbase + (hw0 + p1) * P.A);
        }
        uint32_t m_lo = __ballot_sync(0xffffffffu, r_lo != RANK_INVALID);
        uint32_t m_hi = __ballot_sync(0xffffffffu, r_hi != RANK_INVALID);
        YPP_STAMP(tile_id, 6);
        bool released = false;
        if (!(m_lo | m_hi)) {
            stage_release_fence();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            released = true;
            YPP_STAMP(tile_id, 3);
        }
        while (m_lo | m_hi) {
            // pick up to DEC_BATCH admitted positions (uniform scalar work) ...
            int pos[DEC_BATCH];
            int nb = 0;
#pragma unroll
            for (int q = 0; q < DEC_BATCH; ++q) {
                const bool has = (m_lo | m_hi) != 0u;
                int ps = m_lo ? (__ffs(m_lo) - 1) : (32 + __ffs(m_hi) - 1);
                if (has) {
                    if (m_lo) m_lo &= m_lo - 1;
                    else m_hi &= m_hi - 1;
                    nb = q + 1;
                }
                pos[q] = has ? ps : 0;
            }
            // ... then pull their logits out of the tile in one straight-line block of independent loads
            // (empty slots read position 0: harmless, and it keeps the block free of branches)
            uint32_t rr[DEC_BATCH];
            float tv[DEC_BATCH][DEC_ROUNDS];
#pragma unroll
            for (int q = 0; q < DEC_BATCH; ++q) {
                rr[q] = __shfl_sync(0xffffffffu, pos[q] < 32 ? r_lo : r_hi, pos[q] & 31);
#pragma unroll
                for (int u = 0; u < DEC_ROUNDS; ++u) {
                    const int c = u * 32 + lane;
                    tv[q][u] = (in_regs && !P.agnostic && c < P.C) ? tile_at(stage, g.sub_bytes, 5 + c, pos[q]) : 0.f;
                }
            }
            const LevelDev& lv = P.lv[lvl];
            const SegDev& sg = P.seg[lv.seg];
            if (in_regs) {
                // box / objectness logits: lane group (lane / 8) <-> anchor slot, lane % 8 <-> attribute
                const int gq = lane >> 3, kq = lane & 7;
                const int pg0 = gq == 0 ? pos[0] : (gq == 1 ? pos[1] : (gq == 2 ? pos[2] : pos[3]));
                const float av0 = (kq < 5 && gq < nb) ? tile_at(stage, g.sub_bytes, kq, pg0) : 0.f;
                YPP_STAMP(tile_id, 7);
                // last batch and everything is in registers: give the stage back before the math
                if (!(m_lo | m_hi)) {
                    stage_release_fence();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s]);
                    released = true;
                    YPP_STAMP(tile_id, 3);
                }
                process_batch<MODE, 0>(P, lv, sg, b, a, hw0, nb, pos, rr, av0, tv, lane);
            } else {
                // wide heads (C > 96): classes are read from the tile in place, the stage is held meanwhile
#pragma unroll
                for (int q = 0; q < DEC_BATCH; ++q) {
                    if (q < nb) {
                        const int ps = pos[q];
                        const float av = lane < 5 ? tile_at(stage, g.sub_bytes, lane, ps) : 0.f;
                        process_anchor<MODE>(P, lv, sg, b, a, hw0 + ps, rr[q], lane, av, [&](int u) -> float {
                            return tile_at(stage, g.sub_bytes, 5 + u * 32 + lane, ps);
                        });
                    }
                }
            }
        }
        if (!released) {
            stage_release_fence();
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            YPP_STAMP(tile_id, 3);
        }
        YPP_STAMP(tile_id, 4);
#ifdef YPP_PROFILE
        if (lane == 0 && tile_id < (1 << 16)) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_prof[(size_t)tile_id * 8 + 5] = (long long)smid | ((long long)blockIdx.x << 16);
        }
#endif
    }
#ifdef YPP_PROFILE
    if (lane == 0 && blockIdx.x < 1024) {  // every consumer warp stamps; the last one to leave wins
        g_prof_cta[blockIdx.x * 8 + 2] = clock64();
        g_prof_cta[blockIdx.x * 8 + 3] = ypp_globaltimer();
    }
#endif
}

// Sparse admission on levels the persistent kernel cannot take (more than 256 attributes per anchor): a warp scans
// 32 positions, then gathers each admitted anchor's logits (lanes over classes).
template <int MODE>
__global__ void __launch_bounds__(128) decode_ldg_kernel(const __grid_constant__ DevParams P) {
    const int t = blockIdx.x;
    int l = 0;
    for (int q = 0; q < P.L; ++q)
        if (!P.lv[q].use_tma && !P.lv[q].dense && t >= P.lv[q].tile0) l = q;
    const LevelDev& lv = P.lv[l];
    const int loc = t - lv.tile0;
    const int plane = loc / lv.tpp;
    const int ht = loc - plane * lv.tpp;
    const int b = plane / P.A, a = plane - b * P.A;
    const int hw = ht * 128 + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const SegDev& sg = P.seg[lv.seg];
    const float* slab = lv.ptr + (size_t)plane * P.NA * lv.HW;
    const size_t HW = (size_t)lv.HW;

    uint32_t r = RANK_INVALID;
    if (hw < lv.HW) r = row_of(P, lv, b, a, hw);
    unsigned adm = __ballot_sync(0xffffffffu, r != RANK_INVALID);
    const int hw_w = hw - lane;  // first position of this warp
    while (adm) {
        const int src = __ffs(adm) - 1;
        adm &= adm - 1;
        const uint32_t rr = __shfl_sync(0xffffffffu, r, src);
        const int hwp = hw_w + src;
        const float av = lane < 5 ? __ldg(slab + (size_t)lane * HW + hwp) : 0.f;
        process_anchor<MODE>(P, lv, sg, b, a, hwp, rr, lane, av,
                             [&](int u) -> float { return __ldg(slab + (size_t)(5 + u * 32 + lane) * HW + hwp); });
    }
}

// Channels-last head tensors (P.nhwc: memory order (B, H, W, A*NA), what a cuDNN NHWC convolution writes): the
// 5+C logits of one anchor are CONTIGUOUS, so the decode is driven by the rows instead of streaming the tensor —
// one warp per (image, row): anchor index from the top-k's row table (or arithmetic when the segment keeps every
// anchor), one coalesced 4*NA-byte read, the same process_anchor as everywhere else. HBM traffic: only the
// admitted anchors' logits (1000 x 340 B per image at 608^2 instead of 7.7 MB).
constexpr int ROWS_WARPS = 8;
template <int MODE>
__global__ void __launch_bounds__(32 * ROWS_WARPS) decode_rows_kernel(const __grid_constant__ DevParams P) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * ROWS_WARPS + warp;
    if (w >= (long long)P.B * P.R) return;
    const int b = (int)(w / P.R), r = (int)(w - (long long)b * P.R);
    int s = 0;
    for (int q = 1; q < P.nsegs; ++q)
        if (r >= P.seg[q].row_off) s = q;
    const SegDev& sg = P.seg[s];
    int n;  // concatenated anchor index of the row
    if (sg.has_topk) n = P.row_anchor[(size_t)b * P.R + r];
    else n = P.lv[sg.first_level].n_off + (r - sg.row_off);
    int l = sg.first_level;
    for (int q = sg.num_levels - 1; q >= 0; --q)
        if (n >= P.lv[sg.first_level + q].n_off) {
            l = sg.first_level + q;
            break;
        }
    const LevelDev& lv = P.lv[l];
    const int loc = n - lv.n_off;
    const int hw = loc / P.A, a = loc - hw * P.A;
    const float* src = lv.ptr + (((size_t)b * lv.HW + hw) * P.A + a) * P.NA;
    const float av = lane < 5 ? __ldg(src + lane) : 0.f;
    process_anchor<MODE>(P, lv, sg, b, a, hw, (uint32_t)r, lane, av, [&](int u) -> float { return __ldg(src + 5 + u * 32 + lane); });
}

// Dense admission (no top-k, or a top-k that keeps most of the level's anchors — e.g. YOLOv3's 20x20 level with
// nms_pre = 1000 of 1200): every position is computed, one thread per position, loads coalesced across the warp's
// 32 consecutive positions of a plane. Scores are transposed through shared memory in chunks of DENSE_CH classes
// so that the rows of the score matrix leave as coalesced 4*DENSE_CH-byte runs (a thread owning a whole row would
// write 32 scattered words per instruction). One warp = one tile of 32 positions; 4 independent warps per block.
constexpr int DENSE_CH = 32;
constexpr int DENSE_WARPS = 4;
template <int MODE>
__global__ void __launch_bounds__(32 * DENSE_WARPS, 6) decode_dense_kernel(const __grid_constant__ DevParams P) {
    __shared__ uint32_t sm_all[DENSE_WARPS][32][DENSE_CH + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * DENSE_WARPS + warp;
    if (t >= P.dense_tiles) return;
    uint32_t (*sm)[DENSE_CH + 1] = sm_all[warp];
    int l = 0;
    for (int q = 0; q < P.L; ++q)
        if (!P.lv[q].use_tma && P.lv[q].dense && t >= P.lv[q].tile0) l = q;
    const LevelDev& lv = P.lv[l];
    const SegDev& sg = P.seg[lv.seg];
    const int loc = t - lv.tile0;
    const int plane = loc / lv.tpp;
    const int b = plane / P.A, a = plane - b * P.A;
    const int hw = (loc - plane * lv.tpp) * 32 + lane;
    const float* slab = lv.ptr + (size_t)plane * P.NA * lv.HW;
    const size_t HW = (size_t)lv.HW;
    const bool inb = hw < lv.HW;
    const int hwc = inb ? hw : lv.HW - 1;  // out-of-range lanes recompute the last position (keeps loads uniform)

    uint32_t r = RANK_INVALID;
    if (inb) r = row_of(P, lv, b, a, hw);
    bool adm = r != RANK_INVALID;
    if (!__any_sync(0xffffffffu, adm)) return;
    const float t0 = __ldg(slab + 0 * HW + hwc), t1 = __ldg(slab + 1 * HW + hwc);
    const float t2 = __ldg(slab + 2 * HW + hwc), t3 = __ldg(slab + 3 * HW + hwc);
    const float conf = c_sigmoid(__ldg(slab + 4 * HW + hwc));
    const bool drop = (MODE == 1) && P.conf_thr > 0.f && !(conf >= P.conf_thr);  // yolo_head.py:365-376
    if (adm && !drop) {
        const float a0 = c_sigmoid(t0), a1 = c_sigmoid(t1);
        const float a2 = (MODE == 0) ? c_sigmoid(t2) : c_expf(t2);
        const float a3 = (MODE == 0) ? c_sigmoid(t3) : c_expf(t3);
        int x, y;
        pos_xy(lv, hw, x, y);
        float4 bx = decode_box<MODE>(lv, a, x, y, a0, a1, a2, a3);
        if (P.rescale) bx = rescale_box(bx, P.scale + 4 * b);
        P.row_box[(size_t)b * P.R + r] = bx;
        if (!sg.has_topk) P.row_anchor[(size_t)b * P.R + r] = lv.n_off + hw * P.A + a;
    }
    uint32_t best = 0u, worst = 0u;
    int npass = 0;
    const unsigned admm = __ballot_sync(0xffffffffu, adm);
    if (P.agnostic) {
        // cls_pred = conf_pred[:, None]  (yolocsp_head.py:360): one class, score = objectness
        const bool pass = adm && conf > P.score_thr;
        if (adm) P.mat[((size_t)b * P.R + r) * P.C] = pass ? __float_as_uint(conf) : SCORE_NONE;
        if (pass) {
            best = f2ord(conf);
            worst = ~best;
            npass = 1;
        }
    } else {
        const float* cls = slab + 5 * HW + hwc;
        for (int c0 = 0; c0 < P.C; c0 += DENSE_CH) {
            const int nc = min(DENSE_CH, P.C - c0);
            // the chunk's logits first (DENSE_CH independent coalesced loads in flight per thread), then the math
            float tl[DENSE_CH];
#pragma unroll
            for (int j = 0; j < DENSE_CH; ++j) tl[j] = j < nc ? __ldg(cls + (size_t)(c0 + j) * HW) : 0.f;
#pragma unroll
            for (int j4 = 0; j4 < DENSE_CH; j4 += 4) {
                if (j4 < nc) {  // (uniform; classes beyond nc compute on zeros and are not stored)
                    const float4 sg4 = c_sigmoid4(make_float4(tl[j4], tl[j4 + 1], tl[j4 + 2], tl[j4 + 3]));
                    const float sv[4] = {sg4.x, sg4.y, sg4.z, sg4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int j = j4 + q;
                        const float sgm = sv[q];
                        float score;
                        bool pass;
                        if (MODE == 0) {
                            score = fmul(sgm, conf);     // cls_pred *= conf_pred[:, None]   (yolocsp_head.py:358)
                            pass = score > P.score_thr;  // bbox_nms.py:54
                        } else {
                            pass = sgm > P.score_thr;    // threshold on the class score alone (bbox_nms.py:54) ...
                            score = fmul(sgm, conf);     // ... then scores * score_factors    (bbox_nms.py:57-62)
                        }
                        pass = pass && !drop && j < nc;
                        if (j < nc) sm[lane][j] = pass ? __float_as_uint(score) : SCORE_NONE;
                        if (pass && adm) {
                            const uint32_t o = f2ord(score);
                            best = o > best ? o : best;
                            worst = ~o > worst ? ~o : worst;
                            ++npass;
                        }
                    }
                }
            }
            __syncwarp();
            unsigned todo = admm;
            while (todo) {
                const int q = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint32_t rq = __shfl_sync(0xffffffffu, r, q);
                uint32_t* mrow = P.mat + ((size_t)b * P.R + rq) * P.C + c0;
                for (int j = lane; j < nc; j += 32) mrow[j] = sm[q][j];
            }
            __syncwarp();
        }
    }
    if (adm) P.row_stat[(size_t)b * P.R + r] = make_uint4(best, worst, (uint32_t)npass, 0u);
}

// Candidate scan of the NMS kernel through shared memory: the score-matrix rows of the `nrows` best rows (rows[i],
// low word = row index) are fetched with ONE bulk copy per row (cp.async.bulk: no registers, no per-chunk copy
// instructions, every row in flight at once — the DRAM latency is paid once per pass) into a staging buffer of
// P.nms_stage_rows rows, and every entry inside the key window [lo, hi] lands in the stash `out` as a 64-bit key.
// Rows are 16-byte multiples (C % 4 == 0). More rows than the buffer holds: further passes. Returns the number of
// stashed keys (may exceed cap: the caller then falls back to the generic source scan); their low words hold the position
// inside the row list (see nms_stash_to_flat). All threads call.
// `bar` = the block's staging mbarrier (initialised with count 1), `phase` = its running parity.
__device__ __noinline__ int nms_bulk_scan(const DevParams& P, const uint32_t* mat, const u64* rows, int nrows, u64 lo, u64 hi,
                                          u64* out, int cap, unsigned char* stage, int* count, uint64_t* bar, uint32_t& phase) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int C = P.C, C4 = C >> 2, H = P.nms_stage_rows;
    const uint32_t row_bytes = (uint32_t)C * 4u;
    ScoreWindow win;
    win.set(lo, hi);
    if (tid == 0) *count = 0;
#ifdef YPP_PROFILE
    long long pt[3] = {0, 0, 0}, pt_t = clock64();
#define YPP_ACC2(i) do { long long t2 = clock64(); pt[i] += t2 - pt_t; pt_t = t2; } while (0)
#else
#define YPP_ACC2(i) do { } while (0)
#endif
    int npass = 0;
#pragma unroll 1
    for (int r0 = 0; r0 < nrows; r0 += H, ++npass) {
        const int n = min(H, nrows - r0);
        fence_proxy_async();  // this thread's generic reads of the buffer (earlier pass / image) -> the async-proxy refill
        __syncthreads();                  // (first pass: publishes *count = 0)
        // (complete_tx of a copy may precede the expect_tx: the phase cannot complete before thread 0 has arrived)
        if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)n * row_bytes);
#pragma unroll 1
        for (int i = tid; i < n; i += NMS_THREADS)
            bulk_load_1d(stage + (size_t)i * row_bytes, mat + (size_t)(uint32_t)rows[r0 + i] * C, row_bytes, bar);
        YPP_ACC2(0);
        mbar_wait(bar, phase);
        phase ^= 1u;
        YPP_ACC2(1);
        // the buffer is n * C4 consecutive 16-byte groups: thread-contiguous, conflict-free reads; survivors of up to
        // 8 groups per thread are flushed together (one warp scan + one shared-memory atomic per flush)
        const uint4* st4 = reinterpret_cast<const uint4*>(stage);
        const int ng = n * C4;
#pragma unroll 1
        for (int base = 0; base < ng; base += NMS_THREADS * 8) {
            unsigned em = 0u;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int g = base + q * NMS_THREADS + tid;
                if (g < ng) {
                    const uint4 w = st4[g];
                    uint32_t f0 = 0u;
                    if (win.ends) {
                        const int r = g / C4;
                        f0 = (uint32_t)rows[r0 + r] * (uint32_t)C + 4u * (uint32_t)(g - r * C4);
                    }
                    em |= win.test4(w, f0) << (q * 4);
                }
            }
            if (__any_sync(0xffffffffu, em != 0u)) {
                const int c = __popc(em);
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int sp = 0;
                if (lane == 31) sp = atomicAdd(count, incl);
                sp = __shfl_sync(0xffffffffu, sp, 31) + incl - c;
                // (the stash entry carries the position inside the row list for now: no division, no row lookup here)
#pragma unroll 1
                while (em) {
                    const int pos = __ffs(em) - 1;
                    em &= em - 1;
                    const int e = 4 * (base + (pos >> 2) * NMS_THREADS + tid) + (pos & 3);
                    const uint32_t word = reinterpret_cast<const uint32_t*>(stage)[e];
                    if (sp < cap) out[sp] = score_key(word, (uint32_t)(r0 * C + e));
                    ++sp;
                }
            }
        }
        YPP_ACC2(2);
    }
#ifdef YPP_PROFILE
    if (tid == 0 && blockIdx.x < 256) {
        g_phase[1][blockIdx.x][12] = pt[0];
        g_phase[1][blockIdx.x][13] = pt[1];
        g_phase[1][blockIdx.x][14] = pt[2];
        g_phase[1][blockIdx.x][15] = npass;
    }
#endif
    __syncthreads();
    return *count;
}

// The stash of nms_bulk_scan carries, in the low word of each key, the candidate's position inside the row list
// (list index * C + class): this turns it into the flat candidate index row * C + class. All threads call.
__device__ __forceinline__ void nms_stash_to_flat(u64* keys, int n, const u64* rows, int C) {
#pragma unroll 1
    for (int i = threadIdx.x; i < n; i += NMS_THREADS) {
        const u64 k = keys[i];
        const uint32_t e = (uint32_t)k, r = e / (uint32_t)C;
        keys[i] = (k & 0xFFFFFFFF00000000ull) | (u64)((uint32_t)rows[r] * (uint32_t)C + (e - r * (uint32_t)C));
    }
    __syncthreads();
}

// Greedy NMS of one chunk in the classes-are-independent regime (mmcv batched_nms, n >= split_thr: a kept box only
// suppresses boxes of its own class): the classes are resolved IN PARALLEL instead of walking the chunk in score
// order, and the chunk is never sorted as a whole. A candidate's fate depends only on the better-scored candidates of
// its own class, so the keep flag of every candidate of the chunk is exact; only the kept ones are then sorted by key
// (score desc, flat asc) and the first (cap - kept so far) of them appended to the kept list.
//   1. class member lists (counting sort by class, unordered), then each candidate's position inside its class =
//      number of members with a smaller key -> ordered lists;
//   2. warp per class, lanes over the class's pairs (a < b in class order): IoU -> 64-bit mask `supby[b]` (bit a: the
//      member at class position a would suppress b);
//   3. candidates with an empty mask are kept (unless a box kept from an earlier chunk suppresses them); the
//      undecided members of a class are walked in order by one thread with bit operations only;
//   4. the kept candidates are compacted and sorted: (key, index) pairs in registers, one per thread.
// key[0..m): the chunk's candidates in any order (the complete set of candidates inside a key window), m <= NMS_CH;
// cx1..ccl: their (offset) boxes / areas / classes in the same order; kx1..knext/chead/kcl/kkey: kept list with
// per-class chains, nk boxes so far. Returns the new number of kept boxes, or -1 — kept list untouched — when the
// pass does not apply: a class with more than NMS_CLS_MAX candidates in the chunk, or more than NMS_THREADS kept
// (the caller then sorts the chunk and runs the group-wise pass). The loops are deliberately not unrolled: the
// per-image kernels run their code about once and are bound by instruction issue and fetch.
constexpr int NMS_CLS_MAX = 64;              // candidates of one class (one mask word)
constexpr int NMS_CLS_LABELS = 128;          // classes the parallel pass has scratch for
struct NmsClsSmem {
    int ccnt[NMS_CLS_LABELS];
    int coff[NMS_CLS_LABELS + 1];
    int over, nkept;
    unsigned short ulist[NMS_CH];                     // members of each class, unordered
    unsigned short clist[NMS_CH];                     // members of each class, in score order
    unsigned char cpos[NMS_CH];                       // position of candidate i inside its class
    unsigned char kflag[NMS_CH];                      // result: kept
    unsigned keptm[NMS_CLS_LABELS][2], unc[NMS_CLS_LABELS][2];  // per class: members kept for sure / undecided (bit = position)
    unsigned short spay[NMS_THREADS], tpay[NMS_THREADS];       // kept candidates: index (payload of the sort) + scratch
};
static_assert(sizeof(NmsClsSmem) <= NMS_KCAP * 8, "lives in the row-list buffer, which is free between two scans");
__device__ __noinline__ bool nms_iou_gt(const float* x1, const float* y1, const float* x2, const float* y2, const float* ar, int i,
                                        const Box& bj, float thr, float foff) {
    Box bi;
    bi.x1 = x1[i];
    bi.y1 = y1[i];
    bi.x2 = x2[i];
    bi.y2 = y2[i];
    bi.area = ar[i];
    return iou_gt(bi, bj, thr, foff);
}
__device__ __noinline__ int nms_resolve_classes(int m, int nlab, int nk, int cap, float thr, float foff, const u64* key,
                                                const float* cx1, const float* cy1, const float* cx2, const float* cy2,
                                                const float* car, const int* ccl, float* kx1, float* ky1, float* kx2, float* ky2,
                                                float* kar, int* kcl, u64* kkey, int* knext, int* chead, NmsClsSmem& Q,
                                                u64* scratch /* 2 * NMS_CH keys */, u64 lo, u64 hi, TopSelSmem& S) {
    const int tid = threadIdx.x, lane = tid & 31;
    u64* supby = scratch;                 // [NMS_CH]
    u64* skey = scratch + NMS_CH;         // [NMS_THREADS] kept keys
    u64* tkey = skey + NMS_THREADS;       // [NMS_THREADS] scratch of the sort
    static_assert(2 * NMS_THREADS <= NMS_CH, "scratch layout");
    YPP_SUB(8);
    // 1a. class sizes
    if (tid < nlab) Q.ccnt[tid] = 0;
    if (tid == 0) {
        Q.over = 0;
        Q.nkept = 0;
    }
    __syncthreads();
#pragma unroll 1
    for (int i = tid; i < m; i += NMS_THREADS) atomicAdd(&Q.ccnt[ccl[i]], 1);
    __syncthreads();
    if (tid < 32) {
        // exclusive scan over the classes (each lane owns a contiguous run), largest class
        const int per = (nlab + 31) / 32, c0 = lane * per, c1 = min(nlab, c0 + per);
        int sum = 0, mx = 0;
#pragma unroll 1
        for (int c = c0; c < c1; ++c) {
            const int v = Q.ccnt[c];
            sum += v;
            mx = v > mx ? v : mx;
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - sum;
#pragma unroll 1
        for (int c = c0; c < c1; ++c) {
            const int v = Q.ccnt[c];
            Q.coff[c] = run;
            Q.ccnt[c] = 0;  // becomes the fill cursor
            run += v;
        }
        if (lane == 31) Q.coff[nlab] = incl;
        if (__reduce_max_sync(0xffffffffu, mx) > NMS_CLS_MAX && lane == 0) Q.over = 1;
    }
    __syncthreads();
    YPP_SUB(9);
    if (Q.over) return -1;
    // 1b. member lists (unordered)
#pragma unroll 1
    for (int i = tid; i < m; i += NMS_THREADS) {
        const int c = ccl[i];
        Q.ulist[Q.coff[c] + atomicAdd(&Q.ccnt[c], 1)] = (unsigned short)i;
    }
    if (tid < 2 * nlab) {
        (&Q.keptm[0][0])[tid] = 0u;
        (&Q.unc[0][0])[tid] = 0u;
    }
    __syncthreads();
    YPP_SUB(10);
    // 1c. position inside the class = members with a smaller key (keys are unique) -> ordered lists
#pragma unroll 1
    for (int i = tid; i < m; i += NMS_THREADS) {
        const int c = ccl[i], o0 = Q.coff[c], n = Q.coff[c + 1] - o0;
        const u64 mine = key[i];
        int p = 0;
#pragma unroll 1
        for (int t = 0; t < n; ++t) p += key[Q.ulist[o0 + t]] < mine ? 1 : 0;
        Q.cpos[i] = (unsigned char)p;
        Q.clist[o0 + p] = (unsigned short)i;
        supby[i] = 0ull;
    }
    __syncthreads();
    YPP_SUB(11);
    // 2. suppression masks: warp per class, lanes over the class's pairs, so the work is balanced whatever the
    // class sizes are
    uint32_t* sup32 = reinterpret_cast<uint32_t*>(supby);
#pragma unroll 1
    for (int c = tid >> 5; c < nlab; c += NMS_THREADS / 32) {
        const int o0 = Q.coff[c], n = Q.coff[c + 1] - o0, npairs = (n * (n - 1)) >> 1;
#pragma unroll 1
        for (int e = lane; e < npairs; e += 32) {
            int pb = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)e)) * 0.5f);  // e = pb (pb - 1) / 2 + pa, pa < pb
#pragma unroll 1
            while (((pb * (pb - 1)) >> 1) > e) --pb;
#pragma unroll 1
            while ((((pb + 1) * pb) >> 1) <= e) ++pb;
            const int pa = e - ((pb * (pb - 1)) >> 1);
            const int i = (int)Q.clist[o0 + pa], j = (int)Q.clist[o0 + pb];
            Box bj;
            bj.x1 = cx1[j];
            bj.y1 = cy1[j];
            bj.x2 = cx2[j];
            bj.y2 = cy2[j];
            bj.area = car[j];
            if (nms_iou_gt(cx1, cy1, cx2, cy2, car, i, bj, thr, foff)) atomicOr(&sup32[2 * j + (pa >> 5)], 1u << (pa & 31));
        }
    }
    __syncthreads();
    YPP_SUB(12);
    // 3a. candidates nobody can suppress are kept (if alive), the rest wait for their class's pass
#pragma unroll 1
    for (int j = tid; j < m; j += NMS_THREADS) {
        const int c = ccl[j], p = (int)Q.cpos[j];
        bool dead = false;
        if (nk > 0) {
            Box bj;
            bj.x1 = cx1[j];
            bj.y1 = cy1[j];
            bj.x2 = cx2[j];
            bj.y2 = cy2[j];
            bj.area = car[j];
#pragma unroll 1
            for (int kk = chead[c]; kk >= 0 && !dead; kk = knext[kk]) dead = nms_iou_gt(kx1, ky1, kx2, ky2, kar, kk, bj, thr, foff);
        }
        const bool sure = !dead && supby[j] == 0ull;
        Q.kflag[j] = sure ? 1 : 0;
        if (!dead) atomicOr(sure ? &Q.keptm[c][p >> 5] : &Q.unc[c][p >> 5], 1u << (p & 31));
    }
    __syncthreads();
    // 3b. greedy pass per class over the undecided members, in score order: bit operations only
    if (tid < nlab) {
        u64 todo = ((u64)Q.unc[tid][1] << 32) | (u64)Q.unc[tid][0];
        if (todo) {
            const int o0 = Q.coff[tid];
            u64 kept = ((u64)Q.keptm[tid][1] << 32) | (u64)Q.keptm[tid][0];
#pragma unroll 1
            while (todo) {
                const int p = __ffsll((long long)todo) - 1;
                todo &= todo - 1ull;
                const int j = (int)Q.clist[o0 + p];
                if ((supby[j] & kept) == 0ull) {
                    kept |= 1ull << p;
                    Q.kflag[j] = 1;
                }
            }
        }
    }
    __syncthreads();
    YPP_SUB(13);
    // 4. kept candidates -> (key, index) pairs, compacted (any order), sorted by key. Only the first (cap - nk) in
    // key order are needed: when more than NMS_THREADS candidates were kept (the sort takes one per thread), a
    // 512-bin histogram over the chunk's score window first drops the ones that cannot be among them.
    const int room = cap - nk;
    const int tid_w = tid >> 5;
#pragma unroll 1
    for (int i0 = 0; i0 < m; i0 += NMS_THREADS) {  // (uniform trip count)
        const int i = i0 + tid;
        const unsigned bal = __ballot_sync(0xffffffffu, i < m && Q.kflag[i]);
        if (lane == 0 && bal) atomicAdd(&Q.nkept, __popc(bal));
    }
    __syncthreads();
    const int K0 = Q.nkept;
    const uint32_t w_lo = (uint32_t)(lo >> 32), w_range = (uint32_t)(hi >> 32) - w_lo;
    const int shift = w_range >= 512u ? (32 - __clz(w_range)) - 9 : 0;
    int pb = 0x7FFFFFFF;  // last histogram bin that is kept
    if (K0 > NMS_THREADS) {
        static_assert(NMS_THREADS == 512, "one histogram bin per thread");
        S.hist[tid] = 0;
        if (tid == 0) S.kb = -1;
        __syncthreads();
#pragma unroll 1
        for (int i = tid; i < m; i += NMS_THREADS)
            if (Q.kflag[i]) atomicAdd(&S.hist[((uint32_t)(key[i] >> 32) - w_lo) >> shift], 1);
        __syncthreads();
        const int v = S.hist[tid];
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) S.wsum[tid_w] = incl;
        __syncthreads();
        {
            const int wv = lane < NMS_THREADS / 32 ? S.wsum[lane] : 0;
            int wi = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            incl += __shfl_sync(0xffffffffu, wi - wv, tid_w);
        }
        // the first bin where the count from the best score down reaches `room`; everything up to it must fit the sort
        if (incl - v < room && room <= incl) S.kb = incl <= NMS_THREADS ? tid : -2;
        __syncthreads();
        pb = S.kb;
        __syncthreads();
        if (pb < 0) return -1;  // (-2: too many kept candidates share the pivot bin)
    }
    __syncthreads();  // (every thread has read the count)
    if (tid == 0) Q.nkept = 0;
    __syncthreads();
#pragma unroll 1
    for (int i0 = 0; i0 < m; i0 += NMS_THREADS) {  // (uniform trip count)
        const int i = i0 + tid;
        const bool kf = i < m && Q.kflag[i] && (int)(((uint32_t)(key[i] >> 32) - w_lo) >> shift) <= pb;
        const unsigned bal = __ballot_sync(0xffffffffu, kf);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&Q.nkept, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (kf) {
            const int t = base + __popc(bal & ((1u << lane) - 1u));
            if (t < NMS_THREADS) {
                skey[t] = key[i];
                Q.spay[t] = (unsigned short)i;
            }
        }
    }
    __syncthreads();
    const int K = Q.nkept;
    if (K > NMS_THREADS) return -1;
    const int p2 = next_pow2(K < 32 ? 32 : K);
    if (tid >= K && tid < p2) {
        skey[tid] = ~0ull;
        Q.spay[tid] = 0;
    }
    __syncthreads();
    YPP_SUB(14);
    bitonic_sort_kv(skey, Q.spay, tkey, Q.tpay, p2);
    YPP_SUB(15);
    // the first (cap - nk) of them join the kept list, in score order
    const int take = K < room ? K : room;
    if (tid < take) {
        const int gj = (int)Q.spay[tid], kidx = nk + tid;
        kx1[kidx] = cx1[gj];
        ky1[kidx] = cy1[gj];
        kx2[kidx] = cx2[gj];
        ky2[kidx] = cy2[gj];
        kar[kidx] = car[gj];
        kcl[kidx] = ccl[gj];
        kkey[kidx] = skey[tid];
        knext[kidx] = atomicExch(&chead[ccl[gj]], kidx);  // class chain (order irrelevant)
    }
    __syncthreads();
    YPP_SUB(16);
    return nk + take;
}

// bbox2result (mmdet/core/bbox/transforms.py:110-116): `bboxes[labels == i, :]` for every class, i.e. a STABLE
// counting sort of the <= max_per_img output rows by label, done on the device so that the host gets one block it
// can slice into num_classes views. This part: per-class counts of the nk output rows -> exclusive offsets, written
// to o_cls_offsets[b][0..C] and left in cnt[0..C) (C ints of scratch). The rows themselves are placed by the output
// loop: position = offset of the class + number of kept boxes of that class with a smaller output index, counted
// along the class chain of the kept list. All threads of the NMS block call.
__device__ __noinline__ void nms_label_offsets(const DevParams& P, int b, int nk, const int* kcl, int* cnt) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, C = P.C;
#pragma unroll 1
    for (int c = tid; c < C; c += NMS_THREADS) cnt[c] = 0;
    __syncthreads();
#pragma unroll 1
    for (int i = tid; i < nk; i += NMS_THREADS) atomicAdd(&cnt[kcl[i]], 1);
    __syncthreads();
    if (warp == 0) {
        // exclusive scan over the classes: each lane owns a contiguous run
        const int per = (C + 31) / 32, c0 = lane * per, c1 = min(C, c0 + per);
        int sum = 0;
#pragma unroll 1
        for (int c = c0; c < c1; ++c) sum += cnt[c];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - sum;
        int* offs = P.o_cls_offsets + (size_t)b * (C + 1);
#pragma unroll 1
        for (int c = c0; c < c1; ++c) {
            const int v = cnt[c];
            cnt[c] = run;
            offs[c] = run;
            run += v;
        }
        if (lane == 31) offs[C] = incl;  // == nk
    }
    __syncthreads();
}

// Upper key bound of a chunk of the NMS candidate stream: the W-th best ROW maximum is a lower bound of the W-th
// best candidate score (each of those rows owns a candidate at least that good), so when a chunk must reach
// cumulative rank W only the rows whose best score reaches that bound can matter, and only their entries at or
// above it. The bound need not be exact: a 512-bin histogram of the row maxima gives T = lower edge of the bin in
// which the count from the top reaches W (round 1 sorted the row maxima for this: 20 k cycles per image; this is
// ~4 k). Lists those rows (unordered) in rowkeys[] and returns their number, or -1 (fewer than W candidate rows /
// list longer than NMS_KCAP: the caller scans the whole matrix). bmax / bmin = ord of the best / worst candidate
// score of the image. All threads of the NMS block call.
__device__ __noinline__ int nms_pick_rows(const DevParams& P, int b, int W, uint32_t bmax, uint32_t bmin, u64* rowkeys,
                                          TopSelSmem& S, uint32_t* t_ord, const uint32_t* rmax) {
    static_assert(NMS_THREADS == 512, "one histogram bin per thread");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint4* rs = P.row_stat + (size_t)b * P.R;
    const uint32_t range = bmax - bmin;
    const int shift = range >= 512u ? (32 - __clz(range)) - 9 : 0;
    S.hist[tid] = 0;
    if (tid == 0) {
        S.count = 0;
        S.kb = -1;
    }
    __syncthreads();
    // rmax: the rows' best scores as the reduction pass left them in shared memory (0: no candidate), or null
#pragma unroll 1
    for (int r = tid; r < P.R; r += NMS_THREADS) {
        uint32_t x;
        if (rmax) {
            x = rmax[r];
        } else {
            const uint4 st = rs[r];
            x = st.z ? st.x : 0u;
        }
        if (x) atomicAdd(&S.hist[(bmax - x) >> shift], 1);
    }
    __syncthreads();
    const int v = S.hist[tid];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) S.wsum[warp] = incl;
    __syncthreads();
    int excl = incl - v;
    {
        // bins of the warps before this one: every warp scans the 16 warp sums itself
        const int wv = lane < NMS_THREADS / 32 ? S.wsum[lane] : 0;
        int wi = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        excl += __shfl_sync(0xffffffffu, wi - wv, warp);
    }
    if (excl < W && W <= excl + v) S.kb = tid;
    __syncthreads();
    const int pb = S.kb;
    if (pb < 0) return -1;
#pragma unroll 1
    for (int r = tid; r < P.R; r += NMS_THREADS) {
        uint32_t x;
        if (rmax) {
            x = rmax[r];
        } else {
            const uint4 st = rs[r];
            x = st.z ? st.x : 0u;
        }
        if (x && (int)((bmax - x) >> shift) <= pb) {
            const int sp = atomicAdd(&S.count, 1);
            if (sp < NMS_KCAP) rowkeys[sp] = (u64)(uint32_t)r;
        }
    }
    __syncthreads();
    const int n = S.count;
    const u64 below = (((u64)pb + 1ull) << shift) - 1ull;  // largest (bmax - ord) inside the pivot bin
    *t_ord = below >= (u64)bmax ? 0u : bmax - (uint32_t)below;
    __syncthreads();  // S.count / S.kb are reused by the scan that follows
    return n <= NMS_KCAP ? n : -1;
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
__device__ __forceinline__ void nms_image_body(const DevParams& P, const int b, unsigned char* nms_smem, uint64_t* stage_bar,
                                               uint32_t& stage_phase) {
    __shared__ TopSelSmem S;
    __shared__ u64 s_sup, s_masks[NMS_G];
    __shared__ int s_nk, s_stash;
    __shared__ uint32_t s_red[4];
    const int cap = P.keep_cap;
    // The kept list (key, offset box, area, class, chain link per kept box) lives in shared memory up to NMS_MAX_KEEP
    // boxes; beyond that ("keep all" with many survivors: max_per_img = -1, RPN-sized standalone NMS) it lives in
    // the caller's workspace — same code, the pointers just lead to global memory (L2 resident).
    const bool kept_global = P.nms_kept != nullptr;
    const int cap_s = kept_global ? 0 : cap;
    u64* keys = reinterpret_cast<u64*>(nms_smem);                 // [NMS_KCAP] sorted chunk
    u64* ktmp = keys + NMS_KCAP;                                   // [NMS_KCAP] scratch of the select
    u64* kkey = ktmp + NMS_KCAP;                                   // [cap_s]
    float* cx1 = reinterpret_cast<float*>(kkey + cap_s);           // [NMS_CH] x 5
    float* cy1 = cx1 + NMS_CH;
    float* cx2 = cy1 + NMS_CH;
    float* cy2 = cx2 + NMS_CH;
    float* car = cy2 + NMS_CH;
    int* ccl = reinterpret_cast<int*>(car + NMS_CH);               // [NMS_CH]
    float* kx1 = reinterpret_cast<float*>(ccl + NMS_CH);           // [cap_s] x 5
    int* chead = reinterpret_cast<int*>(kx1 + 7 * (size_t)cap_s);  // [C]   latest kept box of each class (-1: none)
    if (kept_global) {
        unsigned char* gk = P.nms_kept + (size_t)b * (size_t)P.nms_kept_stride;
        kkey = reinterpret_cast<u64*>(gk);
        kx1 = reinterpret_cast<float*>(kkey + cap);
    }
    float* ky1 = kx1 + cap;
    float* kx2 = ky1 + cap;
    float* ky2 = kx2 + cap;
    float* kar = ky2 + cap;
    int* kcl = reinterpret_cast<int*>(kar + cap);                  // [cap]
    int* knext = kcl + cap;                                        // [cap] previous kept box of the same class
    u64* rowkeys = reinterpret_cast<u64*>(nms_smem + P.nms_rowkeys_off);  // [NMS_KCAP] rows of the candidate scan
    NmsClsSmem& Q = *reinterpret_cast<NmsClsSmem*>(rowkeys);              // (scratch of the class-parallel pass, between scans)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = P.C;
    const int slots = P.R * C;
    const bool generic = P.generic != 0;
    const uint32_t* mat = generic ? reinterpret_cast<const uint32_t*>(P.g_scores) : P.mat + (size_t)b * slots;
    const float4* row_box = P.row_box + (size_t)b * P.R;
    const int nlab = generic ? P.num_labels : C;
    // the rows' best scores, kept for the first row pick (the select scratch is free until then)
    uint32_t* rmax = (!generic && P.R <= 2 * NMS_KCAP) ? reinterpret_cast<uint32_t*>(ktmp) : nullptr;

    // candidate count, score range and boxes.max() of the image: reduction over the per-row statistics that the
    // decode kernels stored
    YPP_PHASE(1, b, 0);
#ifdef YPP_PROFILE
    if (threadIdx.x == 0) { S.prof_kernel = 1; S.prof_call = 0; }
#endif
    if (tid == 0) {
        s_nk = 0;
        s_sup = 0ull;
        s_red[0] = 0u;
        s_red[1] = 0u;
        s_red[2] = 0u;
        s_red[3] = 0u;
    }
    __syncthreads();
    {
        uint32_t best = 0u, worst = 0u, mx = 0u, cnt = 0u;
        if (!generic) {
            const uint4* rs = P.row_stat + (size_t)b * P.R;
#pragma unroll 1
            for (int r = tid; r < P.R; r += NMS_THREADS) {
                const uint4 st = rs[r];
                if (rmax) rmax[r] = st.z ? st.x : 0u;
                if (st.z) {
                    best = st.x > best ? st.x : best;
                    worst = st.y > worst ? st.y : worst;
                    cnt += st.z;
                    // boxes.max() runs over candidate boxes only (precomputed per row when boxes are per class)
                    const uint32_t o = st.w ? st.w : f2ord(box_max(row_box[r]));
                    mx = o > mx ? o : mx;
                }
            }
        } else {
            // every input box is a candidate: key range over all scores, boxes.max() over all boxes
#pragma unroll 1
            for (int i = tid; i < slots; i += NMS_THREADS) {
                const uint32_t o = f2ord(__uint_as_float(mat[i]));
                best = o > best ? o : best;
                worst = ~o > worst ? ~o : worst;
                const uint32_t bo = f2ord(box_max(row_box[i]));
                mx = bo > mx ? bo : mx;
                ++cnt;
            }
        }
        best = __reduce_max_sync(0xffffffffu, best);
        worst = __reduce_max_sync(0xffffffffu, worst);
        mx = __reduce_max_sync(0xffffffffu, mx);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && cnt) {
            atomicMax(&s_red[0], best);
            atomicMax(&s_red[1], worst);
            atomicMax(&s_red[2], mx);
            atomicAdd(&s_red[3], cnt);
        }
    }
    __syncthreads();
    YPP_PHASE(1, b, 1);
    const int ntot = (int)s_red[3];
    if (tid == 0 && P.o_ncand) P.o_ncand[b] = ntot;
    if (ntot == 0) {
        if (tid == 0) P.o_count[b] = 0;
        if (P.o_cls_offsets && !generic) nms_label_offsets(P, b, 0, kcl, reinterpret_cast<int*>(ktmp));  // all groups empty
        return;
    }
    const u64 gmin = (u64)(~s_red[0]) << 32;
    u64 gmax = ((u64)s_red[1] << 32) | 0xFFFFFFFFull;
    if (P.nms_score_thr > 0.f) {
        // NMSop.forward: `valid_mask = scores > score_threshold` before the greedy pass — in key space: keys whose high
        // word is >= ~ord(threshold) are out. (The regime decision above `ntot` and boxes.max() stay unfiltered, as
        // in batched_nms, which filters inside nms().)
        const u64 lim = (u64)(~f2ord(P.nms_score_thr)) << 32;
        if (lim == 0ull || lim - 1ull < gmin) {
            if (tid == 0) P.o_count[b] = 0;
            if (P.o_cls_offsets && !generic) nms_label_offsets(P, b, 0, kcl, reinterpret_cast<int*>(ktmp));
            return;
        }
        gmax = gmax < lim - 1ull ? gmax : lim - 1ull;
    }
    const uint32_t img_max_ord = s_red[2];
    const bool per_class = !(ntot < P.split_thr);  // regime of mmcv batched_nms
#pragma unroll 1
    for (int c = tid; c < nlab; c += NMS_THREADS) chead[c] = -1;
    const bool use_off = !P.nms_agnostic;
    const float mp1 = fadd(ord2f(img_max_ord), 1.0f);  // max_coordinate + 1
    const float thr = P.iou_thr, foff = P.foff;

    MatSource msrc;
    msrc.m = mat;
    msrc.slots = slots;
    msrc.vec4 = ((slots & 3) == 0);
    const bool nonneg = s_red[1] <= 0x7FFFFFFFu;  // every candidate score is >= +0: raw words order like the keys
    msrc.windowed = nonneg;  // (the standalone entries accept scores of any sign: exact 64-bit test per element)

    int processed = 0;
    u64 lo = gmin;
    // the first chunk only needs a little more than `cap` candidates; later chunks (heavy suppression) are full
    int chunk = cap + (cap >> 2) + 64;
    chunk = chunk < NMS_CH ? chunk : NMS_CH;
    // ... and its rows should fit the staging buffer in one pass (the pivot bin adds a few rows beyond `chunk`)
    if (P.nms_stage_rows - 24 >= cap + 32 && chunk > P.nms_stage_rows - 24) chunk = P.nms_stage_rows - 24;
    // classes-in-parallel pass: available when the classes are independent and the scratch covers the label range. It
    // takes a whole stash of at most NMS_CH candidates. The stash of a scan that must reach rank W holds every entry of
    // the W best rows above the W-th best row maximum — measured 1.2 W .. 1.5 W candidates — so the first chunk asks
    // for a little under max_num rows (what is missing after it — suppression, or a stash at the low end — is made up by a
    // further, smaller chunk); a stash that comes out too large goes the sorted, group-wise way.
    const bool cls_parallel = per_class && !generic && nlab <= NMS_CLS_LABELS && P.nms_stage_rows > 0;
    const int chunk_sorted = chunk;
#ifndef YPP_NMS_W_NUM
#define YPP_NMS_W_NUM 19 // rows of the first chunk = max_num * 19 / 20 (measured, 64 images each: 608^2 stash 1.5 - 1.65 W,
#define YPP_NMS_W_DEN 20 // 1280^2 1.15 - 1.4 W, YOLOv3 640^2 1.12 - 1.45 W; at 8 / 10 and 17 / 20 some images of the last two needed a
                         // second chunk: 35 -> 56 us per launch)
#endif
    if (cls_parallel && cap <= NMS_THREADS) chunk = cap < 36 ? 32 : (cap * YPP_NMS_W_NUM) / YPP_NMS_W_DEN;
    YPP_PHASE(1, b, 2);
#ifdef YPP_PROFILE
    int prof_chunks = 0, prof_groups = 0;
    long long prof_ab1 = 0, prof_b2 = 0;
#endif
#pragma unroll 1
    while (processed < ntot && s_nk < cap) {
        if (processed > 0) {
            chunk = NMS_CH;
            if (cls_parallel) {
                // further chunks of the class-parallel pass: twice what is still missing (at least 128 rows)
                const int miss = 2 * (cap - s_nk) + 64;
                chunk = miss < 128 ? 128 : (miss > NMS_THREADS ? NMS_THREADS : miss);
            }
        }
        const int want = min(chunk, ntot - processed);
        const int want_sorted = processed == 0 ? min(chunk_sorted, ntot) : want;  // (the stash usually holds more than `want`)
        const int wc = processed + want;  // cumulative rank this chunk must reach
        u64 hi = gmax;
        int nrows = -1;
        if (!generic && nonneg) {
            uint32_t t_ord = 0u;
            YPP_SUB(0);
            nrows = nms_pick_rows(P, b, wc, s_red[0], ~s_red[1], rowkeys, S, &t_ord, rmax);
            rmax = nullptr;  // (the scratch is reused from here on)
            YPP_SUB(1);
            if (nrows >= 0) {
                const u64 hi_t = ((u64)(~t_ord) << 32) | 0xFFFFFFFFull;
                hi = hi_t < gmax ? hi_t : gmax;
            }
        }
        // boxes of the chunk's candidates keys[0..n): boxes + idxs.to(boxes) * (max_coordinate + 1), areas, classes
        // (from_list: the keys still carry positions inside the row list — see nms_stash_to_flat — and become flat here)
        auto stage_boxes = [&](int n, bool from_list) {
#pragma unroll 1
            for (int i = tid; i < n; i += NMS_THREADS) {
                uint32_t flat = key_flat(keys[i]);
                int r = (int)(flat / (uint32_t)C);
                const int c0 = (int)(flat - (uint32_t)r * (uint32_t)C);
                if (from_list) {
                    r = (int)(uint32_t)rowkeys[r];
                    flat = (uint32_t)r * (uint32_t)C + (uint32_t)c0;
                    keys[i] = (keys[i] & 0xFFFFFFFF00000000ull) | (u64)flat;
                }
                const int c = generic ? (P.g_labels ? (int)P.g_labels[flat] : 0) : c0;
                const float4 bx = row_box[P.boxes_per_class ? flat : (uint32_t)r];
                float x1 = bx.x, y1 = bx.y, x2 = bx.z, y2 = bx.w;
                if (use_off) {
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
        };
        int got;
        if (nrows >= 0) {
            // only the listed rows can hold one of the wc best candidates: scan just their matrix rows
            int staged = -1;
            bool keys_flat = false;  // the stash keys carry flat indices already (else: positions inside the row list)
            if (P.nms_stage_rows > 0)
                staged = nms_bulk_scan(P, mat, rowkeys, nrows, lo, hi, keys, NMS_KCAP, nms_smem + P.nms_stage_off, &s_stash,
                                       stage_bar, stage_phase);
            YPP_SUB(2);
            if (cls_parallel && staged > 0 && staged <= NMS_CH && cap - s_nk <= NMS_THREADS) {
                // the stash IS the complete set of candidates inside the window [lo, hi] — a prefix of the global order
                // that reaches rank wc: resolve its classes in parallel, unsorted
                stage_boxes(staged, true);
                keys_flat = true;
                YPP_SUB(3);
                const int nk1 = nms_resolve_classes(staged, nlab, s_nk, cap, thr, foff, keys, cx1, cy1, cx2, cy2, car, ccl, kx1, ky1,
                                                    kx2, ky2, kar, kcl, kkey, knext, chead, Q, ktmp, lo, hi, S);
                if (processed == 0) {
                    YPP_SUBV(26, nrows);
                    YPP_SUBV(27, staged);
                    YPP_SUBV(28, Q.nkept);
                    YPP_SUBV(29, nk1);
                }
                if (nk1 >= 0) {
#ifdef YPP_PROFILE
                    if (prof_chunks == 0) {
                        YPP_PHASE(1, b, 3);
                        YPP_PHASE(1, b, 4);
                    }
                    ++prof_chunks;
#endif
                    if (tid == 0) s_nk = nk1;
                    processed += staged;
                    if (hi == ~0ull) processed = ntot;
                    lo = hi + 1ull;
                    __syncthreads();
                    continue;
                }
            }
            if (staged >= 0 && staged <= NMS_KCAP) {
                if (!keys_flat) nms_stash_to_flat(keys, staged, rowkeys, C);
                StashSource none;
                got = select_sorted_prefix(none, lo, hi, want_sorted, keys, ktmp, NMS_KCAP, S, staged);
            } else {
                RowListSource rl;
                rl.m = mat;
                rl.rows = rowkeys;
                rl.nrows = nrows;
                rl.C = C;
                rl.nsub = (C + 127) / 128;
                rl.vec4 = ((C & 3) == 0);
                rl.win.set(lo, hi);
                __syncthreads();
                got = select_sorted_prefix(rl, lo, hi, want, keys, ktmp, NMS_KCAP, S);
            }
        } else {
            msrc.win.set(lo, hi);
            got = select_sorted_prefix(msrc, lo, hi, want, keys, ktmp, NMS_KCAP, S);
        }
        const int m = got < NMS_CH ? got : NMS_CH;  // boxes staged this round (a prefix of the sorted order)
#ifdef YPP_PROFILE
        if (prof_chunks == 0) YPP_PHASE(1, b, 3);
        ++prof_chunks;
#endif
        if (m == 0) break;
        stage_boxes(m, false);
#ifdef YPP_PROFILE
        if (prof_chunks == 1) YPP_PHASE(1, b, 4);
#endif
#pragma unroll 1
        for (int s0 = 0; s0 < m; s0 += NMS_G) {
            const int nk = s_nk;
            if (nk >= cap) break;
#ifdef YPP_PROFILE
            ++prof_groups;
            long long pg_t0 = clock64();
#endif
            // ---- phase A: the group's 64 candidates against the kept list
            if (per_class) {
                // classes are independent: walk the chain of kept boxes of the candidate's own class (a handful)
                if (tid < NMS_G) {
                    const int j = s0 + tid;
                    bool sup = false;
                    if (j < m) {
                        Box bj;
                        bj.x1 = cx1[j];
                        bj.y1 = cy1[j];
                        bj.x2 = cx2[j];
                        bj.y2 = cy2[j];
                        bj.area = car[j];
#pragma unroll 1
                        for (int k = chead[ccl[j]]; k >= 0 && !sup; k = knext[k]) {
                            Box bk;
                            bk.x1 = kx1[k];
                            bk.y1 = ky1[k];
                            bk.x2 = kx2[k];
                            bk.y2 = ky2[k];
                            bk.area = kar[k];
                            sup = iou_gt(bk, bj, thr, foff);
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, sup);
                    if (lane == 0 && bal) atomicOr(&s_sup, (u64)bal << ((warp & 1) * 32));
                }
            } else {
                // one problem over all classes' offset boxes: 16 kept-subsets x 64 candidates
                const int j = s0 + (tid & (NMS_G - 1));
                const bool valid = j < m;
                Box bj;
                bj.x1 = valid ? cx1[j] : 0.f;
                bj.y1 = valid ? cy1[j] : 0.f;
                bj.x2 = valid ? cx2[j] : 0.f;
                bj.y2 = valid ? cy2[j] : 0.f;
                bj.area = valid ? car[j] : 0.f;
                bool sup = false;
#pragma unroll 1
                for (int k = tid >> 6; k < nk; k += NMS_THREADS / NMS_G) {
                    Box bk;
                    bk.x1 = kx1[k];
                    bk.y1 = ky1[k];
                    bk.x2 = kx2[k];
                    bk.y2 = ky2[k];
                    bk.area = kar[k];
                    if (valid && !sup && iou_gt(bk, bj, thr, foff)) sup = true;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, sup);
                if (lane == 0 && bal) atomicOr(&s_sup, (u64)bal << ((warp & 1) * 32));
            }
            // ---- phase B1: suppression rows inside the group, row i = bits j > i that box i would suppress
#pragma unroll 1
            for (int i = warp; i < NMS_G; i += NMS_THREADS / 32) {
                const int gi = s0 + i;
                u64 row = 0ull;
                if (gi < m) {
                    Box bi;
                    bi.x1 = cx1[gi];
                    bi.y1 = cy1[gi];
                    bi.x2 = cx2[gi];
                    bi.y2 = cy2[gi];
                    bi.area = car[gi];
                    const int ci = ccl[gi];
                    unsigned w2[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int jj = h * 32 + lane, gj = s0 + jj;
                        bool hit = false;
                        if (jj > i && gj < m) {
                            Box bj;
                            bj.x1 = cx1[gj];
                            bj.y1 = cy1[gj];
                            bj.x2 = cx2[gj];
                            bj.y2 = cy2[gj];
                            bj.area = car[gj];
                            const bool same = !per_class || (ccl[gj] == ci);
                            hit = same && iou_gt(bi, bj, thr, foff);
                        }
                        w2[h] = __ballot_sync(0xffffffffu, hit);
                    }
                    row = ((u64)w2[1] << 32) | (u64)w2[0];
                }
                if (lane == 0) s_masks[i] = row;
            }
            __syncthreads();
#ifdef YPP_PROFILE
            long long pg_t1 = clock64();
            prof_ab1 += pg_t1 - pg_t0;
#endif
            // ---- phase B2: greedy scan over the group (warp 0): bit operations only
            if (warp == 0) {
                const int gcount = min(NMS_G, m - s0);
                const u64 validm = gcount >= 64 ? ~0ull : ((1ull << gcount) - 1ull);
                u64 alive = validm & ~s_sup;
                const u64 ma = s_masks[lane], mb = s_masks[lane + 32];
                const int room = cap - nk;
                // Only a row that is alive and whose mask meets an alive candidate can change anything, and `alive`
                // only shrinks: visit just those rows, in order. A candidate still alive at the end is kept.
                u64 nz = ((u64)__ballot_sync(0xffffffffu, (mb & alive) != 0ull) << 32) |
                         (u64)__ballot_sync(0xffffffffu, (ma & alive) != 0ull);
                nz &= alive;
#pragma unroll 1
                while (nz) {
                    const int i = __ffsll((long long)nz) - 1;
                    nz &= nz - 1ull;
                    const u64 mrow = i < 32 ? ma : mb;
                    const unsigned lo32 = __shfl_sync(0xffffffffu, (unsigned)mrow, i & 31);
                    const unsigned hi32 = __shfl_sync(0xffffffffu, (unsigned)(mrow >> 32), i & 31);
                    if ((alive >> i) & 1ull) alive &= ~(((u64)hi32 << 32) | (u64)lo32);
                }
                u64 keptm = alive;
                // only the first `room` kept boxes can enter the first `cap` kept (later ones never affect earlier)
#pragma unroll 1
                for (int kc = __popcll(keptm); kc > room; --kc) keptm &= ~(1ull << (63 - __clzll((long long)keptm)));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int jj = h * 32 + lane;
                    if ((keptm >> jj) & 1ull) {
                        const int kidx = nk + __popcll(keptm & ((1ull << jj) - 1ull));
                        const int gj = s0 + jj;
                        kx1[kidx] = cx1[gj];
                        ky1[kidx] = cy1[gj];
                        kx2[kidx] = cx2[gj];
                        ky2[kidx] = cy2[gj];
                        kar[kidx] = car[gj];
                        kcl[kidx] = ccl[gj];
                        kkey[kidx] = keys[gj];
                        knext[kidx] = atomicExch(&chead[ccl[gj]], kidx);  // push on the class chain (order irrelevant)
                    }
                }
                __syncwarp();  // every lane has read s_sup / s_masks before lane 0 resets them
                if (lane == 0) {
                    s_nk = nk + __popcll(keptm);
                    s_sup = 0ull;
                }
            }
            __syncthreads();
#ifdef YPP_PROFILE
            prof_b2 += clock64() - pg_t1;
#endif
        }
        processed += m;
        lo = keys[m - 1] + 1ull;
        __syncthreads();
    }
    YPP_PHASE(1, b, 5);
#ifdef YPP_PROFILE
    if (tid == 0 && b < 256) {
        g_phase[1][b][7] = prof_chunks;
        g_phase[1][b][8] = prof_groups;
        g_phase[1][b][9] = ntot;
        g_phase[1][b][10] = prof_ab1;
        g_phase[1][b][11] = prof_b2;
    }
#endif
    // outputs: dets = (boxes[keep], scores[keep]), labels[keep]  (bbox_nms.py:84-93)
    int nk = s_nk;
    // "keep all" (no max_num) but the kept list is full while candidates remain: report, do not truncate silently
    if (P.m_eff <= 0 && nk >= cap && processed < ntot && tid == 0 && P.o_status) atomicMax(P.o_status, 3);
    if (nk > P.out_cap) {
        nk = P.out_cap;
        if (tid == 0 && P.o_status) atomicMax(P.o_status, 3);  // YOLOPP_E_OVERFLOW
    }
    const bool grouped = P.o_cls_offsets && !generic;
    int* gcnt = reinterpret_cast<int*>(ktmp);  // (the select scratch is free by now: C <= 4096 ints)
    if (grouped) {
        __syncthreads();
        nms_label_offsets(P, b, nk, kcl, gcnt);
    }
#pragma unroll 1
    for (int i = tid; i < nk; i += NMS_THREADS) {
        const u64 key = kkey[i];
        const uint32_t flat = key_flat(key);
        const int r = (int)(flat / (uint32_t)C);
        const int c = generic ? kcl[i] : (int)(flat - (uint32_t)r * (uint32_t)C);
        const float4 bx = row_box[P.boxes_per_class ? flat : (uint32_t)r];
        const float sc = key_score(key);
        float* d = P.o_dets + ((size_t)b * P.out_cap + i) * 5;
        d[0] = bx.x;
        d[1] = bx.y;
        d[2] = bx.z;
        d[3] = bx.w;
        d[4] = sc;
        if (P.o_labels) P.o_labels[(size_t)b * P.out_cap + i] = (long long)c;
        if (P.o_keep) P.o_keep[(size_t)b * P.out_cap + i] = (long long)flat;
        if (P.o_anchors) P.o_anchors[(size_t)b * P.out_cap + i] = P.row_anchor[(size_t)b * P.R + r];
        if (P.o_rows) P.o_rows[(size_t)b * P.out_cap + i] = r;
        if (grouped) {
            // stable position inside the label's group: kept boxes of the class with a smaller output index
            int pos = gcnt[c];
#pragma unroll 1
            for (int k = chead[c]; k >= 0; k = knext[k]) pos += (k < i) ? 1 : 0;
            float* g = P.o_cls_dets + ((size_t)b * P.out_cap + pos) * 5;
            g[0] = bx.x;
            g[1] = bx.y;
            g[2] = bx.z;
            g[3] = bx.w;
            g[4] = sc;
        }
    }
    if (tid == 0) P.o_count[b] = nk;
    YPP_PHASE(1, b, 6);
}

// One CTA per image — or, when the grid is smaller than the batch (batches in flight on several streams: what counts
// there is SM-time, not latency), a CTA takes several images one after the other: the later ones find the kernel's
// code in the instruction caches, which is most of what a per-image pass waits for.
__global__ void __launch_bounds__(NMS_THREADS, 1) nms_image_kernel(const __grid_constant__ DevParams P) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    __shared__ unsigned long long s_stage_bar;  // bulk copies of the candidate scan
    uint64_t* stage_bar = reinterpret_cast<uint64_t*>(&s_stage_bar);
    uint32_t stage_phase = 0u;
    if (threadIdx.x == 0) {
        mbar_init(stage_bar, 1);
        fence_mbar_init();
    }
    // (the body's first block barrier publishes the initialised mbarrier)
#pragma unroll 1
    for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
        nms_image_body(P, b, nms_smem, stage_bar, stage_phase);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// multiclass_nms front half (bbox_nms.py:34-62): threshold, optional score_factors, per-row statistics.
// One warp per row of multi_scores (n, C+1); writes the (row, class) score matrix the NMS kernel streams.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) multiclass_prep_kernel(const float* __restrict__ multi_scores,
                                                              const float* __restrict__ score_factors,
                                                              const float4* __restrict__ boxes, int per_class, int n,
                                                              int C, float score_thr, uint32_t* __restrict__ mat,
                                                              uint4* __restrict__ row_stat) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const float* srow = multi_scores + (size_t)r * (C + 1);  // last column = background, ignored (bbox_nms.py:42)
    const float fac = score_factors ? score_factors[r] : 1.0f;
    uint32_t best = 0u, worst = 0u, mx = 0u;
    int npass = 0;
    for (int c = lane; c < C; c += 32) {
        float sc = srow[c];
        const bool pass = sc > score_thr;             // valid_mask = scores > score_thr  (bbox_nms.py:54)
        if (score_factors) sc = fmul(sc, fac);        // scores * score_factors AFTER the threshold (:57-62)
        mat[(size_t)r * C + c] = pass ? __float_as_uint(sc) : SCORE_NONE;
        if (pass) {
            const uint32_t o = f2ord(sc);
            best = o > best ? o : best;
            worst = ~o > worst ? ~o : worst;
            ++npass;
            if (per_class) {
                const uint32_t bo = f2ord(box_max(boxes[(size_t)r * C + c]));
                mx = bo > mx ? bo : mx;
            }
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    worst = __reduce_max_sync(0xffffffffu, worst);
    mx = __reduce_max_sync(0xffffffffu, mx);
    npass = __reduce_add_sync(0xffffffffu, npass);
    if (lane == 0) row_stat[r] = make_uint4(best, worst, (uint32_t)npass, mx);
}

// ------------------------------------------------------------------------------------------------
// small standalone kernels
// ------------------------------------------------------------------------------------------------
// yolopp_topk_conf: rows of the segments that keep every anchor (no top-k) are the anchors in order
__global__ void fill_rows_kernel(const __grid_constant__ DevParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)P.B * P.R) return;
    const int r = (int)(i % P.R);
    int s = 0;
    for (int q = 1; q < P.nsegs; ++q)
        if (r >= P.seg[q].row_off) s = q;
    const SegDev& sg = P.seg[s];
    if (!sg.has_topk) P.row_anchor[i] = P.lv[sg.first_level].n_off + (r - sg.row_off);
}

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
