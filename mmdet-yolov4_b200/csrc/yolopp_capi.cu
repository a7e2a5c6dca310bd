// yolopp_capi.cu — the C ABI of include/yolopp.h: validation, workspace plan, TMA descriptors, launches.
// Pure CUDA runtime + one driver entry point (cuTensorMapEncodeTiled, resolved through the runtime so the
// library does not link libcuda). No torch, no device allocation, no host synchronisation.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "../../include/yolopp.h"
#include "yolopp_kernels.cuh"
#include "yolopp_mish.cuh"

#ifndef YPP_STATIC_NUM
#define YPP_STATIC_NUM 3  // statically dealt share of the decode kernel's tile sequence (batch alone)
#define YPP_STATIC_DEN 4
#endif

using namespace ypp;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr size_t SMEM_LIMIT_TOTAL = 232448;  // static + dynamic shared memory a CTA may use on sm_100 (227 KB opt-in)
#ifndef YPP_NMS_STATIC_SMEM
#define YPP_NMS_STATIC_SMEM 10752  // >= the per-image kernels' static shared memory (ptxas -v; checked at start-up)
#endif
constexpr size_t SMEM_LIMIT = SMEM_LIMIT_TOTAL - YPP_NMS_STATIC_SMEM;  // dynamic part

// Staging buffer of the NMS kernel's candidate scan: whole score-matrix rows, one bulk copy each (rows must be
// 16-byte multiples; classes <= 128). The NMS CTA owns its SM anyway (512 threads x 128 registers), so the buffer takes
// what the other arrays leave — the more rows fit, the more often the scan is a single DRAM round trip.
static size_t nms_add_stage(ypp::DevParams& d, size_t smem) {
    d.nms_stage_off = 0;
    d.nms_stage_rows = 0;
#ifdef YPP_NO_STAGE
    return smem;
#endif
    if (d.C % 4 != 0 || d.C > 128 || d.generic) return smem;
    smem = (smem + 127) & ~(size_t)127;
    if (smem >= SMEM_LIMIT) return smem;
    int rows = (int)((SMEM_LIMIT - smem) / ((size_t)d.C * 4));
    if (rows > 1024) rows = 1024;
    if (rows < 16) return smem;
    d.nms_stage_off = (int)smem;
    d.nms_stage_rows = rows;
    return smem + (size_t)rows * d.C * 4;
}

struct Plan {
    DevParams d;
    size_t off_counters, counters_bytes;
    size_t off_ckey, off_rank, off_row_anchor, off_row_box, off_row_stat, off_mat, off_kept, kept_stride;
    size_t total;
    size_t dec_smem, sel_smem, nms_smem;
    int dec_ctas_per_sm;
    int rows_blocks;  // NHWC: blocks of the row-driven decode kernel
};

#define CHECK_ARG(cond) \
    do {                \
        if (!(cond)) return false; \
    } while (0)

// Validates the params and lays the workspace out. Level pointers are filled in by the caller.
bool make_plan(const yolopp_params* p, Plan* plan) {
    CHECK_ARG(p != nullptr);
    CHECK_ARG(p->abi_version == YOLOPP_ABI_VERSION);
    CHECK_ARG(p->mode == YOLOPP_MODE_CSP || p->mode == YOLOPP_MODE_V3);
    CHECK_ARG(p->layout == YOLOPP_LAYOUT_NCHW || p->layout == YOLOPP_LAYOUT_NHWC);
    CHECK_ARG(p->batch >= 1 && p->batch <= 65535);
    CHECK_ARG(p->num_levels >= 1 && p->num_levels <= YOLOPP_MAX_LEVELS);
    CHECK_ARG(p->num_anchors >= 1 && p->num_anchors <= YOLOPP_MAX_ANCHORS);
    CHECK_ARG(p->nms_offset == 0 || p->nms_offset == 1);
    const bool agn = p->class_agnostic != 0;
    CHECK_ARG(!(agn && p->mode == YOLOPP_MODE_V3));  // YOLOV3Head has no class-agnostic variant
    const int C = agn ? 1 : p->num_classes;
    CHECK_ARG(C >= 1 && C <= YOLOPP_MAX_CLASSES);
    const int NA = agn ? 5 : 5 + p->num_classes;

    memset(plan, 0, sizeof(*plan));
    DevParams& d = plan->d;
    d.mode = p->mode;
    d.B = p->batch;
    d.L = p->num_levels;
    d.A = p->num_anchors;
    d.C = C;
    d.NA = NA;
    d.agnostic = agn ? 1 : 0;
    d.nhwc = p->layout == YOLOPP_LAYOUT_NHWC ? 1 : 0;
    d.score_thr = p->score_thr;
    d.conf_thr = p->conf_thr;
    d.iou_thr = p->iou_thr;
    d.nms_score_thr = p->nms_score_thr;
    d.foff = (float)p->nms_offset;
    d.split_thr = p->split_thr;
    d.nms_agnostic = p->nms_class_agnostic ? 1 : 0;
    d.rescale = p->rescale ? 1 : 0;

    long long n_off = 0, m_off = 0;
    for (int l = 0; l < d.L; ++l) {
        CHECK_ARG(p->height[l] >= 1 && p->width[l] >= 1 && p->height[l] <= 16384 && p->width[l] <= 16384);
        CHECK_ARG(p->stride_w[l] >= 1 && p->stride_h[l] >= 1 && p->coder_stride[l] >= 1);
        LevelDev& lv = d.lv[l];
        lv.H = p->height[l];
        lv.W = p->width[l];
        long long hw = (long long)lv.H * lv.W;
        CHECK_ARG(hw * d.A < (1 << 24));
        lv.HW = (int)hw;
        lv.n_off = (int)n_off;
        lv.m_off = (int)m_off;
        lv.sx = (float)p->stride_w[l];
        lv.sy = (float)p->stride_h[l];
        lv.cstride = (float)p->coder_stride[l];
        lv.inv_w = 1.0f / (float)lv.W;
        for (int a = 0; a < d.A; ++a)
            for (int k = 0; k < 4; ++k) lv.base[a][k] = p->base_anchors[l][a][k];
        n_off += hw * d.A;
        m_off = (long long)align_up((size_t)(m_off + hw * d.A), 4);
        CHECK_ARG(n_off < (1 << 24));
    }
    d.N = (int)n_off;
    d.M_pad = (int)m_off + 72;  // slack: the rank row of a partial last tile is fetched whole (+ aligned superset)

    // top-k segments: CSP = one over all levels (yolocsp_head.py:350-355); V3 = one per level (yolo_head.py:281-302)
    d.nsegs = (p->mode == YOLOPP_MODE_CSP) ? 1 : d.L;
    long long row_off = 0;
    d.ntopk = 0;
    for (int s = 0; s < d.nsegs; ++s) {
        SegDev& sg = d.seg[s];
        sg.first_level = (p->mode == YOLOPP_MODE_CSP) ? 0 : s;
        sg.num_levels = (p->mode == YOLOPP_MODE_CSP) ? d.L : 1;
        long long n = 0;
        for (int q = 0; q < sg.num_levels; ++q) {
            n += (long long)d.lv[sg.first_level + q].HW * d.A;
            d.lv[sg.first_level + q].seg = s;
        }
        sg.N = (int)n;
        sg.m_begin = d.lv[sg.first_level].m_off;
        sg.m_end = d.lv[sg.first_level + sg.num_levels - 1].m_off + d.lv[sg.first_level + sg.num_levels - 1].HW * d.A;
        sg.has_topk = (p->nms_pre > 0 && p->nms_pre < n) ? 1 : 0;  // any k: beyond SEL_MAX_K the sort runs in chunks
        sg.k = sg.has_topk ? p->nms_pre : (int)n;
        if (sg.has_topk) d.topk_segs[d.ntopk++] = s;
        sg.row_off = (int)row_off;
        row_off += sg.k;
    }
    CHECK_ARG(row_off <= YOLOPP_MAX_ROWS);
    d.R = (int)row_off;

    int m_eff = -1;
    if (p->max_per_img > 0) m_eff = p->max_per_img;
    if (p->nms_max_num > 0) m_eff = (m_eff > 0 && m_eff < p->nms_max_num) ? m_eff : p->nms_max_num;
    d.m_eff = m_eff;
    d.out_cap = p->out_capacity > 0 ? p->out_capacity : p->max_per_img;
    CHECK_ARG(d.out_cap >= 1 && d.out_cap <= YOLOPP_MAX_ROWS);
    CHECK_ARG(m_eff <= YOLOPP_MAX_ROWS);
    // boxes the NMS pass may keep: max_num when given, else one more than fits (to flag the overflow). Up to
    // NMS_MAX_KEEP the kept list lives in shared memory, beyond that in the workspace.
    d.keep_cap = m_eff > 0 ? m_eff : d.out_cap + 1;
    const int cap_s = d.keep_cap > NMS_MAX_KEEP ? 0 : d.keep_cap;
    plan->kept_stride = d.keep_cap > NMS_MAX_KEEP ? align_up((size_t)d.keep_cap * 36, 256) : 0;
    CHECK_ARG((long long)d.R * d.C < (1ll << 31));
    int max_k = 0;
    for (int s = 0; s < d.nsegs; ++s)
        if (d.seg[s].has_topk && d.seg[s].k > max_k) max_k = d.seg[s].k;
    int kcap = 4096;  // >= 2k: room for the survivors of the sampled pivot (~1.5k expected)
    while (kcap < 2 * max_k && kcap < 2 * SEL_MAX_K) kcap <<= 1;
    d.sel_kcap = kcap;  // (nms_pre > SEL_MAX_K: chunks of kcap / 2 sorted keys)
    plan->sel_smem = (size_t)kcap * 8 * 2;  // sorted keys + scatter scratch
    d.sel_stage = 0;
    d.sel_stride = 1;
    if (max_k <= SEL_MAX_K) {
        // staging buffer of the fast path: the largest top-k segment's objectness logits when they fit (<= 32768
        // slots), else a 1-in-2^j sample of them (<= 16384 slots) and a streamed pass over the segment
        int max_m = 0;
        for (int s = 0; s < d.nsegs; ++s)
            if (d.seg[s].has_topk && d.seg[s].m_end - d.seg[s].m_begin > max_m) max_m = d.seg[s].m_end - d.seg[s].m_begin;
        if (max_m > 0) {
            int slots = max_m;
            if (!(max_m <= 32768 && plan->sel_smem + (size_t)max_m * 4 <= 200 * 1024)) {
                int stride = 2;
                while ((max_m + stride - 1) / stride + 4 * d.L * d.A > 16384) stride <<= 1;
                d.sel_stride = stride;
                slots = (max_m + stride - 1) / stride + 4 * d.L * d.A;  // (+ per-plane alignment padding of the sample)
            }
            d.sel_stage = slots;
            plan->sel_smem += (size_t)slots * 4;
        }
    }
    CHECK_ARG(plan->sel_smem <= SMEM_LIMIT);
    plan->nms_smem = (size_t)NMS_KCAP * 8 * 2 + (size_t)cap_s * 8 + (size_t)NMS_CH * 24 + (size_t)cap_s * 28 +
                     (size_t)d.C * 4;
    plan->nms_smem = align_up(plan->nms_smem, 16);
    d.nms_rowkeys_off = (int)plan->nms_smem;
    plan->nms_smem += (size_t)NMS_KCAP * 8;
    plan->nms_smem = nms_add_stage(d, plan->nms_smem);
    CHECK_ARG(plan->nms_smem <= SMEM_LIMIT);

    // which kernel decodes which level
    d.dec_quad = 0;
    int tma_tiles = 0, ldg_blocks = 0, dense_tiles = 0;
    if (d.nhwc) {
        // channels-last: the row-driven kernel serves every level
        for (int l = 0; l < d.L; ++l) d.lv[l].use_tma = 0;
        plan->rows_blocks = (int)(((long long)d.B * d.R + ROWS_WARPS - 1) / ROWS_WARPS);
    } else {
        bool tma_fits = NA <= 256;
#ifdef YPP_QUAD
        for (int l = 0; l < d.L && tma_fits; ++l) {
            const SegDev& sg = d.seg[d.lv[l].seg];
            if (sg.has_topk && (long long)sg.k * 4 <= sg.N && d.lv[l].HW % 4 != 0) d.dec_quad = 1;
        }
#endif
        // tile size of the persistent kernel: 32 positions x 8 stages where a 64-position tile would hold 2 or more
        // admitted anchors on average (the in-order ring then stalls on the tiles with many), 64 x 4 where tiles are
        // mostly empty and the per-tile costs dominate (see dec_stages_of). Quad-row tiles exist for 64 only.
        {
            double adm = 0.0, pos = 0.0;
            for (int l = 0; l < d.L; ++l) {
                const SegDev& sg = d.seg[d.lv[l].seg];
                if (!(sg.has_topk && (long long)sg.k * 4 <= sg.N)) continue;
                adm += (double)sg.k * d.lv[l].HW * d.A / sg.N;  // the level's share of its segment's top-k (expected)
                pos += (double)d.lv[l].HW * d.A;
            }
            d.tile_t = (!d.dec_quad && pos > 0.0 && 64.0 * adm / pos >= 2.0) ? 32 : 64;
#ifdef YPP_TILE64
            d.tile_t = 64;
#endif
#ifdef YPP_TILE32
            if (!d.dec_quad) d.tile_t = 32;
#endif
        }
        StageGeom geom = stage_geom(NA, d.dec_quad, d.tile_t);
        plan->dec_smem = 1024 /*alignment slack*/ + 1024 /*barriers*/ + (size_t)dec_stages_of(d.tile_t) * geom.stage_bytes;
        tma_fits = tma_fits && plan->dec_smem <= 200 * 1024;
        plan->dec_ctas_per_sm = 2 * (plan->dec_smem + 1024) <= 233472 ? 2 : 1;  // (228 KB per SM, 1 KB reserved per CTA)
        // persistent decode kernel: sparse-admission levels — TMA tiles where the plane stride is 16-byte aligned,
        // gather tiles where it is not (quad-row TMA tiles in the -DYPP_QUAD build: measured slower)
        for (int l = 0; l < d.L; ++l) {
            LevelDev& lv = d.lv[l];
            const SegDev& sg = d.seg[lv.seg];
            const bool sparse = sg.has_topk && (long long)sg.k * 4 <= sg.N;
#ifdef YPP_QUAD
            lv.use_tma = (tma_fits && sparse) ? ((lv.HW % 4 == 0) ? 1 : 3) : 0;  // (opt-in build: quad-row TMA tiles)
#else
            lv.use_tma = (tma_fits && sparse) ? ((lv.HW % 4 == 0) ? 1 : 2) : 0;
#endif
            lv.dense = (!lv.use_tma && !sparse) ? 1 : 0;
            lv.qrows = (int)(((long long)d.B * d.A * NA) / 4);
        }
        // tile enumeration of the persistent kernel: unaligned levels (quad-row / gather tiles) first — they are the
        // small, coarse levels — then the plain TMA levels
        for (int pi = 0; pi < 2; ++pi) {
            for (int l = 0; l < d.L; ++l) {
                LevelDev& lv = d.lv[l];
#ifdef YPP_QUAD_LAST
                const bool mine = pi == 0 ? lv.use_tma == 1 : (lv.use_tma == 2 || lv.use_tma == 3);
#else
                const bool mine = pi == 0 ? (lv.use_tma == 2 || lv.use_tma == 3) : lv.use_tma == 1;
#endif
                if (!mine) continue;
                lv.tile_pos = lv.use_tma == 2 ? GATHER_T : d.tile_t;  // (a level that turns into a gather level later —
                lv.tpp = (lv.HW + lv.tile_pos - 1) / lv.tile_pos;     //  unaligned pointer — keeps the kernel's tile size)
                lv.tile0 = tma_tiles;
                long long t = (long long)lv.tpp * d.B * d.A;
                CHECK_ARG(tma_tiles + t < (1ll << 30));
                tma_tiles += (int)t;
            }
        }
        for (int l = 0; l < d.L; ++l) {
            LevelDev& lv = d.lv[l];
            if (lv.use_tma) continue;
            if (lv.dense) {
                lv.tpp = (lv.HW + 31) / 32;
                lv.tile0 = dense_tiles;
                long long t = (long long)lv.tpp * d.B * d.A;
                CHECK_ARG(dense_tiles + t < (1ll << 30));
                dense_tiles += (int)t;
            } else {
                lv.tpp = (lv.HW + 127) / 128;
                lv.tile0 = ldg_blocks;
                long long t = (long long)lv.tpp * d.B * d.A;
                CHECK_ARG(ldg_blocks + t < (1ll << 30));
                ldg_blocks += (int)t;
            }
        }
    }
    d.tma_tiles = tma_tiles;
    d.ldg_blocks = ldg_blocks;
    d.dense_tiles = dense_tiles;

    // workspace layout
    size_t off = 0;
    const size_t B = (size_t)d.B, Cc = (size_t)d.C, R = (size_t)d.R;
    plan->off_counters = off;
    plan->counters_bytes = 256;  // tile counter of the decode kernel's scheduler
    off += plan->counters_bytes;
    plan->off_ckey = off;
    off += align_up(B * d.M_pad * 8, 256);
    plan->off_rank = off;
    off += align_up(B * d.M_pad * 4, 256);
    plan->off_row_anchor = off;
    off += align_up(B * R * 4, 256);
    plan->off_row_box = off;
    off += align_up(B * R * 16, 256);
    plan->off_row_stat = off;
    off += align_up(B * R * 16, 256);
    plan->off_mat = off;
    off += align_up(B * R * Cc * 4, 256);
    plan->off_kept = off;
    off += B * plan->kept_stride;
    plan->total = off;
    return true;
}

void bind_workspace(Plan* plan, void* ws) {
    unsigned char* w = (unsigned char*)ws;
    DevParams& d = plan->d;
    d.tile_ctr = (unsigned*)(w + plan->off_counters);
    d.ckey = (u64*)(w + plan->off_ckey);
    d.rank = (uint32_t*)(w + plan->off_rank);
    d.row_anchor = (int*)(w + plan->off_row_anchor);
    d.row_box = (float4*)(w + plan->off_row_box);
    d.row_stat = (uint4*)(w + plan->off_row_stat);
    d.mat = (uint32_t*)(w + plan->off_mat);
    d.nms_kept = plan->kept_stride ? w + plan->off_kept : nullptr;
    d.nms_kept_stride = (long long)plan->kept_stride;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static std::atomic<EncodeTiledFn> fn{nullptr};  // resolved once; the driver symbol never changes
    EncodeTiledFn f = fn.load(std::memory_order_acquire);
    if (f) return f;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
    if (qres != cudaDriverEntryPointSuccess) return nullptr;
    fn.store((EncodeTiledFn)sym, std::memory_order_release);
    return (EncodeTiledFn)sym;
}

inline int cuda_rc(cudaError_t e) { return e == cudaSuccess ? YOLOPP_OK : YOLOPP_E_CUDA + (int)e; }

// Per-device, once per process: sm_100 check, SM count, and the opt-in to large dynamic shared memory for the
// kernels that need it. The attribute is set to the DEVICE MAXIMUM (not a call's own size), so concurrent callers
// with different configurations can never lower it under each other; the only process-wide state of the library
// is this idempotent initialisation.
struct DeviceInfo {
    int rc;
    int sms;
};
DeviceInfo device_info() {
    static std::atomic<int> state[64];  // 0: unknown, else (sms << 8) | 1
    DeviceInfo di = {YOLOPP_E_NO_DEVICE, 148};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return di;
    if (dev >= 0 && dev < 64) {
        const int st = state[dev].load(std::memory_order_acquire);
        if (st) {
            di.rc = YOLOPP_OK;
            di.sms = st >> 8;
            return di;
        }
    }
    int major = 0, sms = 148, optin = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return di;
    if (major != 10) return di;  // sm_100a only: no other code path exists
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // (the opt-in limit covers static + dynamic shared memory of a kernel)
    auto opt_in = [&](const void* fn) -> cudaError_t {
        cudaFuncAttributes fa;
        cudaError_t e2 = cudaFuncGetAttributes(&fa, fn);
        if (e2 != cudaSuccess) return e2;
        if (fa.sharedSizeBytes > (size_t)YPP_NMS_STATIC_SMEM) return cudaErrorInvalidConfiguration;  // SMEM_LIMIT assumes it
        return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
    };
    cudaError_t e = opt_in((const void*)select_kernel);
    if (e == cudaSuccess) e = opt_in((const void*)decode_tma_kernel<0, 64>);
    if (e == cudaSuccess) e = opt_in((const void*)decode_tma_kernel<1, 64>);
    if (e == cudaSuccess) e = opt_in((const void*)decode_tma_kernel<0, 32>);
    if (e == cudaSuccess) e = opt_in((const void*)decode_tma_kernel<1, 32>);
    if (e == cudaSuccess) e = opt_in((const void*)nms_image_kernel);
    if (e != cudaSuccess) {
        di.rc = cuda_rc(e);
        return di;
    }
    if (dev >= 0 && dev < 64) state[dev].store((sms << 8) | 1, std::memory_order_release);
    di.rc = YOLOPP_OK;
    di.sms = sms;
    return di;
}

}  // namespace

// Everything a call needs, derived once: the caller-visible handle of yolopp_plan_create and the stack object of
// the one-shot entry points.
struct yolopp_plan {
    Plan plan;
    TmapPack maps;
    int dec_grid;
    int nms_grid;    // CTAs of the NMS kernel: one per image, or one per YPP_NMS_IMAGES_PER_CTA images for batches in flight
    int stop_after;  // 0: whole path, 1: after the top-k, 2: after the decode (stage entries)
    cudaGraphExec_t exec;  // plan handles only: the call's launches as one executable graph (null: launch directly)
};

namespace {

int prepare(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors, const yolopp_outputs* out,
            void* workspace, size_t workspace_bytes, int stop_after, yolopp_plan* pl) {
    Plan& plan = pl->plan;
    if (!make_plan(p, &plan)) return YOLOPP_E_INVALID;
    if (!level_ptrs) return YOLOPP_E_INVALID;
    if (stop_after == 0 && (!out || !out->dets || !out->labels || !out->count || !out->status)) return YOLOPP_E_INVALID;
    if (out && ((out->cls_dets == nullptr) != (out->cls_offsets == nullptr))) return YOLOPP_E_INVALID;
    if (p->rescale && !scale_factors && stop_after != 1) return YOLOPP_E_INVALID;
    if (!workspace || workspace_bytes < plan.total) return YOLOPP_E_WORKSPACE;
    if (((uintptr_t)workspace & 255) != 0) return YOLOPP_E_INVALID;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    pl->stop_after = stop_after;
    pl->exec = nullptr;
    DevParams& d = plan.d;
    for (int l = 0; l < d.L; ++l) {
        if (!level_ptrs[l] || ((uintptr_t)level_ptrs[l] & 3) != 0) return YOLOPP_E_INVALID;
        d.lv[l].ptr = level_ptrs[l];
        // a tensor map needs a 16-byte aligned base: otherwise the level's tiles are gathered
        if (d.lv[l].use_tma && ((uintptr_t)level_ptrs[l] & 15) != 0) d.lv[l].use_tma = 2;
    }
    // top-k staging: which levels' objectness planes can be bulk-copied (16-byte aligned source and destination)
    for (int sgi = 0; sgi < d.nsegs; ++sgi) {
        SegDev& sg = d.seg[sgi];
        sg.sel_bulk_bytes = 0;
        for (int q = 0; q < sg.num_levels; ++q) {
            LevelDev& lv = d.lv[sg.first_level + q];
            lv.sel_bulk = (!d.nhwc && (lv.HW & 3) == 0 && ((uintptr_t)lv.ptr & 15) == 0 && ((lv.m_off - sg.m_begin) & 3) == 0) ? 1 : 0;
            if (lv.sel_bulk) sg.sel_bulk_bytes += (unsigned)(d.A * lv.HW) * 4u;
        }
    }
    bind_workspace(&plan, workspace);
    d.scale = scale_factors;
    if (out) {
        d.o_dets = out->dets;
        d.o_labels = (long long*)out->labels;
        d.o_anchors = out->anchors;
        d.o_rows = out->rows;
        d.o_count = out->count;
        d.o_ncand = out->num_candidates;
        d.o_status = out->status;
        d.o_cls_dets = out->cls_dets;
        d.o_cls_offsets = out->cls_offsets;
    }
    // decode kernel's grid and the statically dealt head of its tile sequence (select_kernel initialises the
    // scheduler counter with it). A batch that runs alone: 3/4 of the tiles, a whole number of rounds over the
    // CTAs, the rest is claimed dynamically so that all CTAs run dry together. Batches in flight on several
    // streams (params.batches_in_flight > 1): everything is dealt statically — CTAs then retire progressively
    // and the per-image kernels of the neighbouring batches move onto the freed SMs while the rest of the decode
    // kernel still streams.
    int dec_grid = di.sms * plan.dec_ctas_per_sm;
#ifdef YPP_EXPERIMENT_KNOBS
    if (getenv("YPP_DEC_GRID")) dec_grid = atoi(getenv("YPP_DEC_GRID"));
#endif
    if (dec_grid > d.tma_tiles) dec_grid = d.tma_tiles;
    pl->dec_grid = dec_grid;
#ifndef YPP_NMS_IMAGES_PER_CTA
#define YPP_NMS_IMAGES_PER_CTA 1  // measured (608^2 b64, 6 batches in flight): 1 -> 579 k, 2 -> 583 k, 4 -> 575 k img/s: no gain
#endif
    {
        const char* env = getenv("YPP_NMS_IPC");  // (experiment knob)
        const int ipc = env ? atoi(env) : YPP_NMS_IMAGES_PER_CTA;
        pl->nms_grid = (p->batches_in_flight > 1 && ipc > 1) ? (d.B + ipc - 1) / ipc : d.B;
    }
    d.dec_first = 0;
    if (d.tma_tiles > 0) {
        d.dec_first = p->batches_in_flight > 1 ? (unsigned)d.tma_tiles
                                               : (unsigned)(((long long)d.tma_tiles * YPP_STATIC_NUM / YPP_STATIC_DEN) / dec_grid) * (unsigned)dec_grid;
        // the persistent kernel only takes levels whose segment ran a top-k, so select_kernel (which initialises
        // the scheduler counter) runs
        if (d.ntopk == 0) return YOLOPP_E_INVALID;
        EncodeTiledFn enc = get_encode_fn();
        if (!enc) return YOLOPP_E_NO_DEVICE;
        memset(&pl->maps, 0, sizeof(pl->maps));
        const StageGeom geom = stage_geom(d.NA, d.dec_quad, d.tile_t);
        for (int l = 0; l < d.L; ++l) {
            const LevelDev& lv = d.lv[l];
            CUresult r = CUDA_SUCCESS;
            if (lv.use_tma == 1) {
                // rows = planes, one (NA x 32) box = all attributes of 32 positions
                cuuint64_t gdim[2] = {(cuuint64_t)lv.HW, (cuuint64_t)d.B * d.A * d.NA};
                cuuint64_t gstr[1] = {(cuuint64_t)lv.HW * 4};
                cuuint32_t box[2] = {(cuuint32_t)TILE_SUB, (cuuint32_t)d.NA};
                cuuint32_t estr[2] = {1, 1};
                r = enc(&pl->maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)lv.ptr, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            } else if (lv.use_tma == 3) {
                // plane stride 4 * HW bytes is not a multiple of 16: view the tensor as rows of FOUR planes (stride
                // 16 * HW bytes); a box = the aligned superset of 64 positions of one plane-in-row x the rows a slab spans
                if (lv.qrows < 1) {
                    d.lv[l].use_tma = 2;
                    continue;
                }
                cuuint64_t gdim[2] = {(cuuint64_t)lv.HW * 4, (cuuint64_t)lv.qrows};
                cuuint64_t gstr[1] = {(cuuint64_t)lv.HW * 16};
                cuuint32_t box[2] = {(cuuint32_t)QUAD_W, (cuuint32_t)geom.quad_rows};
                cuuint32_t estr[2] = {1, 1};
                r = enc(&pl->maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)lv.ptr, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
            if (r != CUDA_SUCCESS) return YOLOPP_E_CUDA + 999;
        }
    }
    return YOLOPP_OK;
}

int launch(const yolopp_plan* pl, cudaStream_t stream, void* const* events, int num_events) {
    const Plan& plan = pl->plan;
    const DevParams& d = plan.d;
    cudaError_t e;
    int ev_i = 0;
#define YPP_MARK()                                                                          \
    do {                                                                                    \
        if (events && ev_i < num_events) {                                                  \
            e = cudaEventRecord((cudaEvent_t)events[ev_i++], stream);                       \
            if (e != cudaSuccess) return cuda_rc(e);                                        \
        }                                                                                   \
    } while (0)
    YPP_MARK();  // 0: start of select
#ifdef YPP_EXPERIMENT_KNOBS
    const char* skip_env = getenv("YPP_SKIP");  // timing experiments only: bit 0 = no top-k, bit 1 = no NMS (stale results)
    const int skip = skip_env ? atoi(skip_env) : 0;
#else
    const int skip = 0;
#endif
    if (d.ntopk > 0 && !(skip & 1)) {
        // (block (0, 0) also clears the status word and initialises the decode kernel's tile counter)
        select_kernel<<<dim3(d.ntopk, d.B), SEL_THREADS, plan.sel_smem, stream>>>(d);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
    } else if (d.o_status) {
        e = cudaMemsetAsync(d.o_status, 0, sizeof(int32_t), stream);
        if (e != cudaSuccess) return cuda_rc(e);
    }
    if (pl->stop_after == 1) {
        // rows of the segments that keep every anchor (decode would write them): the anchors in order
        if (d.ntopk < d.nsegs) {
            const long long n = (long long)d.B * d.R;
            fill_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d);
            if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
        }
        return YOLOPP_OK;
    }
    YPP_MARK();  // 1: start of decode (persistent TMA kernel / NHWC rows kernel)
    if (d.tma_tiles > 0) {
        if (d.mode == YOLOPP_MODE_CSP) {
            if (d.tile_t == 64) decode_tma_kernel<0, 64><<<pl->dec_grid, DEC_THREADS, plan.dec_smem, stream>>>(d, pl->maps);
            else decode_tma_kernel<0, 32><<<pl->dec_grid, DEC_THREADS, plan.dec_smem, stream>>>(d, pl->maps);
        } else {
            if (d.tile_t == 64) decode_tma_kernel<1, 64><<<pl->dec_grid, DEC_THREADS, plan.dec_smem, stream>>>(d, pl->maps);
            else decode_tma_kernel<1, 32><<<pl->dec_grid, DEC_THREADS, plan.dec_smem, stream>>>(d, pl->maps);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
    }
    if (d.nhwc) {
        if (d.mode == YOLOPP_MODE_CSP)
            decode_rows_kernel<0><<<plan.rows_blocks, 32 * ROWS_WARPS, 0, stream>>>(d);
        else
            decode_rows_kernel<1><<<plan.rows_blocks, 32 * ROWS_WARPS, 0, stream>>>(d);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
    }
    YPP_MARK();  // 2: start of decode (other levels)
    if (d.dense_tiles > 0) {
        const int blocks = (d.dense_tiles + DENSE_WARPS - 1) / DENSE_WARPS;
        if (d.mode == YOLOPP_MODE_CSP)
            decode_dense_kernel<0><<<blocks, 32 * DENSE_WARPS, 0, stream>>>(d);
        else
            decode_dense_kernel<1><<<blocks, 32 * DENSE_WARPS, 0, stream>>>(d);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
    }
    if (d.ldg_blocks > 0) {
        if (d.mode == YOLOPP_MODE_CSP)
            decode_ldg_kernel<0><<<d.ldg_blocks, 128, 0, stream>>>(d);
        else
            decode_ldg_kernel<1><<<d.ldg_blocks, 128, 0, stream>>>(d);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
    }
    if (pl->stop_after == 2) return YOLOPP_OK;
    YPP_MARK();  // 3: start of the per-image NMS
    if (!(skip & 2)) {
        nms_image_kernel<<<pl->nms_grid, NMS_THREADS, plan.nms_smem, stream>>>(d);
        if ((e = cudaGetLastError()) != cudaSuccess) return cuda_rc(e);
    }
    YPP_MARK();  // 4: end
#undef YPP_MARK
    return YOLOPP_OK;
}

}  // namespace

extern "C" {

int yolopp_abi_version(void) { return YOLOPP_ABI_VERSION; }

const char* yolopp_strerror(int code) {
    switch (code) {
        case YOLOPP_OK: return "ok";
        case YOLOPP_E_INVALID: return "invalid argument or unsupported configuration";
        case YOLOPP_E_WORKSPACE: return "workspace too small";
        case YOLOPP_E_OVERFLOW: return "output capacity overflow";
        case YOLOPP_E_NO_DEVICE: return "no sm_100 CUDA device";
        default: return code >= YOLOPP_E_CUDA ? "CUDA error (code - 1000 = cudaError_t)" : "unknown error";
    }
}

size_t yolopp_workspace_bytes(const yolopp_params* p) {
    Plan plan;
    if (!make_plan(p, &plan)) return 0;
    return plan.total;
}

int yolopp_get_bboxes(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors,
                      const yolopp_outputs* out, void* workspace, size_t workspace_bytes, void* stream) {
    yolopp_plan pl;
    const int rc = prepare(p, level_ptrs, scale_factors, out, workspace, workspace_bytes, 0, &pl);
    if (rc != YOLOPP_OK) return rc;
    return launch(&pl, (cudaStream_t)stream, nullptr, 0);
}

int yolopp_get_bboxes_profiled(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors,
                               const yolopp_outputs* out, void* workspace, size_t workspace_bytes, void* stream,
                               void* const* events, int num_events) {
    if (!events || num_events < YOLOPP_NUM_STAGE_EVENTS) return YOLOPP_E_INVALID;
    yolopp_plan pl;
    const int rc = prepare(p, level_ptrs, scale_factors, out, workspace, workspace_bytes, 0, &pl);
    if (rc != YOLOPP_OK) return rc;
    return launch(&pl, (cudaStream_t)stream, events, num_events);
}

int yolopp_plan_create(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors,
                       const yolopp_outputs* out, void* workspace, size_t workspace_bytes, yolopp_plan** plan) {
    if (!plan) return YOLOPP_E_INVALID;
    *plan = nullptr;
    yolopp_plan* pl = (yolopp_plan*)malloc(sizeof(yolopp_plan));
    if (!pl) return YOLOPP_E_INVALID;
    const int rc = prepare(p, level_ptrs, scale_factors, out, workspace, workspace_bytes, 0, pl);
    if (rc != YOLOPP_OK) {
        free(pl);
        return rc;
    }
    // The launches of one call, captured once into an executable graph: a run is then ONE driver call instead of
    // three launches (+ a memset where no top-k runs). Captured on a private stream in thread-local mode (other
    // threads' streams are not affected); any failure just leaves the direct-launch path in place.
    {
        cudaStream_t cs = nullptr;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess) {
            if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int lrc = launch(pl, cs, nullptr, 0);
                const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
                if (lrc == YOLOPP_OK && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess)
                    pl->exec = exec;
                if (graph) cudaGraphDestroy(graph);
            }
            cudaStreamDestroy(cs);
        }
        (void)cudaGetLastError();  // a failed capture must not poison the caller's error state
    }
    *plan = pl;
    return YOLOPP_OK;
}

int yolopp_plan_run(const yolopp_plan* plan, void* stream) {
    if (!plan) return YOLOPP_E_INVALID;
    if (plan->exec) return cuda_rc(cudaGraphLaunch(plan->exec, (cudaStream_t)stream));
    return launch(plan, (cudaStream_t)stream, nullptr, 0);
}

int yolopp_plan_run_profiled(const yolopp_plan* plan, void* stream, void* const* events, int num_events) {
    if (!plan || !events || num_events < YOLOPP_NUM_STAGE_EVENTS) return YOLOPP_E_INVALID;
    return launch(plan, (cudaStream_t)stream, events, num_events);
}

void yolopp_plan_destroy(yolopp_plan* plan) {
    if (!plan) return;
    if (plan->exec) cudaGraphExecDestroy(plan->exec);
    free(plan);
}

int yolopp_topk_conf(const yolopp_params* p, const float* const* level_ptrs, int32_t* topk_inds, void* workspace,
                     size_t workspace_bytes, void* stream) {
    if (!topk_inds) return YOLOPP_E_INVALID;
    yolopp_plan pl;
    int rc = prepare(p, level_ptrs, nullptr, nullptr, workspace, workspace_bytes, 1, &pl);
    if (rc != YOLOPP_OK) return rc;
    rc = launch(&pl, (cudaStream_t)stream, nullptr, 0);
    if (rc != YOLOPP_OK) return rc;
    const DevParams& d = pl.plan.d;
    return cuda_rc(cudaMemcpyAsync(topk_inds, d.row_anchor, (size_t)d.B * d.R * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                   (cudaStream_t)stream));
}

int yolopp_decode(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors, float* boxes,
                  float* scores, int32_t* topk_inds, void* workspace, size_t workspace_bytes, void* stream_) {
    if (!boxes || !scores) return YOLOPP_E_INVALID;
    cudaStream_t stream = (cudaStream_t)stream_;
    yolopp_plan pl;
    int rc = prepare(p, level_ptrs, scale_factors, nullptr, workspace, workspace_bytes, 2, &pl);
    if (rc != YOLOPP_OK) return rc;
    rc = launch(&pl, stream, nullptr, 0);
    if (rc != YOLOPP_OK) return rc;
    const DevParams& d = pl.plan.d;
    const size_t rows = (size_t)d.B * d.R;
    cudaError_t e = cudaMemcpyAsync(boxes, d.row_box, rows * 16, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(scores, d.mat, rows * d.C * 4, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess && topk_inds)
        e = cudaMemcpyAsync(topk_inds, d.row_anchor, rows * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream);
    return cuda_rc(e);
}

int yolopp_describe(const yolopp_params* p, yolopp_plan_info* info) {
    Plan plan;
    if (!info || !make_plan(p, &plan)) return YOLOPP_E_INVALID;
    memset(info, 0, sizeof(*info));
    const DevParams& d = plan.d;
    info->anchors_per_image = d.N;
    info->rows_per_image = d.R;
    info->num_attrib = d.NA;
    info->tma_tiles = d.tma_tiles;
    info->ldg_blocks = d.ldg_blocks;
    info->dense_tiles = d.dense_tiles;
    info->decode_smem_bytes = (int32_t)plan.dec_smem;
    info->decode_ctas_per_sm = plan.dec_ctas_per_sm;
    info->decode_tile_positions = d.tma_tiles > 0 ? d.tile_t : 0;
    info->workspace_bytes = (int64_t)plan.total;
    for (int l = 0; l < d.L; ++l) {
        int64_t bytes = (int64_t)4 * d.A * d.NA * d.lv[l].HW;
        if (d.lv[l].use_tma) {
            info->tma_level_mask |= 1 << l;
            info->tma_bytes_per_image += bytes;
        } else if (!d.nhwc) {
            info->ldg_bytes_per_image += bytes;
        }
    }
    if (d.nhwc) info->ldg_bytes_per_image = (int64_t)4 * d.NA * d.R;
    int launches = 1;  // nms_image
    if (d.ntopk > 0) ++launches;
    if (d.tma_tiles > 0) ++launches;
    if (d.ldg_blocks > 0) ++launches;
    if (d.dense_tiles > 0) ++launches;
    if (d.nhwc) ++launches;
    info->kernel_launches = launches;
    return YOLOPP_OK;
}

int yolopp_coder_decode(int mode, const float* bboxes, const float* pred, float stride, int64_t n, float* out,
                        void* stream) {
    if (mode != YOLOPP_MODE_CSP && mode != YOLOPP_MODE_V3) return YOLOPP_E_INVALID;
    if (n < 0 || (n > 0 && (!bboxes || !pred || !out))) return YOLOPP_E_INVALID;
    if (n == 0) return YOLOPP_OK;
    if ((((uintptr_t)bboxes | (uintptr_t)pred | (uintptr_t)out) & 15) != 0) return YOLOPP_E_INVALID;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    long long blocks = (n + 255) / 256;
    int cap = di.sms * 8;
    if (blocks > cap) blocks = cap;
    coder_decode_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mode, (const float4*)bboxes, (const float4*)pred,
                                                                      stride, n, (float4*)out);
    return cuda_rc(cudaGetLastError());
}

static int unary(const float* in, float* out, int64_t n, int op, void* stream) {
    if (n < 0 || (n > 0 && (!in || !out))) return YOLOPP_E_INVALID;
    if (n == 0) return YOLOPP_OK;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    long long blocks = (n + 255) / 256;
    int cap = di.sms * 8;
    if (blocks > cap) blocks = cap;
    unary_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n, op);
    return cuda_rc(cudaGetLastError());
}
int yolopp_sigmoid(const float* in, float* out, int64_t n, void* stream) { return unary(in, out, n, 0, stream); }
int yolopp_exp(const float* in, float* out, int64_t n, void* stream) { return unary(in, out, n, 1, stream); }

// grid of a streaming elementwise kernel: enough blocks for every vector, at most MISH_GRID_MULT blocks per SM (grid-
// stride loop beyond that), and a whole number of blocks per SM when the tensor is large
static int stream_grid(long long nvec, int per_block, int sms) {
    long long blocks = (nvec + per_block - 1) / per_block;
    if (blocks < 1) blocks = 1;
#ifndef MISH_GRID_MULT
#define MISH_GRID_MULT 32  // measured on B200 (tools/mish_bench.py): 8 -> 0.92 / 0.94, 32 -> 0.95 / 0.99 of the copy peak (fwd / bwd, f32)
#endif
    const long long cap = (long long)sms * MISH_GRID_MULT;
    if (blocks > cap) blocks = cap;
    else if (blocks > sms) blocks = blocks / sms * sms;
    return (int)blocks;
}

int yolopp_mish_forward(const void* in, void* out, int64_t n, int dtype, void* stream_) {
    if (n < 0 || (n > 0 && (!in || !out))) return YOLOPP_E_INVALID;
    if (dtype != YOLOPP_DTYPE_F32 && dtype != YOLOPP_DTYPE_F16 && dtype != YOLOPP_DTYPE_BF16) return YOLOPP_E_INVALID;
    if (n == 0) return YOLOPP_OK;
    if ((((uintptr_t)in | (uintptr_t)out) & 15) != 0) return YOLOPP_E_INVALID;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int vec = dtype == YOLOPP_DTYPE_F32 ? 4 : 8;
    const int grid = stream_grid(n / vec + 1, MISH_THREADS * MISH_UNROLL, di.sms);
    if (dtype == YOLOPP_DTYPE_F32)
        mish_fwd_kernel<float><<<grid, MISH_THREADS, 0, stream>>>((const float*)in, (float*)out, n);
    else if (dtype == YOLOPP_DTYPE_F16)
        mish_fwd_kernel<__half><<<grid, MISH_THREADS, 0, stream>>>((const __half*)in, (__half*)out, n);
    else
        mish_fwd_kernel<__nv_bfloat16><<<grid, MISH_THREADS, 0, stream>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, n);
    return cuda_rc(cudaGetLastError());
}

int yolopp_mish_backward(const void* grad_out, const void* in, void* grad_in, int64_t n, int dtype, void* stream_) {
    if (n < 0 || (n > 0 && (!grad_out || !in || !grad_in))) return YOLOPP_E_INVALID;
    if (dtype != YOLOPP_DTYPE_F32 && dtype != YOLOPP_DTYPE_F16 && dtype != YOLOPP_DTYPE_BF16) return YOLOPP_E_INVALID;
    if (n == 0) return YOLOPP_OK;
    if ((((uintptr_t)grad_out | (uintptr_t)in | (uintptr_t)grad_in) & 15) != 0) return YOLOPP_E_INVALID;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int vec = dtype == YOLOPP_DTYPE_F32 ? 4 : 8;
    const int grid = stream_grid(n / vec + 1, MISH_THREADS * 2, di.sms);
    if (dtype == YOLOPP_DTYPE_F32)
        mish_bwd_kernel<float><<<grid, MISH_THREADS, 0, stream>>>((const float*)grad_out, (const float*)in, (float*)grad_in, n);
    else if (dtype == YOLOPP_DTYPE_F16)
        mish_bwd_kernel<__half><<<grid, MISH_THREADS, 0, stream>>>((const __half*)grad_out, (const __half*)in, (__half*)grad_in, n);
    else
        mish_bwd_kernel<__nv_bfloat16><<<grid, MISH_THREADS, 0, stream>>>((const __nv_bfloat16*)grad_out, (const __nv_bfloat16*)in,
                                                                         (__nv_bfloat16*)grad_in, n);
    return cuda_rc(cudaGetLastError());
}

int yolopp_synth_level(float* out, int32_t batch, int32_t num_anchors, int32_t num_attrib, int32_t hw,
                       const float* mean3, const float* std3, uint64_t seed, void* stream) {
    if (!out || !mean3 || !std3 || batch < 1 || num_anchors < 1 || num_attrib < 5 || hw < 1) return YOLOPP_E_INVALID;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    Synth3 st;
    for (int i = 0; i < 3; ++i) {
        st.mean[i] = mean3[i];
        st.std[i] = std3[i];
    }
    long long n = (long long)batch * num_anchors * num_attrib * hw;
    long long blocks = (n + 255) / 256;
    int cap = di.sms * 16;
    if (blocks > cap) blocks = cap;
    synth_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(out, n, num_attrib, hw, st, (u64)seed);
    return cuda_rc(cudaGetLastError());
}

#ifdef YPP_PROFILE
int yolopp_prof_read(long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, ypp::g_prof, sizeof(long long) * (size_t)n);
}
int yolopp_prof_cta_read(long long* host) { return (int)cudaMemcpyFromSymbol(host, ypp::g_prof_cta, sizeof(long long) * 1024 * 8); }
int yolopp_ssp_read(long long* host) { return (int)cudaMemcpyFromSymbol(host, ypp::g_ssp, sizeof(long long) * 2 * 64 * 4 * 10); }
int yolopp_ssp2_read(long long* host) { return (int)cudaMemcpyFromSymbol(host, ypp::g_ssp2, sizeof(long long) * 2 * 64 * 4 * 4); }
int yolopp_sub_read(long long* host) { return (int)cudaMemcpyFromSymbol(host, ypp::g_sub, sizeof(long long) * 256 * 32); }
int yolopp_phase_read(long long* host) { return (int)cudaMemcpyFromSymbol(host, ypp::g_phase, sizeof(long long) * 2 * 256 * 16); }
#endif

// ---------------------------------------------------------------------------------------------------
// standalone NMS entry points: the per-image NMS kernel with B = 1
// ---------------------------------------------------------------------------------------------------
static size_t nms_only_smem(int keep_cap, int nlab, int* rowkeys_off) {
    if (keep_cap > NMS_MAX_KEEP) keep_cap = 0;  // kept list in the workspace
    size_t sm = (size_t)NMS_KCAP * 8 * 2 + (size_t)keep_cap * 8 + (size_t)NMS_CH * 24 + (size_t)keep_cap * 28 + (size_t)nlab * 4;
    sm = align_up(sm, 16);
    *rowkeys_off = (int)sm;
    return sm + (size_t)NMS_KCAP * 8;
}

static int launch_nms_only(DevParams& d, int nlab, cudaStream_t stream) {
    int off = 0;
    size_t smem = nms_only_smem(d.keep_cap, nlab, &off);
    d.nms_rowkeys_off = off;
    smem = nms_add_stage(d, smem);
    if (smem > SMEM_LIMIT) return YOLOPP_E_INVALID;
    nms_image_kernel<<<1, NMS_THREADS, smem, stream>>>(d);
    return cuda_rc(cudaGetLastError());
}

static size_t nms_kept_bytes(long long cap) { return cap > NMS_MAX_KEEP ? align_up((size_t)cap * 36, 256) : 0; }

size_t yolopp_nms_workspace_bytes(int64_t n, int32_t num_classes) {
    if (n < 0 || num_classes < 0) return 0;
    // multiclass_nms: score matrix + per-row statistics; batched_nms / nms (num_classes == 0): nothing — plus, in both,
    // the kept list when more than 4096 boxes may be kept (worst case: every candidate)
    const long long cand = num_classes == 0 ? n : n * num_classes;
    const size_t base = num_classes == 0 ? 256 : align_up((size_t)n * num_classes * 4, 256) + align_up((size_t)n * 16, 256) + 256;
    return base + nms_kept_bytes(cand);
}

int yolopp_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n, int32_t num_labels,
                       float iou_thr, float score_threshold, int nms_offset, int split_thr, int class_agnostic,
                       int max_num, float* dets, int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                       void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || n >= (1ll << 31) || !num_keep || (nms_offset != 0 && nms_offset != 1)) return YOLOPP_E_INVALID;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    if (n == 0) return cuda_rc(cudaMemsetAsync(num_keep, 0, 2 * sizeof(int32_t), stream));
    if (!boxes || !scores || !dets || !keep || (((uintptr_t)boxes) & 15)) return YOLOPP_E_INVALID;
    const int nlab = idxs ? num_labels : 1;
    if (nlab < 1 || nlab > YOLOPP_MAX_CLASSES) return YOLOPP_E_INVALID;
    const long long cap = (max_num > 0 && max_num < n) ? max_num : n;
    if (nms_kept_bytes(cap) && (!workspace || workspace_bytes < nms_kept_bytes(cap) || (((uintptr_t)workspace) & 255)))
        return YOLOPP_E_WORKSPACE;  // more than 4096 boxes may be kept: the kept list lives in the workspace
    cudaError_t e0 = cudaMemsetAsync(num_keep, 0, 2 * sizeof(int32_t), stream);
    if (e0 != cudaSuccess) return cuda_rc(e0);
    DevParams d;
    memset(&d, 0, sizeof(d));
    d.nms_kept = nms_kept_bytes(cap) ? (unsigned char*)workspace : nullptr;
    d.nms_kept_stride = (long long)nms_kept_bytes(cap);
    d.B = 1;
    d.R = (int)n;
    d.C = 1;
    d.generic = 1;
    d.num_labels = nlab;
    d.g_scores = scores;
    d.g_labels = (const long long*)idxs;
    d.row_box = (float4*)boxes;
    d.iou_thr = iou_thr;
    d.nms_score_thr = score_threshold;
    d.foff = (float)nms_offset;
    d.split_thr = split_thr;
    d.nms_agnostic = (class_agnostic || !idxs) ? 1 : 0;
    d.m_eff = max_num > 0 ? max_num : -1;
    d.keep_cap = (int)cap;
    d.out_cap = (int)cap;
    d.o_dets = dets;
    d.o_keep = (long long*)keep;
    d.o_count = num_keep;
    d.o_status = num_keep + 1;
    return launch_nms_only(d, nlab, stream);
}

int yolopp_multiclass_nms(const float* multi_bboxes, int boxes_per_class, const float* multi_scores, int64_t n,
                          int32_t num_classes, float score_thr, const float* score_factors, float iou_thr,
                          float nms_score_threshold, int nms_offset, int split_thr, int class_agnostic, int nms_max_num,
                          int max_num, float* dets,
                          int64_t* labels, int64_t* flat_inds, int32_t* num_keep, int32_t* num_candidates,
                          void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || num_classes < 1 || num_classes > YOLOPP_MAX_CLASSES || n * num_classes >= (1ll << 31) || !num_keep ||
        (nms_offset != 0 && nms_offset != 1))
        return YOLOPP_E_INVALID;
    const DeviceInfo di = device_info();
    if (di.rc != YOLOPP_OK) return di.rc;
    if (n == 0) {
        if (num_candidates) cudaMemsetAsync(num_candidates, 0, sizeof(int32_t), stream);
        return cuda_rc(cudaMemsetAsync(num_keep, 0, 2 * sizeof(int32_t), stream));
    }
    if (!multi_bboxes || !multi_scores || !dets || !labels || (((uintptr_t)multi_bboxes) & 15)) return YOLOPP_E_INVALID;
    if (!workspace || workspace_bytes < yolopp_nms_workspace_bytes(n, num_classes) || (((uintptr_t)workspace) & 255))
        return YOLOPP_E_WORKSPACE;
    int m_eff = -1;
    if (max_num > 0) m_eff = max_num;
    if (nms_max_num > 0) m_eff = (m_eff > 0 && m_eff < nms_max_num) ? m_eff : nms_max_num;
    const long long total = n * num_classes;
    const long long cap = (m_eff > 0 && m_eff < total) ? m_eff : total;
    cudaError_t e0 = cudaMemsetAsync(num_keep, 0, 2 * sizeof(int32_t), stream);
    if (e0 != cudaSuccess) return cuda_rc(e0);
    unsigned char* w = (unsigned char*)workspace;
    uint32_t* mat = (uint32_t*)w;
    uint4* row_stat = (uint4*)(w + align_up((size_t)total * 4, 256));
    unsigned char* kept = w + align_up((size_t)total * 4, 256) + align_up((size_t)n * 16, 256) + 256;  // (after the stats)
    multiclass_prep_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(multi_scores, score_factors, (const float4*)multi_bboxes,
                                                                        boxes_per_class ? 1 : 0, (int)n, num_classes,
                                                                        score_thr, mat, row_stat);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_rc(e);
    DevParams d;
    memset(&d, 0, sizeof(d));
    d.nms_kept = nms_kept_bytes(cap) ? kept : nullptr;
    d.nms_kept_stride = (long long)nms_kept_bytes(cap);
    d.B = 1;
    d.R = (int)n;
    d.C = num_classes;
    d.boxes_per_class = boxes_per_class ? 1 : 0;
    d.row_box = (float4*)multi_bboxes;
    d.row_stat = row_stat;
    d.mat = mat;
    d.iou_thr = iou_thr;
    d.nms_score_thr = nms_score_threshold;
    d.foff = (float)nms_offset;
    d.split_thr = split_thr;
    d.nms_agnostic = class_agnostic ? 1 : 0;
    d.m_eff = m_eff;
    d.keep_cap = (int)cap;
    d.out_cap = (int)cap;
    d.o_dets = dets;
    d.o_labels = (long long*)labels;
    d.o_keep = (long long*)flat_inds;
    d.o_count = num_keep;
    d.o_status = num_keep + 1;
    d.o_ncand = num_candidates;
    return launch_nms_only(d, num_classes, stream);
}

}  // extern "C"
