/*
 * oracle.c — CPU restatement of the reference's YOLO post-processing path.   *** TEST INFRASTRUCTURE ***
 *
 * This file is the parity oracle and the CPU baseline ("port"). It is NOT part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (mmdet-yolov4_b200/) never includes, links or calls anything in oracle/.
 *
 * It restates, operation by operation and in the reference's own order (dense sigmoid of every
 * attribute, materialised anchors, stable top-k, (row, class) expansion, nonzero, class-offset boxes,
 * greedy O(n^2) NMS), the following reference code (paths relative to /root/reference):
 *
 *   YOLOCSPHead.get_bboxes / _get_bboxes_single   mmdet/models/dense_heads/yolocsp_head.py:225-382
 *   YOLOV3Head.get_bboxes / _get_bboxes           mmdet/models/dense_heads/yolo_head.py:171-393
 *   YOLOV4BBoxCoder.decode                        mmdet/core/bbox/coder/yolov4_bbox_coder.py:39-67
 *   YOLOBBoxCoder.decode                          mmdet/core/bbox/coder/yolo_bbox_coder.py:60-89
 *   AnchorGenerator.single_level_grid_anchors     mmdet/core/anchor/anchor_generator.py:233-270
 *   multiclass_nms                                mmdet/core/post_processing/bbox_nms.py:7-93
 *   get_k_for_topk                                mmdet/core/export/onnx_helper.py:45-78
 *
 * Third-party arithmetic NOT in the reference tree: mmcv-full, pinned >=1.3.2,<=1.4.0 by
 * mmdet/__init__.py:18-19. batched_nms / nms / NMSop.forward follow upstream mmcv/ops/nms.py and
 * nms_cpu follows mmcv/ops/csrc/pytorch/nms.cpp of that series (published algorithm, restated from
 * memory of the upstream source; the call site is bbox_nms.py:2,84). PARITY UNPINNED at that boundary:
 * the reference holds no test or golden vector for nms/batched_nms/multiclass_nms. What
 * IS pinned: the tests/golden npz fixtures were produced by executing the reference's own head / coder / anchor /
 * multiclass_nms source files from /root/reference (tests/golden/make_golden.py) with this mmcv
 * restatement behind them and torchvision.ops.nms as an independent inner kernel; plus the reference's
 * own known-answer tests tests/test_utils/test_coder.py:8-23 and tests/test_utils/test_anchor.py:148-188.
 *
 * Two deliberate, documented canonicalisations:
 *   1. exp(): torch's CPU exp is a SIMD approximation whose bits depend on the host ISA and on the
 *      position of an element in the vector loop, so no portable bit-exact target exists. The oracle and
 *      the CUDA kernels both use the polynomial below (<=1.02 ulp, monotone on [-30,30]); all other
 *      operations are the reference's fp32 operations, one IEEE rounding each, no FMA contraction.
 *   2. ties: torch.topk / sort(descending) tie order is implementation defined; the oracle uses the
 *      stable order (value desc, index asc) everywhere.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/yolopp.h"

/* ------------------------------------------------------------------------------------------------ */
/* canonical transcendental                                                                           */
/* ------------------------------------------------------------------------------------------------ */
static inline float bits2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* exp(x): round-to-nearest range reduction by the 1.5*2^23 magic constant, Cody-Waite ln2 split,
   degree-5 polynomial (Cephes expf coefficients) evaluated with explicit fmaf, two-step 2^j scaling. */
float oracle_expf(float x) {
    if (!(x == x)) return x;
    float xc = x < -104.0f ? -104.0f : x;
    xc = xc > 89.0f ? 89.0f : xc;
    float t = fmaf(xc, 1.44269502f, 12582912.0f);
    float j = t - 12582912.0f;
    float r = fmaf(j, -0.693145752f, xc);
    r = fmaf(j, -1.42860677e-06f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float e = fmaf(p, r2, r);
    e = e + 1.0f;
    int ji = (int)j;
    int j1 = ji / 2;
    int j2 = ji - j1;
    float s1 = bits2f((uint32_t)(j1 + 127) << 23);
    float s2 = bits2f((uint32_t)(j2 + 127) << 23);
    return (e * s1) * s2;
}

/* torch.sigmoid: 1 / (1 + exp(-x))   (yolocsp_head.py:267, yolo_head.py:267,276-278) */
float oracle_sigmoid(float x) {
    float e = oracle_expf(-x);
    float d = 1.0f + e;
    return 1.0f / d;
}

void oracle_expf_array(const float* in, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_expf(in[i]);
}
void oracle_sigmoid_array(const float* in, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_sigmoid(in[i]);
}

/* ------------------------------------------------------------------------------------------------ */
/* bbox coders                                                                                        */
/* ------------------------------------------------------------------------------------------------ */
/* YOLOV4BBoxCoder.decode (yolov4_bbox_coder.py:52-65) when mode==CSP, YOLOBBoxCoder.decode
   (yolo_bbox_coder.py:74-87) when mode==V3; one fp32 rounding per reference op. */
static inline void decode_one(int mode, const float* a, const float* p, float stride, float* o) {
    float xc = (a[0] + a[2]) * 0.5f;
    float yc = (a[1] + a[3]) * 0.5f;
    float w = a[2] - a[0];
    float h = a[3] - a[1];
    float xcp, ycp, wp, hp;
    if (mode == YOLOPP_MODE_CSP) {
        xcp = p[0] * stride + xc;
        ycp = p[1] * stride + yc;
        wp = p[2] * w;
        hp = p[3] * h;
    } else {
        xcp = (p[0] - 0.5f) * stride + xc;
        ycp = (p[1] - 0.5f) * stride + yc;
        wp = oracle_expf(p[2]) * w;
        hp = oracle_expf(p[3]) * h;
    }
    float hw = wp / 2.0f, hh = hp / 2.0f;
    o[0] = xcp - hw;
    o[1] = ycp - hh;
    o[2] = xcp + hw;
    o[3] = ycp + hh;
}

void oracle_coder_decode(int mode, const float* bboxes, const float* pred, float stride, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) decode_one(mode, bboxes + 4 * i, pred + 4 * i, stride, out + 4 * i);
}

/* AnchorGenerator.single_level_grid_anchors (anchor_generator.py:254-269): shifts are exact ints, the add is fp32.
   out: [H*W*A][4], row (y*W+x)*A+a */
void oracle_grid_anchors(const float* base /*[A][4]*/, int A, int H, int W, int stride_w, int stride_h, float* out) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int a = 0; a < A; ++a) {
                float sx = (float)((int64_t)x * stride_w), sy = (float)((int64_t)y * stride_h);
                float* o = out + 4 * (((int64_t)y * W + x) * A + a);
                o[0] = base[4 * a + 0] + sx;
                o[1] = base[4 * a + 1] + sy;
                o[2] = base[4 * a + 2] + sx;
                o[3] = base[4 * a + 3] + sy;
            }
}

/* ------------------------------------------------------------------------------------------------ */
/* stable descending argsort (canonical tie-break: value desc, index asc)                             */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    float v;
    int64_t i;
} vi_t;

static int cmp_desc(const void* pa, const void* pb) {
    const vi_t* a = (const vi_t*)pa;
    const vi_t* b = (const vi_t*)pb;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    /* NaN never appears in supported inputs; equal (or unordered) values fall through to the index */
    return (a->i > b->i) - (a->i < b->i);
}

static void argsort_desc(const float* v, int64_t n, int64_t* order) {
    vi_t* t = (vi_t*)malloc(sizeof(vi_t) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; ++i) {
        t[i].v = v[i];
        t[i].i = i;
    }
    qsort(t, (size_t)n, sizeof(vi_t), cmp_desc);
    for (int64_t i = 0; i < n; ++i) order[i] = t[i].i;
    free(t);
}

/* ------------------------------------------------------------------------------------------------ */
/* mmcv nms_cpu / nms / batched_nms (third party, restated)                                           */
/* ------------------------------------------------------------------------------------------------ */
/* nms_cpu (mmcv/ops/csrc/pytorch/nms.cpp): returns indices of kept boxes in descending-score order. */
int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, float iou_thr, int offset, int64_t* keep) {
    if (n == 0) return 0;
    float* areas = (float*)malloc(sizeof(float) * (size_t)n);
    int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
    unsigned char* select = (unsigned char*)malloc((size_t)n);
    float foff = (float)offset;
    for (int64_t i = 0; i < n; ++i) {
        const float* b = boxes + 4 * i;
        areas[i] = (b[2] - b[0] + foff) * (b[3] - b[1] + foff);
        select[i] = 1;
    }
    argsort_desc(scores, n, order);
    for (int64_t _i = 0; _i < n; ++_i) {
        if (!select[_i]) continue;
        int64_t i = order[_i];
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float iarea = areas[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            if (!select[_j]) continue;
            int64_t j = order[_j];
            const float* b = boxes + 4 * j;
            float xx1 = ix1 > b[0] ? ix1 : b[0];
            float yy1 = iy1 > b[1] ? iy1 : b[1];
            float xx2 = ix2 < b[2] ? ix2 : b[2];
            float yy2 = iy2 < b[3] ? iy2 : b[3];
            float w = xx2 - xx1 + foff;
            w = w > 0.f ? w : 0.f;
            float h = yy2 - yy1 + foff;
            h = h > 0.f ? h : 0.f;
            float inter = w * h;
            float ovr = inter / (iarea + areas[j] - inter);
            if (ovr > iou_thr) select[_j] = 0;
        }
    }
    int64_t k = 0;
    for (int64_t _i = 0; _i < n; ++_i)
        if (select[_i]) keep[k++] = order[_i];
    free(areas);
    free(order);
    free(select);
    return k;
}

/* mmcv.ops.nms.nms (python wrapper + NMSop.forward): max_num cut; dets = cat(boxes[inds], scores[inds]). */
static int64_t nms_op(const float* boxes, const float* scores, int64_t n, float iou_thr, int offset, int max_num,
                      float score_threshold, int64_t* keep) {
    int64_t k;
    if (score_threshold > 0.f) {
        /* NMSop.forward: valid_mask = scores > score_threshold; nms on the subset; inds = valid_inds[inds] */
        float* vb = (float*)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
        float* vs = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        int64_t* vi = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
        int64_t m = 0;
        for (int64_t i = 0; i < n; ++i)
            if (scores[i] > score_threshold) {
                memcpy(vb + 4 * m, boxes + 4 * i, sizeof(float) * 4);
                vs[m] = scores[i];
                vi[m] = i;
                ++m;
            }
        k = oracle_nms(vb, vs, m, iou_thr, offset, keep);
        if (max_num > 0 && k > max_num) k = max_num;
        for (int64_t q = 0; q < k; ++q) keep[q] = vi[keep[q]];
        free(vb);
        free(vs);
        free(vi);
        return k;
    }
    k = oracle_nms(boxes, scores, n, iou_thr, offset, keep);
    if (max_num > 0 && k > max_num) k = max_num;
    return k;
}

/*
 * mmcv.ops.nms.batched_nms. idxs may be NULL (== class_agnostic). Returns number kept; keep[] holds indices
 * into the inputs in output order; dets[k][5] = (boxes[keep], score).
 */
int64_t oracle_batched_nms2(const float* boxes, const float* scores, const int64_t* idxs, int64_t n, float iou_thr,
                            int offset, int split_thr, int class_agnostic, int max_num, float score_threshold, float* dets,
                            int64_t* keep) {
    if (n == 0) return 0;
    float* bfn = (float*)malloc(sizeof(float) * 4 * (size_t)n);
    if (class_agnostic || idxs == NULL) {
        memcpy(bfn, boxes, sizeof(float) * 4 * (size_t)n);
    } else {
        float m = boxes[0];
        for (int64_t i = 1; i < 4 * n; ++i) m = boxes[i] > m ? boxes[i] : m; /* boxes.max() */
        float mp1 = m + 1.0f;
        for (int64_t i = 0; i < n; ++i) {
            float off = (float)idxs[i] * mp1; /* idxs.to(boxes) * (max_coordinate + 1) */
            for (int k = 0; k < 4; ++k) bfn[4 * i + k] = boxes[4 * i + k] + off;
        }
    }
    int64_t nk = 0;
    if (n < split_thr) {
        nk = nms_op(bfn, scores, n, iou_thr, offset, max_num, score_threshold, keep);
    } else {
        /* per-class loop over torch.unique(idxs) (sorted ascending) */
        unsigned char* total = (unsigned char*)calloc((size_t)n, 1);
        int64_t* uniq = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
        int64_t nu = 0;
        {
            /* unique labels, ascending */
            int64_t* tmp = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
            for (int64_t i = 0; i < n; ++i) tmp[i] = idxs ? idxs[i] : 0;
            /* simple sort */
            for (int64_t gap = n / 2; gap > 0; gap /= 2)
                for (int64_t i = gap; i < n; ++i) {
                    int64_t v = tmp[i], j = i;
                    for (; j >= gap && tmp[j - gap] > v; j -= gap) tmp[j] = tmp[j - gap];
                    tmp[j] = v;
                }
            for (int64_t i = 0; i < n; ++i)
                if (i == 0 || tmp[i] != tmp[i - 1]) uniq[nu++] = tmp[i];
            free(tmp);
        }
        int64_t* mask = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
        float* sb = (float*)malloc(sizeof(float) * 4 * (size_t)n);
        float* ss = (float*)malloc(sizeof(float) * (size_t)n);
        int64_t* sk = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
        for (int64_t u = 0; u < nu; ++u) {
            int64_t m = 0;
            for (int64_t i = 0; i < n; ++i)
                if ((idxs ? idxs[i] : 0) == uniq[u]) mask[m++] = i; /* (idxs == id).nonzero() */
            for (int64_t q = 0; q < m; ++q) {
                memcpy(sb + 4 * q, bfn + 4 * mask[q], sizeof(float) * 4);
                ss[q] = scores[mask[q]];
            }
            /* per-class nms_op: nms_cfg_ no longer holds max_num here (popped before the loop) */
            int64_t k = nms_op(sb, ss, m, iou_thr, offset, -1, score_threshold, sk);
            for (int64_t q = 0; q < k; ++q) total[mask[sk[q]]] = 1;
        }
        /* keep = total_mask.nonzero(); scores[keep].sort(descending=True)  (canonical: stable) */
        int64_t m = 0;
        for (int64_t i = 0; i < n; ++i)
            if (total[i]) {
                mask[m] = i;
                ss[m] = scores[i];
                ++m;
            }
        argsort_desc(ss, m, sk);
        for (int64_t q = 0; q < m; ++q) keep[q] = mask[sk[q]];
        nk = m;
        if (max_num > 0 && nk > max_num) nk = max_num;
        free(total);
        free(uniq);
        free(mask);
        free(sb);
        free(ss);
        free(sk);
    }
    for (int64_t q = 0; q < nk; ++q) {
        memcpy(dets + 5 * q, boxes + 4 * keep[q], sizeof(float) * 4);
        dets[5 * q + 4] = scores[keep[q]];
    }
    free(bfn);
    return nk;
}

int64_t oracle_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n, float iou_thr,
                           int offset, int split_thr, int class_agnostic, int max_num, float* dets, int64_t* keep) {
    return oracle_batched_nms2(boxes, scores, idxs, n, iou_thr, offset, split_thr, class_agnostic, max_num, 0.f, dets, keep);
}

/*
 * multiclass_nms (bbox_nms.py:7-93).
 *   multi_bboxes (n,4) [boxes_per_class==0] or (n, 4*C); multi_scores (n, C+1), last column = background
 *   score_factors (n) or NULL
 *   outputs (capacity n*C): dets[k][5], labels[k], inds[k] = the `keep` the reference returns with
 *   return_inds=True (index into the thresholded candidate list), flat[k] = row*C + class of each detection.
 */
int64_t oracle_multiclass_nms2(const float* multi_bboxes, int per_class_boxes, const float* multi_scores, int64_t n,
                               int C, float score_thr, float iou_thr, int offset, int split_thr, int class_agnostic,
                               int nms_max_num, int max_num, float nms_score_threshold, const float* score_factors,
                               float* dets, int64_t* labels, int64_t* inds, int64_t* flat, int64_t* num_candidates) {
    int64_t total = n * (int64_t)C, m = 0;
    float* bb = (float*)malloc(sizeof(float) * 4 * (size_t)(total > 0 ? total : 1));
    float* sc = (float*)malloc(sizeof(float) * (size_t)(total > 0 ? total : 1));
    int64_t* lb = (int64_t*)malloc(sizeof(int64_t) * (size_t)(total > 0 ? total : 1));
    int64_t* fl = (int64_t*)malloc(sizeof(int64_t) * (size_t)(total > 0 ? total : 1));
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < C; ++c) {
            float s = multi_scores[i * (C + 1) + c];
            if (s > score_thr) { /* valid_mask = scores > score_thr */
                if (score_factors) s = s * score_factors[i];
                const float* b = per_class_boxes ? multi_bboxes + (i * C + c) * 4 : multi_bboxes + i * 4;
                memcpy(bb + 4 * m, b, sizeof(float) * 4);
                sc[m] = s;
                lb[m] = c;
                fl[m] = i * C + c;
                ++m;
            }
        }
    if (num_candidates) *num_candidates = m;
    int64_t nk = 0;
    if (m > 0) {
        int64_t* keep = (int64_t*)malloc(sizeof(int64_t) * (size_t)m);
        float* d = (float*)malloc(sizeof(float) * 5 * (size_t)m);
        nk = oracle_batched_nms2(bb, sc, lb, m, iou_thr, offset, split_thr, class_agnostic, nms_max_num,
                                 nms_score_threshold, d, keep);
        if (max_num > 0 && nk > max_num) nk = max_num;
        for (int64_t q = 0; q < nk; ++q) {
            memcpy(dets + 5 * q, d + 5 * q, sizeof(float) * 5);
            labels[q] = lb[keep[q]];
            if (inds) inds[q] = keep[q];
            if (flat) flat[q] = fl[keep[q]];
        }
        free(keep);
        free(d);
    }
    free(bb);
    free(sc);
    free(lb);
    free(fl);
    return nk;
}

int64_t oracle_multiclass_nms(const float* multi_bboxes, int per_class_boxes, const float* multi_scores, int64_t n,
                              int C, float score_thr, float iou_thr, int offset, int split_thr, int class_agnostic,
                              int nms_max_num, int max_num, const float* score_factors, float* dets,
                              int64_t* labels, int64_t* inds, int64_t* flat, int64_t* num_candidates) {
    return oracle_multiclass_nms2(multi_bboxes, per_class_boxes, multi_scores, n, C, score_thr, iou_thr, offset, split_thr,
                                  class_agnostic, nms_max_num, max_num, 0.f, score_factors, dets, labels, inds, flat,
                                  num_candidates);
}

/* ------------------------------------------------------------------------------------------------ */
/* the whole path                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
static int num_attrib(const yolopp_params* p) { return p->class_agnostic ? 5 : 5 + p->num_classes; }

/* one image; out_* have capacity `cap` rows. Returns number of detections (<= cap) or -1. */
/* intermediate results of one image (SURVEY.md A.3 parity taps); any pointer may be NULL */
typedef struct oracle_taps {
    int32_t* topk_inds; /* [R]     topk_inds of yolocsp_head.py:350-355 / yolo_head.py:281-302 (concatenated anchor index) */
    float* boxes;       /* [R][4]  bbox_pred entering multiclass_nms (after rescale) */
    float* scores;      /* [R][C]  score of candidate (row, class), NaN where valid_mask is false (bbox_nms.py:54-67) */
} oracle_taps;

static int64_t get_bboxes_single(const yolopp_params* p, const float* const* levels, int b, const float* scale,
                                 int64_t cap, float* out_dets, int64_t* out_labels, int32_t* out_anchor,
                                 int32_t* out_row, int32_t* out_ncand, const oracle_taps* taps) {
    const int L = p->num_levels, A = p->num_anchors, NA = num_attrib(p);
    const int C = p->class_agnostic ? 1 : p->num_classes;
    int64_t N = 0;
    for (int l = 0; l < L; ++l) N += (int64_t)p->height[l] * p->width[l] * A;

    float* conf = (float*)malloc(sizeof(float) * (size_t)N);
    float* cls = (float*)malloc(sizeof(float) * (size_t)N * (size_t)C);
    float* box = (float*)malloc(sizeof(float) * 4 * (size_t)N);
    int64_t* src = (int64_t*)malloc(sizeof(int64_t) * (size_t)N); /* row -> concatenated anchor index */
    int64_t R = 0;                                                  /* rows after the (per-level) top-k */

    int64_t lvl_off = 0;
    for (int l = 0; l < L; ++l) {
        const int H = p->height[l], W = p->width[l];
        const int64_t HW = (int64_t)H * W, NL = HW * A;
        const float* base = levels[l] + (int64_t)b * A * NA * HW;
        float* anchors = (float*)malloc(sizeof(float) * 4 * (size_t)NL);
        oracle_grid_anchors(&p->base_anchors[l][0][0], A, H, W, p->stride_w[l], p->stride_h[l], anchors);
        float* lconf = (float*)malloc(sizeof(float) * (size_t)NL);
        float* lcls = (float*)malloc(sizeof(float) * (size_t)NL * (size_t)C);
        float* lbox = (float*)malloc(sizeof(float) * 4 * (size_t)NL);
        const float stride = (float)p->coder_stride[l];
        for (int64_t hw = 0; hw < HW; ++hw)
            for (int a = 0; a < A; ++a) {
                /* permute(0,2,3,1).reshape(B,-1,NA): row = hw*A + a, attr k = channel a*NA + k */
                const int64_t row = hw * A + a;
                const float* t = base + ((int64_t)a * NA) * HW + hw;
                float pr[4];
                if (p->mode == YOLOPP_MODE_CSP) {
                    float s0 = oracle_sigmoid(t[0 * HW]), s1 = oracle_sigmoid(t[1 * HW]);
                    float s2 = oracle_sigmoid(t[2 * HW]), s3 = oracle_sigmoid(t[3 * HW]);
                    pr[0] = s0 * 2.0f - 1.0f; /* yolocsp_head.py:274 */
                    pr[1] = s1 * 2.0f - 1.0f;
                    float w2 = s2 * 2.0f, h2 = s3 * 2.0f; /* :275  (x*2)**2 */
                    pr[2] = w2 * w2;
                    pr[3] = h2 * h2;
                } else {
                    pr[0] = oracle_sigmoid(t[0 * HW]); /* yolo_head.py:267 */
                    pr[1] = oracle_sigmoid(t[1 * HW]);
                    pr[2] = t[2 * HW];
                    pr[3] = t[3 * HW];
                }
                decode_one(p->mode, anchors + 4 * row, pr, stride, lbox + 4 * row);
                lconf[row] = oracle_sigmoid(t[4 * HW]);
                if (!p->class_agnostic)
                    for (int c = 0; c < C; ++c) lcls[row * C + c] = oracle_sigmoid(t[(int64_t)(5 + c) * HW]);
            }
        if (p->mode == YOLOPP_MODE_V3 && p->nms_pre > 0 && p->nms_pre < NL) {
            /* per-level top-k (yolo_head.py:281-302), rows re-ordered to top-k order */
            int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)NL);
            argsort_desc(lconf, NL, order);
            for (int64_t q = 0; q < p->nms_pre; ++q) {
                int64_t r = order[q];
                conf[R] = lconf[r];
                memcpy(box + 4 * R, lbox + 4 * r, sizeof(float) * 4);
                if (!p->class_agnostic) memcpy(cls + R * C, lcls + r * C, sizeof(float) * (size_t)C);
                src[R] = lvl_off + r;
                ++R;
            }
            free(order);
        } else {
            for (int64_t r = 0; r < NL; ++r) {
                conf[R] = lconf[r];
                memcpy(box + 4 * R, lbox + 4 * r, sizeof(float) * 4);
                if (!p->class_agnostic) memcpy(cls + R * C, lcls + r * C, sizeof(float) * (size_t)C);
                src[R] = lvl_off + r;
                ++R;
            }
        }
        lvl_off += NL;
        free(anchors);
        free(lconf);
        free(lcls);
        free(lbox);
    }

    if (p->mode == YOLOPP_MODE_CSP && p->nms_pre > 0 && p->nms_pre < R) {
        /* conf_pred.topk(nms_pre) over all levels (yolocsp_head.py:350-355) */
        int64_t* order = (int64_t*)malloc(sizeof(int64_t) * (size_t)R);
        argsort_desc(conf, R, order);
        int64_t K = p->nms_pre;
        float* c2 = (float*)malloc(sizeof(float) * (size_t)K);
        float* b2 = (float*)malloc(sizeof(float) * 4 * (size_t)K);
        float* s2 = (float*)malloc(sizeof(float) * (size_t)K * (size_t)C);
        int64_t* r2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)K);
        for (int64_t q = 0; q < K; ++q) {
            int64_t r = order[q];
            c2[q] = conf[r];
            memcpy(b2 + 4 * q, box + 4 * r, sizeof(float) * 4);
            if (!p->class_agnostic) memcpy(s2 + q * C, cls + r * C, sizeof(float) * (size_t)C);
            r2[q] = src[r];
        }
        memcpy(conf, c2, sizeof(float) * (size_t)K);
        memcpy(box, b2, sizeof(float) * 4 * (size_t)K);
        if (!p->class_agnostic) memcpy(cls, s2, sizeof(float) * (size_t)K * (size_t)C);
        memcpy(src, r2, sizeof(int64_t) * (size_t)K);
        R = K;
        free(order);
        free(c2);
        free(b2);
        free(s2);
        free(r2);
    }

    if (p->rescale && scale) /* bbox_pred /= bbox_pred.new_tensor(scale_factor) */
        for (int64_t r = 0; r < R; ++r)
            for (int k = 0; k < 4; ++k) box[4 * r + k] = box[4 * r + k] / scale[k];

    if (taps) {
        /* rows are indexed in top-k rank order BEFORE the V3 conf_thr row filter (like the `rows` tap) */
        for (int64_t r = 0; r < R; ++r) {
            if (taps->topk_inds) taps->topk_inds[r] = (int32_t)src[r];
            if (taps->boxes) memcpy(taps->boxes + 4 * r, box + 4 * r, sizeof(float) * 4);
            if (taps->scores) {
                const int dropped = p->mode == YOLOPP_MODE_V3 && p->conf_thr > 0.f && !(conf[r] >= p->conf_thr);
                for (int c = 0; c < C; ++c) {
                    float sc;
                    int valid;
                    if (p->mode == YOLOPP_MODE_CSP) {
                        sc = p->class_agnostic ? conf[r] : cls[r * C + c] * conf[r]; /* yolocsp_head.py:358 / :360 */
                        valid = sc > p->score_thr;                                    /* bbox_nms.py:54 */
                    } else {
                        valid = cls[r * C + c] > p->score_thr; /* bbox_nms.py:54 on the class score ... */
                        sc = cls[r * C + c] * conf[r];         /* ... then * score_factors (:57-62) */
                    }
                    taps->scores[r * C + c] = (!dropped && valid) ? sc : NAN;
                }
            }
        }
    }

    /* scores (R, C+1) with zero background column */
    float* ms = (float*)malloc(sizeof(float) * (size_t)(R > 0 ? R : 1) * (size_t)(C + 1));
    const float* factors = NULL;
    int64_t R2 = R;
    /* parity tap: row index in top-k rank order BEFORE the V3 conf_thr filter */
    int64_t* rowid = (int64_t*)malloc(sizeof(int64_t) * (size_t)(R > 0 ? R : 1));
    for (int64_t r = 0; r < R; ++r) rowid[r] = r;
    if (p->mode == YOLOPP_MODE_CSP) {
        for (int64_t r = 0; r < R; ++r) {
            for (int c = 0; c < C; ++c)
                ms[r * (C + 1) + c] = p->class_agnostic ? conf[r] : cls[r * C + c] * conf[r]; /* :358 / :360 */
            ms[r * (C + 1) + C] = 0.f;
        }
    } else {
        /* conf_thr row filter, order preserved (yolo_head.py:365-376) */
        if (p->conf_thr > 0.f) {
            int64_t w = 0;
            for (int64_t r = 0; r < R; ++r)
                if (conf[r] >= p->conf_thr) {
                    conf[w] = conf[r];
                    memmove(box + 4 * w, box + 4 * r, sizeof(float) * 4);
                    memmove(cls + w * C, cls + r * C, sizeof(float) * (size_t)C);
                    src[w] = src[r];
                    rowid[w] = rowid[r];
                    ++w;
                }
            R2 = w;
        }
        for (int64_t r = 0; r < R2; ++r) {
            for (int c = 0; c < C; ++c) ms[r * (C + 1) + c] = cls[r * C + c];
            ms[r * (C + 1) + C] = 0.f;
        }
        factors = conf; /* score_factors=mlvl_conf_scores (yolo_head.py:378-384) */
    }

    int64_t total = R2 * (int64_t)C;
    float* d = (float*)malloc(sizeof(float) * 5 * (size_t)(total > 0 ? total : 1));
    int64_t* lab = (int64_t*)malloc(sizeof(int64_t) * (size_t)(total > 0 ? total : 1));
    int64_t* flat = (int64_t*)malloc(sizeof(int64_t) * (size_t)(total > 0 ? total : 1));
    int64_t ncand = 0;
    int64_t nk = oracle_multiclass_nms2(box, 0, ms, R2, C, p->score_thr, p->iou_thr, p->nms_offset, p->split_thr,
                                        p->nms_class_agnostic, p->nms_max_num, p->max_per_img, p->nms_score_thr, factors, d,
                                        lab, NULL, flat, &ncand);
    if (nk > cap) nk = -1;
    for (int64_t q = 0; q < nk; ++q) {
        memcpy(out_dets + 5 * q, d + 5 * q, sizeof(float) * 5);
        out_labels[q] = lab[q];
        if (out_row) out_row[q] = (int32_t)rowid[flat[q] / C];
        if (out_anchor) out_anchor[q] = (int32_t)src[flat[q] / C];
    }
    if (out_ncand) *out_ncand = (int32_t)ncand;
    free(conf);
    free(cls);
    free(box);
    free(src);
    free(ms);
    free(rowid);
    free(d);
    free(lab);
    free(flat);
    return nk;
}

/*
 * Host mirror of yolopp_get_bboxes: all pointers are HOST pointers. out arrays are [B][cap] like
 * yolopp_outputs. num_threads <= 0: all OpenMP threads. Returns 0, or YOLOPP_E_INVALID when cap is too small.
 */
int oracle_get_bboxes(const yolopp_params* p, const float* const* levels, const float* scale_factors, int64_t cap,
                      float* dets, int64_t* labels, int32_t* anchors, int32_t* rows, int32_t* count,
                      int32_t* num_candidates, int num_threads) {
    int rc = 0;
#ifdef _OPENMP
    int nt = num_threads > 0 ? num_threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
#else
    (void)num_threads;
#endif
    for (int b = 0; b < p->batch; ++b) {
        int64_t nk = get_bboxes_single(p, levels, b, scale_factors ? scale_factors + 4 * b : NULL, cap,
                                       dets + (int64_t)b * cap * 5, labels + (int64_t)b * cap,
                                       anchors ? anchors + (int64_t)b * cap : NULL, rows ? rows + (int64_t)b * cap : NULL,
                                       num_candidates ? num_candidates + b : NULL, NULL);
        if (nk < 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            rc = YOLOPP_E_INVALID;
            nk = 0;
        }
        count[b] = (int32_t)nk;
    }
    return rc;
}

/* Host mirror of yolopp_decode / yolopp_topk_conf: the taps of every image. rows = rows entering multiclass_nms
 * per image (sum over segments of min(nms_pre, N_segment)); arrays are [B][rows]... like the device entries. */
int oracle_get_taps(const yolopp_params* p, const float* const* levels, const float* scale_factors, int64_t rows,
                    int32_t* topk_inds, float* boxes, float* scores) {
    const int C = p->class_agnostic ? 1 : p->num_classes;
    const int64_t cap = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < p->batch; ++b) {
        oracle_taps t;
        t.topk_inds = topk_inds ? topk_inds + (int64_t)b * rows : NULL;
        t.boxes = boxes ? boxes + (int64_t)b * rows * 4 : NULL;
        t.scores = scores ? scores + (int64_t)b * rows * C : NULL;
        float d[5];
        int64_t lab[1];
        yolopp_params q = *p;
        q.max_per_img = 1; /* the detections themselves are not wanted here */
        q.nms_max_num = -1;
        get_bboxes_single(&q, levels, b, scale_factors ? scale_factors + 4 * b : NULL, cap, d, lab, NULL, NULL, NULL, &t);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* Mish (mmdet/ops/mish_cuda/src/mish.h:17-29) as mish_cpu.cc:6-29 instantiates it for float tensors. NOTE the
 * types: the header calls exp / log1p / tanh UNQUALIFIED, which in the host build resolve to the C library's
 * double functions, so every call computes in double and only assignments to `scalar_t` round to float. Verified
 * bit for bit against the reference header compiled here (oracle/_ref/libmish_ref.so, tests/test_cpu_oracle.py). */
/* ------------------------------------------------------------------------------------------------ */
float oracle_mish_fwd(float inp) {
    const double sp = inp < 20.0f ? log1p(exp((double)inp)) : (double)inp;
    return (float)((double)inp * tanh(sp));
}
float oracle_mish_bwd(float grad_out, float inp) {
    const float sp = (float)(inp < 20.0f ? log1p(exp((double)inp)) : (double)inp);
    const float grad_sp = (float)(1 - exp((double)-sp));
    const float tsp = (float)tanh((double)sp);
    const float grad_tsp = (1 - tsp * tsp) * grad_sp;
    const float grad = inp * grad_tsp + tsp;
    return grad_out * grad;
}
void oracle_mish_fwd_array(const float* in, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = oracle_mish_fwd(in[i]);
}
void oracle_mish_bwd_array(const float* grad_out, const float* in, float* grad_in, int64_t n) {
    for (int64_t i = 0; i < n; ++i) grad_in[i] = oracle_mish_bwd(grad_out[i], in[i]);
}

/* ------------------------------------------------------------------------------------------------ */
/* synthetic head tensors (same bits as yolopp_synth_level and oracle/synth.py)                        */
/* ------------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}

void oracle_synth_level(float* out, int32_t batch, int32_t num_anchors, int32_t na, int32_t hw, const float* mean,
                        const float* std, uint64_t seed) {
    int64_t n = (int64_t)batch * num_anchors * na * hw;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t x = splitmix64(seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL);
        int32_t s = (int32_t)(x & 0xFFFF) + (int32_t)((x >> 16) & 0xFFFF) + (int32_t)((x >> 32) & 0xFFFF) +
                    (int32_t)(x >> 48);
        float z = (float)(s - 131070) * 2.64290273e-05f; /* 1/37837.23: unit variance Irwin-Hall(4) */
        int k = (int)((i / hw) % na);
        float v = std[k] * z;
        out[i] = mean[k] + v;
    }
}
