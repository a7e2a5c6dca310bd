/*
 * yolopp.h — C ABI of the B200-native YOLO detection post-processing path.
 *
 * Drop-in boundary for ONE hot path of zhanggefan/mmdet-yolov4 (an MMDetection 2.12 fork):
 *
 *   YOLOCSPHead.get_bboxes / _get_bboxes_single   mmdet/models/dense_heads/yolocsp_head.py:225-382
 *   YOLOV3Head.get_bboxes / _get_bboxes           mmdet/models/dense_heads/yolo_head.py:171-393
 *   YOLOV4BBoxCoder.decode                        mmdet/core/bbox/coder/yolov4_bbox_coder.py:39-67
 *   YOLOBBoxCoder.decode                          mmdet/core/bbox/coder/yolo_bbox_coder.py:60-89
 *   AnchorGenerator.grid_anchors (YOLO)           mmdet/core/anchor/anchor_generator.py:207-270,639-665
 *   multiclass_nms                                mmdet/core/post_processing/bbox_nms.py:7-93
 *   mmcv.ops.nms.batched_nms / nms / nms_cpu      third party, mmcv-full 1.3.2..1.4.0 (call site bbox_nms.py:2,84)
 * and, around the path:
 *   bbox2result (per-class split of the detections)  mmdet/core/bbox/transforms.py:99-116
 *   Mish activation forward / backward               mmdet/ops/mish_cuda/src/mish.h:17-29, kernel/mish_cuda.cu:26-71
 *   channels-last (NHWC) head outputs                mmdet/models/dense_heads/yolocsp_head.py:216-222 (convs_pred output)
 *
 * Conventions
 *   - plain C, no torch types; every pointer marked "device" is a CUDA device pointer on the
 *     current device, every pointer marked "host" is ordinary host memory.
 *   - the library never allocates, frees or retains device memory: the caller owns inputs, outputs
 *     and the workspace (size from yolopp_workspace_bytes).
 *   - every call is asynchronous on the cudaStream_t it is given (passed as void*), performs no host
 *     synchronisation, and is re-entrant (no global state).
 *   - return value: 0 = ok, YOLOPP_E_* otherwise. CUDA launch errors are returned as
 *     YOLOPP_E_CUDA + cudaError_t. Data-dependent failures (candidate-buffer overflow) cannot be
 *     known without a sync; they are written to the `status` word of the output block, which the
 *     caller reads together with the counts.
 *   - arithmetic is IEEE fp32 with one rounding per reference operation and NO fused multiply-add
 *     contraction; exp() is the canonical polynomial defined in DESIGN.md §"Canonical arithmetic"
 *     (the same on CPU and GPU, <= 1.02 ulp), so results are bit-reproducible.
 *   - ties are broken canonically (score desc, then flat candidate index asc) — DESIGN.md §"Ties".
 */
#ifndef YOLOPP_H_
#define YOLOPP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YOLOPP_ABI_VERSION 3

#define YOLOPP_MAX_LEVELS 8
#define YOLOPP_MAX_ANCHORS 8      /* base anchors per level (mmdet asserts the same count on every level) */
#define YOLOPP_MAX_CLASSES 4096   /* 12-bit class field in the sort key */
#define YOLOPP_MAX_ROWS (1 << 20) /* 20-bit row field in the sort key: rows entering multiclass_nms per image */

/* status / error codes */
#define YOLOPP_OK 0
#define YOLOPP_E_INVALID 1        /* bad argument / unsupported configuration */
#define YOLOPP_E_WORKSPACE 2      /* workspace too small */
#define YOLOPP_E_OVERFLOW 3       /* (status word) a candidate bin overflowed its capacity */
#define YOLOPP_E_NO_DEVICE 4      /* no CUDA device / not an sm_100 device */
#define YOLOPP_E_CUDA 1000        /* + cudaError_t */

/* decode conventions */
#define YOLOPP_MODE_CSP 0 /* YOLOCSPHead + YOLOV4BBoxCoder: sigmoid on every attr, xy=(2s-1)*stride+c, wh=(2s)^2*a,
                             ONE objectness top-k per image over all levels, score = cls*conf, then score > thr */
#define YOLOPP_MODE_V3 1  /* YOLOV3Head + YOLOBBoxCoder: xy=(s-0.5)*stride+c, wh=exp(t)*a, objectness top-k PER LEVEL,
                             conf >= conf_thr row filter, cls > thr tested BEFORE score = cls*conf */

/*
 * Everything the reference reads from the head instance, the test_cfg and img_metas.
 *   head instance : yolocsp_head.py:112-114,151,162,170-178 ; yolo_head.py:52-59
 *   test_cfg      : yolocsp_head.py:345-348,374-376 ; yolo_head.py:281,365,378-384
 */
typedef struct yolopp_params {
    int32_t abi_version;   /* = YOLOPP_ABI_VERSION */
    int32_t mode;          /* YOLOPP_MODE_* */
    int32_t batch;         /* B: images in this call */
    int32_t num_levels;    /* L, in HEAD ORDER (CSP: strides 8,16,32; V3: 32,16,8) */
    int32_t num_anchors;   /* A base anchors per level */
    int32_t num_classes;   /* C (ignored when class_agnostic: one class whose score is the objectness) */
    int32_t class_agnostic;/* YOLOCSPHead(class_agnostic=True): 5 attrs/anchor (yolocsp_head.py:175-178,360) */
    int32_t height[YOLOPP_MAX_LEVELS];   /* feature-map H per level */
    int32_t width[YOLOPP_MAX_LEVELS];    /* feature-map W per level */
    int32_t stride_w[YOLOPP_MAX_LEVELS]; /* anchor-generator stride (x) — anchor_generator.py:256 */
    int32_t stride_h[YOLOPP_MAX_LEVELS]; /* anchor-generator stride (y) — anchor_generator.py:257 */
    int32_t coder_stride[YOLOPP_MAX_LEVELS]; /* featmap_strides[l], the `stride` passed to bbox_coder.decode */
    /* fp32 base anchors exactly as YOLOAnchorGenerator.gen_single_level_base_anchors builds them
       (x1,y1,x2,y2 computed in double, rounded once to fp32 — anchor_generator.py:655-662). */
    float base_anchors[YOLOPP_MAX_LEVELS][YOLOPP_MAX_ANCHORS][4];
    /* test_cfg */
    int32_t nms_pre;       /* <=0: no top-k */
    float score_thr;       /* candidates need score > score_thr (strict, fp32) — bbox_nms.py:54 */
    float conf_thr;        /* V3 only: rows need conf >= conf_thr when conf_thr > 0 — yolo_head.py:365-376 */
    float iou_thr;         /* nms_cfg['iou_threshold']: suppress when iou > iou_thr (strict) */
    int32_t nms_offset;    /* nms_cfg.get('offset', 0): 0 or 1 (mmcv nms) */
    int32_t split_thr;     /* nms_cfg.get('split_thr', 10000) (mmcv batched_nms) */
    int32_t nms_class_agnostic; /* nms_cfg.get('class_agnostic', False): no class offsets, one NMS problem */
    int32_t nms_max_num;   /* nms_cfg.get('max_num', -1) */
    int32_t max_per_img;   /* cfg.max_per_img (<=0: keep all) */
    int32_t rescale;       /* divide boxes by scale_factor[b][0..3] before NMS */
    /* capacity knobs (0 = worst case) */
    int32_t out_capacity;  /* rows of the per-image output block; 0 -> max_per_img (must be > 0 then). With max_per_img <= 0
                              ("keep all") a status of YOLOPP_E_OVERFLOW reports images with more survivors than this */
    /* scheduling hint (no effect on results): how many batches the caller keeps in flight on different streams.
       <= 1: the batch runs alone -> lowest latency schedule; > 1: schedule that lets neighbouring batches overlap */
    int32_t batches_in_flight;
    /* memory layout of the level tensors: YOLOPP_LAYOUT_NCHW = contiguous (B, A*(5+C), H, W) as the reference's
       convs_pred return them (yolocsp_head.py:216-222); YOLOPP_LAYOUT_NHWC = the channels-last memory of the same
       logical tensor, (B, H, W, A*(5+C)) — what a cuDNN NHWC convolution writes. With NHWC the 5+C logits of an
       anchor are contiguous and only the admitted anchors are read (no pass over the whole tensor). */
    int32_t layout;
    /* nms_cfg.get('score_threshold', 0) (mmcv NMSop.forward): when > 0 only candidates with score > threshold take
       part in the greedy pass (the regime decision and boxes.max() still see every candidate) */
    float nms_score_thr;
    int32_t reserved[4];
} yolopp_params;

#define YOLOPP_LAYOUT_NCHW 0
#define YOLOPP_LAYOUT_NHWC 1

/* Per-detection outputs. All arrays are device memory, [batch][out_capacity] row-major, caller-owned.
   Rows >= count[b] are left untouched. */
typedef struct yolopp_outputs {
    float* dets;        /* [B][cap][5]  x1,y1,x2,y2,score  (bbox_nms.py:86-93) */
    int64_t* labels;    /* [B][cap]     class id (int64 like the reference) */
    int32_t* anchors;   /* [B][cap]     parity tap: concatenated anchor index (level-major, (y*W+x)*A+a); may be NULL */
    int32_t* rows;      /* [B][cap]     parity tap: row index in top-k rank order (before the V3 conf_thr row filter); may be NULL */
    int32_t* count;     /* [B]          number of detections */
    int32_t* num_candidates; /* [B]     candidates that entered batched_nms (bbox_nms.py:66); may be NULL */
    int32_t* status;    /* [1]          0 or YOLOPP_E_OVERFLOW, data dependent; written every call */
    /* bbox2result on the device (mmdet/core/bbox/transforms.py:99-116): the same detections grouped by label, each
       group in score order (== bboxes[labels == c]); group c = rows cls_offsets[b][c] .. cls_offsets[b][c+1]-1.
       Both NULL or both set. */
    float* cls_dets;       /* [B][cap][5] */
    int32_t* cls_offsets;  /* [B][C+1]   (C = 1 when class agnostic) */
} yolopp_outputs;

/* ABI version of the loaded library. */
int yolopp_abi_version(void);

/* Human-readable name of a return code (static storage). */
const char* yolopp_strerror(int code);

/* Validates `p` and returns the workspace size in bytes needed by yolopp_get_bboxes (0 on invalid params). */
size_t yolopp_workspace_bytes(const yolopp_params* p);

/*
 * The whole path (replaces YOLOCSPHead.get_bboxes yolocsp_head.py:225-310 and
 * YOLOV3Head.get_bboxes yolo_head.py:171-207, with_nms=True).
 *   level_ptrs   host array[L] of DEVICE pointers; level l is the raw head output (B, A*(5+C), H_l, W_l),
 *                contiguous NCHW fp32, channel = a*(5+C)+k   (yolocsp_head.py:264-265)
 *   scale_factors DEVICE pointer [B][4] fp32 (w,h,w,h) or NULL when !rescale (yolocsp_head.py:365-366)
 *   workspace    DEVICE pointer, >= yolopp_workspace_bytes(p) bytes, 256-byte aligned
 */
int yolopp_get_bboxes(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors,
                      const yolopp_outputs* out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Same as yolopp_get_bboxes, and records `events[i]` (cudaEvent_t created by the caller with timing enabled)
 * on `stream` at the stage boundaries, so a benchmark can time each kernel on the launching stream without a
 * profiler:  0 select | 1 decode (TMA levels) | 2 decode (other levels) | 3 per-image NMS + output | 4 end.
 * num_events must be >= YOLOPP_NUM_STAGE_EVENTS.
 */
#define YOLOPP_NUM_STAGE_EVENTS 5
int yolopp_get_bboxes_profiled(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors,
                               const yolopp_outputs* out, void* workspace, size_t workspace_bytes, void* stream,
                               void* const* events, int num_events);

/*
 * Plan handle for serving loops: everything yolopp_get_bboxes derives per call (validation, workspace layout, TMA
 * tensor maps, grid sizes) is computed once for a fixed set of buffers; yolopp_plan_run is then three kernel
 * launches — captured once into an executable CUDA graph, so a run is a single cudaGraphLaunch — and nothing else on
 * the host. The handle is host memory owned by the caller (create / destroy) plus that graph object; it holds no
 * device buffers and no global state; the buffers it was created for must stay alive while it is used.
 * One plan may be run on any stream, but not concurrently with itself (it owns its workspace).
 */
typedef struct yolopp_plan yolopp_plan;
int yolopp_plan_create(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors,
                       const yolopp_outputs* out, void* workspace, size_t workspace_bytes, yolopp_plan** plan);
int yolopp_plan_run(const yolopp_plan* plan, void* stream);
/* as yolopp_get_bboxes_profiled */
int yolopp_plan_run_profiled(const yolopp_plan* plan, void* stream, void* const* events, int num_events);
void yolopp_plan_destroy(yolopp_plan* plan);

/*
 * Stage entries = the parity taps of SURVEY.md A.3. Same params / level tensors / workspace as yolopp_get_bboxes.
 *
 * yolopp_topk_conf: the objectness top-k alone (yolocsp_head.py:348-355: `_, topk_inds = conf_pred.topk(nms_pre)`;
 *   yolo_head.py:281-302 per level). topk_inds DEVICE [B][R] int32, R = yolopp_plan_info.rows_per_image: row r of
 *   image b = concatenated anchor index (level-major, (y*W+x)*A+a) of the r-th row that enters multiclass_nms, in
 *   canonical (conf desc, anchor asc) order per top-k segment; segments without a top-k list their anchors in
 *   order.
 *
 * yolopp_decode: top-k + fused decode, i.e. the tensors that enter multiclass_nms / batched_nms (bbox_nms.py:34-67):
 *   boxes  DEVICE [B][R][4]  decoded (and rescaled) box of every row
 *   scores DEVICE [B][R][C]  score of candidate (row, class) — cls*conf (CSP) / cls*score_factor (V3) — as fp32, and
 *          the bit pattern 0xFFFFFFFF (a NaN) where the pair is NOT a candidate (failed score_thr / conf_thr); the
 *          reference's candidate list `inds = valid_mask.nonzero()` is the row-major order of the non-NaN entries.
 *   topk_inds as above (may be NULL).
 */
int yolopp_topk_conf(const yolopp_params* p, const float* const* level_ptrs, int32_t* topk_inds, void* workspace,
                     size_t workspace_bytes, void* stream);
int yolopp_decode(const yolopp_params* p, const float* const* level_ptrs, const float* scale_factors, float* boxes,
                  float* scores, int32_t* topk_inds, void* workspace, size_t workspace_bytes, void* stream);

/* How a configuration is executed (for DESIGN.md / bench.py roofline arithmetic). */
typedef struct yolopp_plan_info {
    int32_t anchors_per_image;   /* N */
    int32_t rows_per_image;      /* R: rows entering multiclass_nms */
    int32_t num_attrib;          /* 5 + C (5 when class agnostic) */
    int32_t tma_level_mask;      /* bit l set: level l is streamed by the TMA decode kernel (plain or quad-row tiles) */
    int32_t tma_tiles;           /* tiles of the TMA decode kernel */
    int32_t ldg_blocks;          /* blocks of the generic decode kernel */
    int32_t decode_smem_bytes;   /* dynamic shared memory of the TMA decode kernel */
    int32_t decode_ctas_per_sm;
    int32_t kernel_launches;     /* kernels launched per yolopp_get_bboxes call */
    int32_t dense_tiles;         /* 32-position tiles of the dense-admission decode kernel */
    int64_t tma_bytes_per_image; /* algorithmic bytes read per image by the persistent decode kernel (TMA tiles +
                                    gather tiles of unaligned levels): 4*A*(5+C)*sum(HW) over its levels */
    int64_t ldg_bytes_per_image; /* ... by the other decode kernels (dense admission, > 256 attributes; NHWC: the
                                    admitted rows, 4*(5+C)*R) */
    int64_t workspace_bytes;
    int32_t decode_tile_positions; /* positions per tile of the TMA decode kernel (64 or 32; 0: kernel not launched) */
    int32_t reserved_;
} yolopp_plan_info;
int yolopp_describe(const yolopp_params* p, yolopp_plan_info* info);


/*
 * bbox_coder.decode as a standalone elementwise op (YOLOV4BBoxCoder.decode yolov4_bbox_coder.py:39-67 when
 * mode == YOLOPP_MODE_CSP, YOLOBBoxCoder.decode yolo_bbox_coder.py:60-89 when mode == YOLOPP_MODE_V3).
 * bboxes, pred, out: DEVICE [n][4] fp32. `stride` is the scalar stride.
 */
int yolopp_coder_decode(int mode, const float* bboxes, const float* pred, float stride, int64_t n, float* out,
                        void* stream);

/*
 * mmcv.ops.nms.batched_nms on device — also backs mmcv.ops.nms.nms (idxs == NULL, split_thr = INT_MAX).
 * Replaces the third-party call at mmdet/core/post_processing/bbox_nms.py:84 and the direct callers
 * (rpn_head.py:247, cascade_rpn_head.py:670, ...). Same semantics as the in-path NMS: class-offset boxes
 * boxes + idx*(max+1) unless class_agnostic, ONE greedy pass when n < split_thr else classes independent,
 * result in (score desc, index asc) order, first max_num kept. score_threshold > 0: mmcv's NMSop prefilter
 * (only boxes with score > score_threshold enter the greedy pass).
 *   boxes  DEVICE [n][4] fp32 (16-byte aligned), scores DEVICE [n], idxs DEVICE [n] int64 in [0, num_labels) or NULL
 *   dets   DEVICE [cap][5], keep DEVICE [cap] int64 (indices into the inputs), cap = max_num if 0 < max_num < n else n
 *   num_keep DEVICE int32[2]: [0] = number kept, [1] = status (0)
 *   workspace: only needed when cap > 4096 (the kept list then lives there instead of shared memory):
 *          >= yolopp_nms_workspace_bytes(n, 0) bytes, 256-byte aligned; may be NULL otherwise
 */
int yolopp_batched_nms(const float* boxes, const float* scores, const int64_t* idxs, int64_t n, int32_t num_labels,
                       float iou_thr, float score_threshold, int nms_offset, int split_thr, int class_agnostic,
                       int max_num, float* dets, int64_t* keep, int32_t* num_keep, void* workspace, size_t workspace_bytes,
                       void* stream);

/*
 * multiclass_nms (mmdet/core/post_processing/bbox_nms.py:7-93) on device, for the other heads that share the
 * back-end: (row, class) expansion, scores > score_thr, optional score_factors AFTER the threshold, batched_nms,
 * first max_num.
 *   multi_bboxes DEVICE [n][4] (boxes_per_class = 0) or [n][C][4] (boxes_per_class = 1), 16-byte aligned
 *   multi_scores DEVICE [n][C+1] (last column = background, ignored); score_factors DEVICE [n] or NULL
 *   dets DEVICE [cap][5], labels DEVICE [cap] int64, flat_inds DEVICE [cap] int64 (row*C + class; may be NULL),
 *   num_keep DEVICE int32[2] ([0] count, [1] 0 or YOLOPP_E_OVERFLOW), num_candidates DEVICE [1] (may be NULL);
 *   cap = effective max_num when 0 < it < n*C, else n*C
 *   workspace >= yolopp_nms_workspace_bytes(n, C), 256-byte aligned
 */
size_t yolopp_nms_workspace_bytes(int64_t n, int32_t num_classes);
int yolopp_multiclass_nms(const float* multi_bboxes, int boxes_per_class, const float* multi_scores, int64_t n,
                          int32_t num_classes, float score_thr, const float* score_factors, float iou_thr,
                          float nms_score_threshold, int nms_offset, int split_thr, int class_agnostic, int nms_max_num,
                          int max_num, float* dets,
                          int64_t* labels, int64_t* flat_inds, int32_t* num_keep, int32_t* num_candidates,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * Mish activation (mmdet/ops/mish_cuda): y = x * tanh(softplus(x)), softplus threshold 20 (mish.h:15-19);
 * backward dx = dy * (x * (1 - tanh(sp)^2) * (1 - exp(-sp)) + tanh(sp)) (mish.h:22-29). Half / bfloat16 compute in
 * fp32 (mish.h:33-50). Elementwise over n contiguous elements, 16-byte aligned pointers, on `stream` (the
 * reference launches on the default stream: mish_cuda.cu:53,70).
 */
#define YOLOPP_DTYPE_F32 0
#define YOLOPP_DTYPE_F16 1
#define YOLOPP_DTYPE_BF16 2
int yolopp_mish_forward(const void* in, void* out, int64_t n, int dtype, void* stream);
int yolopp_mish_backward(const void* grad_out, const void* in, void* grad_in, int64_t n, int dtype, void* stream);

/*
 * Bit-reproducible synthetic head tensors (bench / tests): element i of a level tensor gets
 * mean[g] + std[g] * z(seed, i) where g is the group of its attribute (0: box t0..t3, 1: objectness t4,
 * 2: class logits) and z is a fixed-point Irwin-Hall(4) variate from a counter-based hash — identical bits
 * from the C/numpy restatement used by the tests.
 *   out DEVICE (B, A*num_attrib, H, W) with hw = H*W; mean3/std3 HOST [3]
 */
int yolopp_synth_level(float* out, int32_t batch, int32_t num_anchors, int32_t num_attrib, int32_t hw,
                       const float* mean3, const float* std3, uint64_t seed, void* stream);

/* Canonical sigmoid / exp applied elementwise (tests: bit-parity of the transcendental with the oracle). */
int yolopp_sigmoid(const float* in, float* out, int64_t n, void* stream);
int yolopp_exp(const float* in, float* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YOLOPP_H_ */
