"""torch custom ops over the C ABI (include/yolopp.h). CUDA dispatch key ONLY — calling them with CPU tensors
raises NotImplementedError from the dispatcher; there is no fallback implementation.

    torch.ops.yolopp.get_bboxes(pred_maps, scale_factors, params_blob)
        -> (dets[B,cap,5] f32, labels[B,cap] i64, anchors[B,cap] i32, rows[B,cap] i32, count[B] i32,
            num_candidates[B] i32, status[1] i32, cls_dets[B,cap,5] f32, cls_offsets[B,C+1] i32)
    torch.ops.yolopp.topk_conf(pred_maps, params_blob) -> topk_inds[B,R] i32          (parity tap)
    torch.ops.yolopp.decode(pred_maps, scale_factors, params_blob)
        -> (boxes[B,R,4] f32, scores[B,R,C] f32 (NaN = not a candidate), topk_inds[B,R] i32)
    torch.ops.yolopp.mish_forward(x) / mish_backward(grad_out, x)
    torch.ops.yolopp.coder_decode(bboxes, pred, stride, mode) -> decoded
    torch.ops.yolopp.sigmoid(x) / torch.ops.yolopp.exp(x)      (canonical transcendentals; tests)

torch here is plumbing: device memory (caching allocator), the current stream, the dispatcher.
"""
import ctypes

import torch

from . import _capi

_LIB_DEF = torch.library.Library('yolopp', 'DEF')
_LIB_DEF.define('get_bboxes(Tensor[] pred_maps, Tensor? scale_factors, Tensor params) -> '
                '(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)')
_LIB_DEF.define('topk_conf(Tensor[] pred_maps, Tensor params) -> Tensor')
_LIB_DEF.define('decode(Tensor[] pred_maps, Tensor? scale_factors, Tensor params) -> (Tensor, Tensor, Tensor)')
_LIB_DEF.define('mish_forward(Tensor input) -> Tensor')
_LIB_DEF.define('mish_backward(Tensor grad_out, Tensor input) -> Tensor')
_LIB_DEF.define('coder_decode(Tensor bboxes, Tensor pred, float stride, int mode) -> Tensor')
_LIB_DEF.define('sigmoid(Tensor x) -> Tensor')
_LIB_DEF.define('exp(Tensor x) -> Tensor')


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _params_from_blob(blob):
    p = _capi.YoloppParams()
    raw = blob.cpu().numpy().tobytes()
    if len(raw) != ctypes.sizeof(p):
        raise ValueError('params blob has the wrong size')
    ctypes.memmove(ctypes.byref(p), raw, len(raw))
    return p


def params_to_blob(p):
    """yolopp_params struct -> uint8 CPU tensor (the custom op's schema only carries tensors and scalars)."""
    return torch.frombuffer(bytearray(bytes(p)), dtype=torch.uint8).clone()


def _norm_dev(device):
    dev = torch.device(device)
    if dev.type == 'cuda' and dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    return dev


def check_maps(p, pred_maps, device=None):
    """Validates the level tensors against the params the way the reference's asserts would (yolocsp_head.py:216,
    yolov4_bbox_coder.py:50-51) plus what a raw-pointer interface must insist on: CUDA, float32, one device, the
    logical shape (B, A*(5+C), H, W), and memory that is either NCHW-contiguous or channels-last contiguous.
    Returns (maps, layout): tensors whose data_ptr() may be handed to the C ABI (copies only where a tensor is
    neither) and the YOLOPP_LAYOUT_* they share."""
    L = p.num_levels
    if len(pred_maps) != L:
        raise AssertionError(f'expected {L} prediction maps, got {len(pred_maps)}')
    dev = pred_maps[0].device if device is None else torch.device(device)
    if dev.type == 'cuda' and dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    if dev.type != 'cuda':
        raise NotImplementedError('yolopp runs on CUDA tensors only (no CPU fallback)')
    nhwc = []
    for l, t in enumerate(pred_maps):
        if t.dtype != torch.float32:
            raise TypeError(f'pred_maps[{l}] must be float32 (got {t.dtype}); the reference path is fp32')
        if tuple(t.shape) != p.level_shape(l):
            raise AssertionError(f'pred_maps[{l}] has shape {tuple(t.shape)}, expected {p.level_shape(l)}')
        if t.device != dev:
            raise AssertionError('all prediction maps must be on the same device')
        nhwc.append((not t.is_contiguous()) and t.is_contiguous(memory_format=torch.channels_last))
    if all(nhwc):
        return list(pred_maps), _capi.LAYOUT_NHWC  # consumed in place: the row-driven decode reads channels-last
    return [t.contiguous() for t in pred_maps], _capi.LAYOUT_NCHW


def alloc_outputs(p, dev, zero=False):
    """The fixed-capacity output block of one call (include/yolopp.h yolopp_outputs)."""
    B, cap, C = p.batch, p.capacity, p.eff_classes
    mk = torch.zeros if zero else torch.empty
    out = dict(dets=mk((B, cap, 5), dtype=torch.float32, device=dev), labels=mk((B, cap), dtype=torch.int64, device=dev),
               anchors=mk((B, cap), dtype=torch.int32, device=dev), rows=mk((B, cap), dtype=torch.int32, device=dev),
               count=mk((B, ), dtype=torch.int32, device=dev), num_candidates=mk((B, ), dtype=torch.int32, device=dev),
               status=mk((1, ), dtype=torch.int32, device=dev),
               cls_dets=mk((B, cap, 5), dtype=torch.float32, device=dev),
               cls_offsets=mk((B, C + 1), dtype=torch.int32, device=dev))
    c_out = _capi.YoloppOutputs(*[out[k].data_ptr() for k in OUTPUT_NAMES])
    return out, c_out


OUTPUT_NAMES = ('dets', 'labels', 'anchors', 'rows', 'count', 'num_candidates', 'status', 'cls_dets', 'cls_offsets')


def _with_layout(p, layout):
    if p.layout == layout:
        return p
    q = type(p).from_buffer_copy(p)
    q.layout = layout
    return q


def _get_bboxes_cuda(pred_maps, scale_factors, params):
    lib = _capi.load_library()
    p = _params_from_blob(params)
    maps, layout = check_maps(p, pred_maps)
    p = _with_layout(p, layout)
    dev = maps[0].device
    B, L = p.batch, p.num_levels
    with torch.cuda.device(dev):
        ws_bytes = lib.yolopp_workspace_bytes(ctypes.byref(p))
        if ws_bytes == 0:
            raise ValueError('yolopp: invalid or unsupported configuration (see include/yolopp.h limits)')
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out, c_out = alloc_outputs(p, dev)
        sf = None
        if p.rescale:
            if scale_factors is None:
                raise ValueError('rescale=True needs scale_factors')
            sf = scale_factors.to(device=dev, dtype=torch.float32).contiguous()
            if tuple(sf.shape) != (B, 4):
                raise AssertionError('scale_factors must be (B, 4)')
        ptrs = (ctypes.c_void_p * L)(*[m.data_ptr() for m in maps])
        rc = lib.yolopp_get_bboxes(ctypes.byref(p), ptrs, ctypes.c_void_p(sf.data_ptr() if sf is not None else None),
                                   ctypes.byref(c_out), ctypes.c_void_p(ws.data_ptr()), ws_bytes, _stream())
        _capi.check(rc, 'yolopp_get_bboxes')
        # the workspace / inputs must outlive the asynchronous kernels: tie them to the current stream
        cur = torch.cuda.current_stream()
        for t in maps + [ws] + ([sf] if sf is not None else []):
            t.record_stream(cur)
    return tuple(out[k] for k in OUTPUT_NAMES)


def _coder_decode_cuda(bboxes, pred, stride, mode):
    lib = _capi.load_library()
    assert pred.size(0) == bboxes.size(0)
    assert pred.size(-1) == bboxes.size(-1) == 4
    if bboxes.dtype != torch.float32 or pred.dtype != torch.float32:
        raise TypeError('coder_decode needs float32 tensors')
    b = bboxes.expand_as(pred).contiguous() if bboxes.shape != pred.shape else bboxes.contiguous()
    q = pred.contiguous()
    out = torch.empty_like(q)
    with torch.cuda.device(q.device):
        rc = lib.yolopp_coder_decode(int(mode), ctypes.c_void_p(b.data_ptr()), ctypes.c_void_p(q.data_ptr()),
                                     ctypes.c_float(float(stride)), q.numel() // 4, ctypes.c_void_p(out.data_ptr()),
                                     _stream())
    _capi.check(rc, 'yolopp_coder_decode')
    return out


def _stage_call(pred_maps, scale_factors, params, which):
    lib = _capi.load_library()
    p = _params_from_blob(params)
    maps, layout = check_maps(p, pred_maps)
    p = _with_layout(p, layout)
    dev = maps[0].device
    info = _capi.describe(p)
    B, R, C, L = p.batch, info.rows_per_image, p.eff_classes, p.num_levels
    with torch.cuda.device(dev):
        ws_bytes = lib.yolopp_workspace_bytes(ctypes.byref(p))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        ptrs = (ctypes.c_void_p * L)(*[m.data_ptr() for m in maps])
        inds = torch.empty((B, R), dtype=torch.int32, device=dev)
        if which == 'topk':
            rc = lib.yolopp_topk_conf(ctypes.byref(p), ptrs, ctypes.c_void_p(inds.data_ptr()),
                                      ctypes.c_void_p(ws.data_ptr()), ws_bytes, _stream())
            _capi.check(rc, 'yolopp_topk_conf')
            res = inds
        else:
            sf = None
            if p.rescale:
                sf = scale_factors.to(device=dev, dtype=torch.float32).contiguous()
                assert tuple(sf.shape) == (B, 4)
            boxes = torch.empty((B, R, 4), dtype=torch.float32, device=dev)
            scores = torch.empty((B, R, C), dtype=torch.float32, device=dev)
            rc = lib.yolopp_decode(ctypes.byref(p), ptrs, ctypes.c_void_p(sf.data_ptr() if sf is not None else None),
                                   ctypes.c_void_p(boxes.data_ptr()), ctypes.c_void_p(scores.data_ptr()),
                                   ctypes.c_void_p(inds.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws_bytes, _stream())
            _capi.check(rc, 'yolopp_decode')
            res = (boxes, scores, inds)
            if sf is not None:
                sf.record_stream(torch.cuda.current_stream())
        cur = torch.cuda.current_stream()
        for t in maps + [ws]:
            t.record_stream(cur)
    return res


def _topk_conf_cuda(pred_maps, params):
    return _stage_call(pred_maps, None, params, 'topk')


def _decode_cuda(pred_maps, scale_factors, params):
    return _stage_call(pred_maps, scale_factors, params, 'decode')


_MISH_DTYPES = {torch.float32: _capi.DTYPE_F32, torch.float16: _capi.DTYPE_F16, torch.bfloat16: _capi.DTYPE_BF16}


def _mish_forward_cuda(inp):
    lib = _capi.load_library()
    if inp.dtype not in _MISH_DTYPES:
        raise TypeError(f'mish: unsupported dtype {inp.dtype} (float32 / float16 / bfloat16)')
    x = inp.contiguous()  # mish.py:24-25
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.yolopp_mish_forward(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), x.numel(),
                                     _MISH_DTYPES[x.dtype], _stream())
    _capi.check(rc, 'yolopp_mish_forward')
    return out


def _mish_backward_cuda(grad_out, inp):
    lib = _capi.load_library()
    if inp.dtype not in _MISH_DTYPES or grad_out.dtype != inp.dtype:
        raise TypeError('mish_backward: grad_out and input must share a float32 / float16 / bfloat16 dtype')
    x, g = inp.contiguous(), grad_out.contiguous()  # mish.py:33-34
    assert x.numel() == g.numel()
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = lib.yolopp_mish_backward(ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(x.data_ptr()),
                                      ctypes.c_void_p(out.data_ptr()), x.numel(), _MISH_DTYPES[x.dtype], _stream())
    _capi.check(rc, 'yolopp_mish_backward')
    return out


def _unary(name):

    def f(x):
        lib = _capi.load_library()
        if x.dtype != torch.float32:
            raise TypeError('float32 only')
        xc = x.contiguous()
        out = torch.empty_like(xc)
        with torch.cuda.device(xc.device):
            rc = getattr(lib, name)(ctypes.c_void_p(xc.data_ptr()), ctypes.c_void_p(out.data_ptr()), xc.numel(),
                                    _stream())
        _capi.check(rc, name)
        return out

    return f


_LIB_IMPL = torch.library.Library('yolopp', 'IMPL')
_LIB_IMPL.impl('get_bboxes', _get_bboxes_cuda, 'CUDA')
_LIB_IMPL.impl('coder_decode', _coder_decode_cuda, 'CUDA')
_LIB_IMPL.impl('topk_conf', _topk_conf_cuda, 'CUDA')
_LIB_IMPL.impl('decode', _decode_cuda, 'CUDA')
_LIB_IMPL.impl('mish_forward', _mish_forward_cuda, 'CUDA')
_LIB_IMPL.impl('mish_backward', _mish_backward_cuda, 'CUDA')
_LIB_IMPL.impl('sigmoid', _unary('yolopp_sigmoid'), 'CUDA')
_LIB_IMPL.impl('exp', _unary('yolopp_exp'), 'CUDA')


# ----------------------------------------------------------------------------------------------------
# thin python entry points
# ----------------------------------------------------------------------------------------------------
def get_bboxes_raw(params, pred_maps, scale_factors=None):
    """Runs the whole path; returns the fixed-capacity device tensors (no host sync)."""
    blob = params if isinstance(params, torch.Tensor) else params_to_blob(params)
    return dict(zip(OUTPUT_NAMES, torch.ops.yolopp.get_bboxes(list(pred_maps), scale_factors, blob)))


def topk_conf(params, pred_maps):
    """Parity tap: `topk_inds` of yolocsp_head.py:350-355 / yolo_head.py:281-302, (B, R) int32."""
    blob = params if isinstance(params, torch.Tensor) else params_to_blob(params)
    return torch.ops.yolopp.topk_conf(list(pred_maps), blob)


def decode(params, pred_maps, scale_factors=None):
    """Parity tap: what enters multiclass_nms — boxes (B,R,4), scores (B,R,C) with NaN where (row, class) is not a
    candidate, topk_inds (B,R)."""
    blob = params if isinstance(params, torch.Tensor) else params_to_blob(params)
    return torch.ops.yolopp.decode(list(pred_maps), scale_factors, blob)


def coder_decode(bboxes, pred, stride, mode):
    return torch.ops.yolopp.coder_decode(bboxes, pred, float(stride), int(mode))


def sigmoid(x):
    return torch.ops.yolopp.sigmoid(x)


def exp(x):
    return torch.ops.yolopp.exp(x)


class Session:
    """Pre-allocated workspace + output block for one configuration (serving loops / the benchmark). `run()` looks up
    (or creates) a yolopp_plan for the exact buffers it is given — validation, workspace layout, tensor maps and
    grid sizes are derived once — and is then ONE C-ABI call (three kernel launches): no allocation, no host sync.
    `run(profile=True)` additionally records the per-stage CUDA events; `stage_ms()` reads them after a
    synchronize. Plans keep their input tensors alive; at most `max_plans` are cached (least recently used first
    out, released to the allocator through record_stream)."""

    def __init__(self, params, device='cuda', max_plans=8):
        self.lib = _capi.load_library()
        self.p = params
        self.dev = _norm_dev(device)
        self.max_plans = int(max_plans)
        with torch.cuda.device(self.dev):
            self.ws_bytes = self.lib.yolopp_workspace_bytes(ctypes.byref(params))
            if self.ws_bytes == 0:
                raise ValueError('yolopp: invalid or unsupported configuration')
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev)
            self.out, self._outs = alloc_outputs(params, self.dev, zero=True)
        self._events = None
        self._plans = {}   # key -> [handle, tensors kept alive, last stream]
        self.info = _capi.describe(params)

    def __del__(self):
        try:
            for h, _, _ in self._plans.values():
                self.lib.yolopp_plan_destroy(h)
        except Exception:
            pass

    def _make_events(self):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(_capi.NUM_STAGE_EVENTS)]
        for e in evs:
            e.record()  # forces creation of the underlying cudaEvent_t
        torch.cuda.synchronize(self.dev)
        return evs

    def _plan(self, pred_maps, scale_factors):
        # (pointer, shape, strides) per map: a view of a known tensor with another shape must not hit its plan
        key = tuple((m.data_ptr(), m.shape, m.stride()) for m in pred_maps) + \
            ((scale_factors.data_ptr(), ) if scale_factors is not None else ())
        ent = self._plans.get(key)
        if ent is not None:
            return ent
        maps, layout = check_maps(self.p, pred_maps, self.dev)
        if any(a is not b for a, b in zip(maps, pred_maps)):
            raise ValueError('Session.run needs NCHW-contiguous or channels-last contiguous level tensors')
        p = _with_layout(self.p, layout)
        if layout != self.p.layout:
            with torch.cuda.device(self.dev):
                need = self.lib.yolopp_workspace_bytes(ctypes.byref(p))
            if need > self.ws_bytes:
                raise ValueError('workspace of this Session is too small for the other layout')
        sf = scale_factors
        if self.p.rescale:
            if sf is None or sf.dtype != torch.float32 or not sf.is_contiguous() or sf.device != self.dev or \
                    tuple(sf.shape) != (self.p.batch, 4):
                raise ValueError('rescale=True needs contiguous float32 scale_factors (B, 4) on the device')
        if len(self._plans) >= self.max_plans:
            old_key = next(iter(self._plans))
            h, refs, st = self._plans.pop(old_key)
            self.lib.yolopp_plan_destroy(h)
            if st is not None:
                for t in refs:
                    t.record_stream(st)  # the allocator may reuse them once that stream has passed this point
        ptrs = (ctypes.c_void_p * p.num_levels)(*[m.data_ptr() for m in maps])
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.dev):
            rc = self.lib.yolopp_plan_create(ctypes.byref(p), ptrs, ctypes.c_void_p(sf.data_ptr() if sf is not None else None),
                                             ctypes.byref(self._outs), ctypes.c_void_p(self.ws.data_ptr()), self.ws_bytes,
                                             ctypes.byref(handle))
        _capi.check(rc, 'yolopp_plan_create')
        ent = [handle, list(maps) + ([sf] if sf is not None else []), None]
        self._plans[key] = ent
        return ent

    def run(self, pred_maps, scale_factors=None, profile=False, stream=None, _sp=None):
        ent = self._plan(pred_maps, scale_factors)
        st = torch.cuda.current_stream(self.dev) if stream is None else stream
        ent[2] = st
        sp = _sp if _sp is not None else ctypes.c_void_p(st.cuda_stream)
        if profile:
            if self._events is None:
                self._events = self._make_events()
            evp = (ctypes.c_void_p * len(self._events))(*[e.cuda_event for e in self._events])
            rc = self.lib.yolopp_plan_run_profiled(ent[0], sp, evp, len(self._events))
        else:
            rc = self.lib.yolopp_plan_run(ent[0], sp)
        if rc != 0:
            _capi.check(rc, 'yolopp_plan_run')
        return self.out

    def stage_ms(self):
        """dict stage -> milliseconds of the last profiled run (call after torch.cuda.synchronize())."""
        ev = self._events
        return {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(_capi.STAGE_NAMES)}


class Pipeline:
    """Software pipelining ACROSS batches for serving loops: `depth` Sessions (own workspace + output block each)
    on `depth` CUDA streams, used round-robin. The per-image kernels of the path (objectness top-k, NMS: one CTA
    per image) leave more than half of the SMs idle and are latency bound, the decode kernel is HBM bound; with
    consecutive batches on different streams the top-k / NMS of one batch run beside the decode of another.

        pipe = Pipeline(params, depth=3)
        t = pipe.submit(pred_maps)        # asynchronous; the Session's plan keeps pred_maps alive
        out = pipe.result(t)              # waits (host) for that batch; dict of fixed-capacity device tensors

    A slot's outputs are overwritten when the slot is reused, i.e. `depth` submits later.
    """

    def __init__(self, params, depth=3, device='cuda'):
        self.depth = int(depth)
        self.dev = _norm_dev(device)
        p = type(params).from_buffer_copy(params)
        p.batches_in_flight = self.depth  # scheduling hint: decode CTAs retire progressively, neighbours overlap
        self.sessions = [Session(p, device) for _ in range(self.depth)]
        with torch.cuda.device(self.dev):
            self.streams = [torch.cuda.Stream(self.dev) for _ in range(self.depth)]
            self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self._sp = [ctypes.c_void_p(s.cuda_stream) for s in self.streams]
        self.n = 0

    def submit(self, pred_maps, scale_factors=None, inputs_ready=False):
        """Queues one batch on the next slot. `inputs_ready=True`: the caller guarantees that the inputs are complete
        (e.g. produced earlier and synchronised) — skips making the slot's stream wait for the caller's stream."""
        slot = self.n % self.depth
        self.n += 1
        st = self.streams[slot]
        if not inputs_ready:
            st.wait_stream(torch.cuda.current_stream(self.dev))  # the inputs were produced on the caller's stream
        self.sessions[slot].run(pred_maps, scale_factors, stream=st, _sp=self._sp[slot])
        self.done[slot].record(st)
        return slot

    def result(self, ticket):
        self.done[ticket].synchronize()
        return self.sessions[ticket].out

    def join(self):
        """Makes the caller's current stream wait for everything submitted so far (no host sync)."""
        cur = torch.cuda.current_stream(self.dev)
        for st in self.streams:
            cur.wait_stream(st)


class HostPipeline:
    """Host buffers in, host results out, double buffered: slot i owns device copies of the level tensors, a Session
    and a pinned result block, all driven by ONE stream per slot (H2D -> path -> D2H in order). With two slots
    the host->device copy of batch i+1 (H2D copy engine) overlaps the kernels and the device->host copy of batch i
    (SMs, D2H copy engine).

        hp = HostPipeline(params, depth=2)
        t = hp.submit(host_levels)        # pinned (B, A*(5+C), H, W) float32 CPU tensors; returns at once
        res = hp.result(t)                # list[(dets ndarray (n,5), labels ndarray (n,))], independent copies
    """

    def __init__(self, params, depth=2, device='cuda'):
        self.depth = int(depth)
        self.dev = _norm_dev(device)
        self.p = params
        B, cap = params.batch, params.capacity
        self.slots = []
        with torch.cuda.device(self.dev):
            for _ in range(self.depth):
                dev_in = [torch.empty(params.level_shape(l), dtype=torch.float32, device=self.dev)
                          for l in range(params.num_levels)]
                self.slots.append(dict(
                    dev_in=dev_in, sess=Session(params, self.dev), stream=torch.cuda.Stream(self.dev),
                    done=torch.cuda.Event(),
                    h_dets=torch.empty((B, cap, 5), dtype=torch.float32).pin_memory(),
                    h_labels=torch.empty((B, cap), dtype=torch.int64).pin_memory(),
                    h_meta=torch.empty((2 * B + 1, ), dtype=torch.int32).pin_memory()))
        self.n = 0
        self.h2d_bytes = sum(4 * int(torch.tensor(params.level_shape(l)).prod()) for l in range(params.num_levels))
        self.d2h_bytes = B * cap * (5 * 4 + 8) + (2 * B + 1) * 4

    def submit(self, host_levels):
        s = self.slots[self.n % self.depth]
        ticket = self.n % self.depth
        self.n += 1
        B = self.p.batch
        with torch.cuda.stream(s['stream']):
            for d, h in zip(s['dev_in'], host_levels):
                d.copy_(h, non_blocking=True)
            out = s['sess'].run(s['dev_in'], None, stream=s['stream'])
            s['h_dets'].copy_(out['dets'], non_blocking=True)
            s['h_labels'].copy_(out['labels'], non_blocking=True)
            s['h_meta'][:B].copy_(out['count'], non_blocking=True)
            s['h_meta'][B:2 * B].copy_(out['num_candidates'], non_blocking=True)
            s['h_meta'][2 * B:].copy_(out['status'], non_blocking=True)
            s['done'].record(s['stream'])
        return ticket

    def result(self, ticket):
        s = self.slots[ticket]
        s['done'].synchronize()
        B = self.p.batch
        meta = s['h_meta'].numpy()
        if int(meta[-1]) != 0:
            raise RuntimeError(f'yolopp_get_bboxes: {_capi.load_library().yolopp_strerror(int(meta[-1])).decode()}')
        d, l, cnt = s['h_dets'].numpy(), s['h_labels'].numpy(), meta[:B]
        return [(d[b, :cnt[b]].copy(), l[b, :cnt[b]].copy()) for b in range(B)]
