"""torch custom ops over the C ABI (include/yolopp.h). CUDA dispatch key ONLY — calling them with CPU tensors
raises NotImplementedError from the dispatcher; there is no fallback implementation.

    torch.ops.yolopp.get_bboxes(pred_maps, scale_factors, params_blob)
        -> (dets[B,cap,5] f32, labels[B,cap] i64, anchors[B,cap] i32, rows[B,cap] i32, count[B] i32,
            num_candidates[B] i32, status[1] i32)
    torch.ops.yolopp.coder_decode(bboxes, pred, stride, mode) -> decoded
    torch.ops.yolopp.sigmoid(x) / torch.ops.yolopp.exp(x)      (canonical transcendentals; tests)

torch here is plumbing: device memory (caching allocator), the current stream, the dispatcher.
"""
import ctypes

import torch

from . import _capi

_LIB_DEF = torch.library.Library('yolopp', 'DEF')
_LIB_DEF.define('get_bboxes(Tensor[] pred_maps, Tensor? scale_factors, Tensor params) -> '
                '(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)')
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


def _get_bboxes_cuda(pred_maps, scale_factors, params):
    lib = _capi.load_library()
    p = _params_from_blob(params)
    L = p.num_levels
    if len(pred_maps) != L:
        raise AssertionError(f'expected {L} prediction maps, got {len(pred_maps)}')
    dev = pred_maps[0].device
    maps = []
    for l, t in enumerate(pred_maps):
        if t.dtype != torch.float32:
            raise TypeError(f'pred_maps[{l}] must be float32 (got {t.dtype}); the reference path is fp32')
        if tuple(t.shape) != p.level_shape(l):
            raise AssertionError(f'pred_maps[{l}] has shape {tuple(t.shape)}, expected {p.level_shape(l)}')
        if t.device != dev:
            raise AssertionError('all prediction maps must be on the same device')
        maps.append(t.contiguous())  # NCHW contiguous (channels_last inputs are re-laid out)
    B, cap = p.batch, p.capacity
    with torch.cuda.device(dev):
        ws_bytes = lib.yolopp_workspace_bytes(ctypes.byref(p))
        if ws_bytes == 0:
            raise ValueError('yolopp: invalid or unsupported configuration (see include/yolopp.h limits)')
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        dets = torch.empty((B, cap, 5), dtype=torch.float32, device=dev)
        labels = torch.empty((B, cap), dtype=torch.int64, device=dev)
        anchors = torch.empty((B, cap), dtype=torch.int32, device=dev)
        rows = torch.empty((B, cap), dtype=torch.int32, device=dev)
        count = torch.empty((B, ), dtype=torch.int32, device=dev)
        ncand = torch.empty((B, ), dtype=torch.int32, device=dev)
        status = torch.empty((1, ), dtype=torch.int32, device=dev)
        sf = None
        if p.rescale:
            if scale_factors is None:
                raise ValueError('rescale=True needs scale_factors')
            sf = scale_factors.to(device=dev, dtype=torch.float32).contiguous()
            if tuple(sf.shape) != (B, 4):
                raise AssertionError('scale_factors must be (B, 4)')
        ptrs = (ctypes.c_void_p * L)(*[m.data_ptr() for m in maps])
        out = _capi.YoloppOutputs(dets.data_ptr(), labels.data_ptr(), anchors.data_ptr(), rows.data_ptr(),
                                  count.data_ptr(), ncand.data_ptr(), status.data_ptr())
        rc = lib.yolopp_get_bboxes(ctypes.byref(p), ptrs, ctypes.c_void_p(sf.data_ptr() if sf is not None else None),
                                   ctypes.byref(out), ctypes.c_void_p(ws.data_ptr()), ws_bytes, _stream())
        _capi.check(rc, 'yolopp_get_bboxes')
        # the workspace / inputs must outlive the asynchronous kernels: tie them to the current stream
        cur = torch.cuda.current_stream()
        for t in maps + [ws] + ([sf] if sf is not None else []):
            t.record_stream(cur)
    return dets, labels, anchors, rows, count, ncand, status


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
_LIB_IMPL.impl('sigmoid', _unary('yolopp_sigmoid'), 'CUDA')
_LIB_IMPL.impl('exp', _unary('yolopp_exp'), 'CUDA')


# ----------------------------------------------------------------------------------------------------
# thin python entry points
# ----------------------------------------------------------------------------------------------------
def get_bboxes_raw(params, pred_maps, scale_factors=None):
    """Runs the whole path; returns the fixed-capacity device tensors (no host sync)."""
    blob = params if isinstance(params, torch.Tensor) else params_to_blob(params)
    names = ('dets', 'labels', 'anchors', 'rows', 'count', 'num_candidates', 'status')
    return dict(zip(names, torch.ops.yolopp.get_bboxes(list(pred_maps), scale_factors, blob)))


def coder_decode(bboxes, pred, stride, mode):
    return torch.ops.yolopp.coder_decode(bboxes, pred, float(stride), int(mode))


def sigmoid(x):
    return torch.ops.yolopp.sigmoid(x)


def exp(x):
    return torch.ops.yolopp.exp(x)


class Session:
    """Pre-allocated workspace + output block for one configuration (serving loops / the benchmark):
    `run()` is a single C-ABI call, no allocation, no host sync. `run(profile=True)` additionally records the
    per-stage CUDA events (yolopp_get_bboxes_profiled); `stage_ms()` reads them after a synchronize."""

    def __init__(self, params, device='cuda'):
        self.lib = _capi.load_library()
        self.p = params
        self.dev = torch.device(device)
        B, cap = params.batch, params.capacity
        with torch.cuda.device(self.dev):
            self.ws_bytes = self.lib.yolopp_workspace_bytes(ctypes.byref(params))
            if self.ws_bytes == 0:
                raise ValueError('yolopp: invalid or unsupported configuration')
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev)
            self.out = dict(
                dets=torch.zeros((B, cap, 5), dtype=torch.float32, device=self.dev),
                labels=torch.zeros((B, cap), dtype=torch.int64, device=self.dev),
                anchors=torch.zeros((B, cap), dtype=torch.int32, device=self.dev),
                rows=torch.zeros((B, cap), dtype=torch.int32, device=self.dev),
                count=torch.zeros((B, ), dtype=torch.int32, device=self.dev),
                num_candidates=torch.zeros((B, ), dtype=torch.int32, device=self.dev),
                status=torch.zeros((1, ), dtype=torch.int32, device=self.dev))
        o = self.out
        self._outs = _capi.YoloppOutputs(o['dets'].data_ptr(), o['labels'].data_ptr(), o['anchors'].data_ptr(),
                                         o['rows'].data_ptr(), o['count'].data_ptr(), o['num_candidates'].data_ptr(),
                                         o['status'].data_ptr())
        self._events = None
        self.info = _capi.describe(params)

    def _make_events(self):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(_capi.NUM_STAGE_EVENTS)]
        for e in evs:
            e.record()  # forces creation of the underlying cudaEvent_t
        torch.cuda.synchronize(self.dev)
        return evs

    def run(self, pred_maps, scale_factors=None, profile=False):
        p = self.p
        ptrs = (ctypes.c_void_p * p.num_levels)(*[m.data_ptr() for m in pred_maps])
        sf = ctypes.c_void_p(scale_factors.data_ptr() if scale_factors is not None else None)
        with torch.cuda.device(self.dev):
            if profile:
                if self._events is None:
                    self._events = self._make_events()
                evp = (ctypes.c_void_p * len(self._events))(*[e.cuda_event for e in self._events])
                rc = self.lib.yolopp_get_bboxes_profiled(ctypes.byref(p), ptrs, sf, ctypes.byref(self._outs),
                                                         ctypes.c_void_p(self.ws.data_ptr()), self.ws_bytes, _stream(),
                                                         evp, len(self._events))
            else:
                rc = self.lib.yolopp_get_bboxes(ctypes.byref(p), ptrs, sf, ctypes.byref(self._outs),
                                                ctypes.c_void_p(self.ws.data_ptr()), self.ws_bytes, _stream())
        _capi.check(rc, 'yolopp_get_bboxes')
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
        t = pipe.submit(pred_maps)        # asynchronous; pred_maps must stay alive until the ticket is done
        out = pipe.result(t)              # waits (host) for that batch; dict of fixed-capacity device tensors

    A slot's outputs are overwritten when the slot is reused, i.e. `depth` submits later.
    """

    def __init__(self, params, depth=3, device='cuda'):
        self.depth = int(depth)
        self.dev = torch.device(device)
        p = type(params).from_buffer_copy(params)
        p.batches_in_flight = self.depth  # scheduling hint: decode CTAs retire progressively, neighbours overlap
        self.sessions = [Session(p, device) for _ in range(self.depth)]
        with torch.cuda.device(self.dev):
            self.streams = [torch.cuda.Stream(self.dev) for _ in range(self.depth)]
            self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self.n = 0

    def submit(self, pred_maps, scale_factors=None):
        slot = self.n % self.depth
        self.n += 1
        st = self.streams[slot]
        st.wait_stream(torch.cuda.current_stream(self.dev))  # the inputs were produced on the caller's stream
        with torch.cuda.stream(st):
            self.sessions[slot].run(pred_maps, scale_factors)
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
