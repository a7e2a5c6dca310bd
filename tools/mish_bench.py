"""Mish forward / backward bandwidth on the B200: algorithmic bytes (fwd: read x + write y; bwd: read
dy, x + write dx) / event-timed duration, against the measured copy peak (MEASURED_PEAKS.json) and against
torch's own Mish (x * tanh(softplus(x)) fused by ATen's mish kernel) on the same tensors.
    python tools/mish_bench.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200')]
import torch, yolopp

peak = 6539.2
try:
    peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass


def timed(f, iters=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


rows = []
for name, dtype, n in (('f32 1Gi', torch.float32, 1 << 30), ('f32 act 64x256x76x76', torch.float32, 64 * 256 * 76 * 76),
                       ('f16 1Gi', torch.float16, 1 << 30), ('bf16 1Gi', torch.bfloat16, 1 << 30)):
    x = torch.randn(n, device='cuda', dtype=torch.float32).to(dtype)
    g = torch.randn(n, device='cuda', dtype=torch.float32).to(dtype)
    es = x.element_size()
    t_f = timed(lambda: yolopp.mish_forward(x))
    t_b = timed(lambda: yolopp.mish_backward(g, x))
    t_ref = timed(lambda: torch.nn.functional.mish(x))
    y = torch.empty_like(x)
    t_copy = timed(lambda: y.copy_(x))
    row = dict(case=name, elements=n, fwd_ms=t_f, fwd_gbs=2 * es * n / t_f / 1e6, fwd_frac_of_copy_peak=2 * es * n / t_f / 1e6 / peak,
               bwd_ms=t_b, bwd_gbs=3 * es * n / t_b / 1e6, bwd_frac_of_copy_peak=3 * es * n / t_b / 1e6 / peak,
               torch_mish_fwd_ms=t_ref, torch_copy_ms=t_copy, copy_gbs_here=2 * es * n / t_copy / 1e6)
    rows.append(row)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)
    del x, g, y
    torch.cuda.empty_cache()
out = dict(peak_gbs=peak, note='includes the torch.empty_like allocation of the output (caching allocator) inside each call', rows=rows)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
