#!/usr/bin/env python
"""bench.py — YOLOv4 decode+NMS throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            own arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N ...            reference arm: the reference's CPU implementation of the
                                                             path (oracle port, all host threads), rank 0 only

A "step" is one pass of the whole path (objectness top-k -> fused decode/score/threshold/binning -> per-class
NMS -> top-300) over one batch of 64 synthetic 608x608 head-output sets per GPU (BASELINE.json configs[1]:
score_thr 0.001, nms_pre 1000, iou 0.65, max_per_img 300, "COCO-like sparse" logits). Multi-GPU = the batch is
sharded, one process per GPU, no collective on the data path (torch.distributed is used only for the barrier
and the max-over-ranks of the timing). Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200'), os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = 'yolov4_decode_nms_images_per_sec_608_b64'
UNIT = 'images/s'
WORKLOADS = {
    # name -> (case key in tests/cases.py, per-GPU batch)
    'yolov4_608_b64_coco_sparse': ('csp608_sparse', 64),
    'yolov4_608_b64_dense': ('csp608_dense', 64),
    'yolov5_640_b128_sparse': ('csp640_sparse', 128),
    'yolov3_640_b128_sparse': ('v3_640_sparse', 128),
    'yolov4_1280_b128_sparse': ('csp1280_sparse', 128),  # configs[4]: 1024 images sharded over 8 GPUs
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='yolopp', choices=['yolopp', 'reference'])
    ap.add_argument('--workload', default='yolov4_608_b64_coco_sparse')
    ap.add_argument('--cpu-baseline-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--pipeline-depth', type=int, default=3,
                    help='batches in flight on separate CUDA streams (1 = strictly one batch after the other)')
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        try:
            with open(path) as f:
                return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def load_traffic():
    """dram read+write bytes per launch of the decode kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, 'profiles', 'decode_traffic.json')
    if os.path.isfile(path):
        try:
            with open(path) as f:
                return json.load(f).get('dram_bytes_per_launch')
        except Exception:
            pass
    return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU during the timed region."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_id):
        self.gpu_id = gpu_id
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu_id), f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))


def get_case(workload):
    import cases
    from yolopp import _capi as capi
    key, batch = WORKLOADS[workload]
    case = dict(cases.CASES[key])
    case['batch'] = batch
    return case


def config_dict(args, case, extra=None):
    cfg = dict(workload=args.workload, image_size=case['sizes'][0][0] * case['strides'][0],
               batch_per_gpu=case['batch'], global_batch=case['batch'] * args.gpus, num_classes=case['num_classes'],
               score_thr=case['score_thr'], nms_pre=case['nms_pre'], iou_threshold=case['nms']['iou_threshold'],
               max_per_img=case['max_per_img'], distribution=case['dist'] if isinstance(case['dist'], str) else 'custom',
               sharding=f'batch sharded over {args.gpus} GPU(s), no collective',
               l2_policy='inputs (495 MB/batch at 608^2 b64) exceed the 126 MB L2 and two input sets alternate; '
                         'no explicit flush')
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------------
def cpu_port_run(case, seconds=None, steps=None, warmup=0, threads=0):
    """Times oracle.get_bboxes (C restatement of the reference path, OpenMP over images) on one batch of the
    workload. Returns (images_per_s, ms_per_step list, cores, sample description)."""
    import cases
    from oracle import oracle
    p = cases.build_params(case)
    levels = cases.host_levels(case, p)
    sf = cases.scale_factors(case)
    cores = threads if threads > 0 else len(os.sched_getaffinity(0))
    times = []
    for _ in range(warmup):
        oracle.get_bboxes(p, levels, sf, num_threads=cores)
    t_begin = time.perf_counter()
    n = 0
    while True:
        t0 = time.perf_counter()
        oracle.get_bboxes(p, levels, sf, num_threads=cores)
        times.append(time.perf_counter() - t0)
        n += 1
        if steps is not None and n >= steps:
            break
        if steps is None and (time.perf_counter() - t_begin) >= seconds:
            break
    ips = case['batch'] * n / sum(times)
    sample = f'{n} x one batch of {case["batch"]} images of the workload ({sum(times):.1f} s of host work)'
    return ips, [t * 1e3 for t in times], cores, sample


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    case = get_case(args.workload)
    ips, ms, cores, sample = cpu_port_run(case, steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)))
    line = dict(impl='reference', metric=METRIC, value=ips, unit=UNIT, n_gpus=args.gpus, steps=len(ms),
                warmup=min(args.warmup, 1), ms_per_step=statistics.mean(ms), p50_ms=statistics.median(ms),
                higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=config_dict(args, case, dict(
                    note='reference arm = CPU port of the reference path (oracle/oracle.c, OpenMP over images); '
                         'the reference itself is Python+mmcv and cannot be installed here (mmcv-full absent)')),
                cpu_baseline=dict(value=ips, unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=ips, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
def run_yolopp(args):
    import torch
    import cases
    import yolopp
    from yolopp import _capi as capi
    from yolopp.ops import Session, Pipeline

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    case = get_case(args.workload)
    case['seed'] = case['seed'] + 1000 * rank  # every rank decodes its own shard of the global batch
    p = cases.build_params(case)
    # two input sets used alternately: consecutive steps never read the same 495 MB (each set alone exceeds L2)
    inputs = [yolopp.synth.synth_levels(p, case['seed'] + 7 * j, case['dist'], device=dev) for j in range(2)]
    levels = inputs[0]
    depth = max(1, args.pipeline_depth)
    pipe = Pipeline(p, depth, dev)
    sess = yolopp.ops.Session(p, dev)  # a batch that runs alone (batches_in_flight = 0): latency / per-kernel timing
    info = sess.info
    sf = None

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the clock sampler (an nvidia-smi child process polling every 100 ms) is started before the warm-up so that its
    # start-up cost on the host does not fall into the timed region; it keeps sampling through the timed region
    # and the load phase that follows it
    uuid = str(torch.cuda.get_device_properties(dev).uuid)
    sampler = ClockSampler(uuid if uuid.startswith('GPU-') else 'GPU-' + uuid)
    sampler.start()
    # ---- warm-up ----
    for i in range(max(3, args.warmup)):
        sess.run(inputs[i % 2], sf, profile=True)
    for i in range(2 * depth):
        pipe.submit(inputs[i % 2], sf)
    torch.cuda.synchronize(dev)

    # ---- timed region: K steps, device-resident inputs, issued round-robin on `depth` streams (software
    #      pipelining across batches); CUDA events: start on the caller's stream before the first submit, end
    #      after the caller's stream has joined every pipeline stream ----
    K = args.steps
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_done = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stage_acc = {n: 0.0 for n in capi.STAGE_NAMES}
    barrier()
    ev0.record()
    t_host0 = time.perf_counter()
    for i in range(K):
        slot = pipe.submit(inputs[i % 2], sf)
        ev_done[i].record(pipe.streams[slot])
    host_submit_ms = (time.perf_counter() - t_host0) * 1e3 / K  # host time to issue one step (not a GPU time)
    pipe.join()
    ev1.record()
    torch.cuda.synchronize(dev)
    barrier()
    done_ms = [ev0.elapsed_time(e) for e in ev_done]
    step_ms = [b - a for a, b in zip([0.0] + done_ms[:-1], done_ms)]
    total_ms = ev0.elapsed_time(ev1)

    # ---- latency of ONE batch (no overlap): K steps strictly one after the other on one stream ----
    n_lat = min(K, 50)
    lat_ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_lat + 1)]
    lat_ev[0].record()
    for i in range(n_lat):
        sess.run(inputs[i % 2], sf, profile=False)
        lat_ev[i + 1].record()
    torch.cuda.synchronize(dev)
    latency_ms = lat_ev[0].elapsed_time(lat_ev[n_lat]) / n_lat
    lat_each = sorted(lat_ev[i].elapsed_time(lat_ev[i + 1]) for i in range(n_lat))  # per-batch latency distribution

    # ---- per-kernel durations (events between the kernels of each step, same stream), separate loop ----
    n_prof = min(K, 20)
    for i in range(n_prof):
        sess.run(inputs[i % 2], sf, profile=True)
        torch.cuda.synchronize(dev)
        for n, v in sess.stage_ms().items():
            stage_acc[n] += v
    stage_ms = {n: v / n_prof for n, v in stage_acc.items()}
    # keep the GPU under the same load until the sampler has a few readings (it polls every 100 ms)
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        for i in range(20):
            pipe.submit(inputs[i % 2], sf)
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()

    if distributed:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms_max = float(t.item())
    else:
        total_ms_max = total_ms
    images = case['batch'] * world * K
    value = images / (total_ms_max * 1e-3)

    # ---- end to end through the public head API with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        head = yolopp.YOLOCSPHead(num_classes=case['num_classes'], test_cfg=cases.ref_cfg(case)) \
            if case['mode'] == capi.MODE_CSP else yolopp.YOLOV3Head(num_classes=case['num_classes'],
                                                                    test_cfg=cases.ref_cfg(case))
        metas = [dict(scale_factor=1.0) for _ in range(case['batch'])]
        host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in levels]
        for h, x in zip(host, levels):
            h.copy_(x)
        dev_in = [torch.empty_like(x) for x in levels]
        torch.cuda.synchronize(dev)
        h2d = sum(h.numel() * 4 for h in host)
        Ke = max(3, min(K, 10))

        def e2e_step():
            for d, h in zip(dev_in, host):
                d.copy_(h, non_blocking=True)
            return head.get_results_host(dev_in, metas)

        for _ in range(2):
            res = e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(Ke):
            res = e2e_step()
        e1.record()
        torch.cuda.synchronize(dev)
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max(e0.elapsed_time(e1), wall_ms)  # the host-side read of the result is inside the region
        if distributed:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        d2h = p.batch * p.capacity * (5 * 4 + 8) + (2 * p.batch + 1) * 4
        e2e = dict(value=case['batch'] * world * Ke / (e2e_ms * 1e-3), unit=UNIT, h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=d2h, steps=Ke, ms_per_step=e2e_ms / Ke,
                   api='yolopp.YOLOCSPHead.get_results_host (pinned host -> device copy of the head tensors, '
                       'custom op, one pinned device -> host copy of the detections)')
        assert len(res) == case['batch']

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (decode, TMA levels) ----
    peak, peak_src = load_peaks()
    alg_bytes = info.tma_bytes_per_image * p.batch
    dec_ms = stage_ms['decode_tma']
    achieved = alg_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else 0.0
    all_dec_bytes = (info.tma_bytes_per_image + info.ldg_bytes_per_image) * p.batch
    all_dec_ms = stage_ms['decode_tma'] + stage_ms['decode_ldg']
    roofline = dict(bound='hbm', kernel='decode_tma_kernel', achieved=achieved, peak=peak, unit='GB/s',
                    frac=achieved / peak,
                    traffic=load_traffic() if args.workload == 'yolov4_608_b64_coco_sparse' else None,  # captured for that workload only
                    peak_source=peak_src,
                    algorithmic_bytes_per_launch=alg_bytes, kernel_ms=dec_ms,
                    decode_all_levels=dict(bytes=all_dec_bytes, ms=all_dec_ms,
                                           achieved=all_dec_bytes / (all_dec_ms * 1e-3) / 1e9 if all_dec_ms > 0 else 0.0,
                                           frac=(all_dec_bytes / (all_dec_ms * 1e-3) / 1e9) / peak if all_dec_ms > 0 else 0.0),
                    stage_ms=stage_ms,
                    measured='CUDA events around each kernel on its stream, batches run one at a time (the pipelined '
                             'region overlaps kernels of different batches, so per-kernel durations are not defined there)')

    cpu_baseline = None
    if not args.no_cpu_baseline:
        base_case = get_case(args.workload)
        ips, ms, cores, sample = cpu_port_run(base_case, seconds=args.cpu_baseline_seconds)
        cpu_baseline = dict(value=ips, unit=UNIT, cores=cores, kind='port', sample=sample,
                            ms_per_batch=statistics.mean(ms))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=max(3, args.warmup),
                ms_per_step=total_ms_max / K, host_submit_ms_per_step=host_submit_ms, latency_ms=latency_ms, p50_ms=statistics.median(lat_each),
                p90_ms=lat_each[int(0.9 * (n_lat - 1))], p99_ms=lat_each[int(0.99 * (n_lat - 1))], higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', config=config_dict(args, case, dict(pipeline_depth=depth, pipeline='steps issued round-robin on '
                                                                  f'{depth} CUDA streams, one workspace per stream; latency_ms / p50_ms / p90_ms / p99_ms = one batch alone on one stream')), clocks=clocks, e2e=e2e,
                gpu_launches=info.kernel_launches * K, roofline=roofline, cpu_baseline=cpu_baseline)
    emit(line)
    if distributed:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL banners, warnings)
    was diverted to stderr at start-up."""
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # fd 1 -> stderr for the duration of the run (native libraries write to fd 1 directly)
    if args.impl == 'reference':
        return run_reference(args)
    return run_yolopp(args)


if __name__ == '__main__':
    sys.exit(main())
