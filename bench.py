#!/usr/bin/env python
"""bench.py — YOLOv4 decode+NMS throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            own arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N ...            reference arm: the reference's CPU implementation of the
                                                             path (oracle port, all host threads), rank 0 only

A "step" is one pass of the whole path (objectness top-k -> fused decode/score/threshold -> per-image NMS ->
top-300) over one batch of 64 synthetic 608x608 head-output sets per GPU (BASELINE.json configs[1]: score_thr 0.001,
nms_pre 1000, iou 0.65, max_per_img 300, "COCO-like sparse" logits). Multi-GPU = the batch is sharded, one process
per GPU, no collective on the data path (torch.distributed is used only for the barrier and the max-over-ranks of
the timing). `--workload yolov4_1280_b1024_sparse` is BASELINE.json configs[4]: a FIXED global batch of 1024 images
at 1280^2 split over the GPUs ("scaling": "strong"; a step = the whole global batch, processed 128 images at a
time per GPU). Prints ONE JSON line on rank 0.
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
for _p in (ROOT, os.path.join(ROOT, 'mmdet-yolov4_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = 'yolov4_decode_nms_images_per_sec_608_b64'
UNIT = 'images/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='yolopp', choices=['yolopp', 'reference'])
    ap.add_argument('--workload', default='yolov4_608_b64_coco_sparse')
    ap.add_argument('--layout', default='nchw', choices=['nchw', 'nhwc'],
                    help='memory layout of the synthetic head tensors (nhwc = channels-last, row-driven decode)')
    ap.add_argument('--cpu-baseline-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-verify', action='store_true')
    ap.add_argument('--no-affinity', action='store_true')
    ap.add_argument('--pipeline-depth', type=int, default=6,
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
    import workloads
    return dict(workloads.WORKLOADS[workload])


def config_dict(args, case):
    """Identical for both arms (the driver compares them); free-text remarks live in the line's `notes`."""
    strong = 'global_batch' in case
    return dict(workload=args.workload, image_size=case['sizes'][0][0] * case['strides'][0],
                batch_per_gpu=(case['global_batch'] // args.gpus) if strong else case['batch'],
                global_batch=case['global_batch'] if strong else case['batch'] * args.gpus,
                images_per_call=case['batch'], num_classes=case['num_classes'],
                score_thr=case['score_thr'], nms_pre=case['nms_pre'], iou_threshold=case['nms']['iou_threshold'],
                max_per_img=case['max_per_img'], distribution=case['dist'] if isinstance(case['dist'], str) else 'custom',
                layout=args.layout,
                sharding=f'batch sharded over {args.gpus} GPU(s), no collective',
                l2_policy='inputs (495 MB/batch at 608^2 b64) exceed the 126 MB L2 and two input sets alternate; '
                          'no explicit flush')


# ------------------------------------------------------------------------------------------------------
# host placement: one rank = one GPU = its own share of the host cores
# ------------------------------------------------------------------------------------------------------
def pin_rank_to_cores(local_rank, local_world, gpu_pci_bus_id=None):
    """Gives every rank its own slice of the cores this job may use (ranks otherwise all float over the same cores
    and allocate their pinned staging buffers wherever they happen to run). Prefers the cores of the GPU's NUMA node
    when sysfs exposes one. Returns a description for the JSON line."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        node = None
        if gpu_pci_bus_id:
            path = f'/sys/bus/pci/devices/{gpu_pci_bus_id.lower()}/numa_node'
            if os.path.isfile(path):
                with open(path) as f:
                    v = int(f.read().strip())
                node = v if v >= 0 else None
        pool = cores
        if node is not None:
            try:
                with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
                    nc = set()
                    for part in f.read().strip().split(','):
                        lo, _, hi = part.partition('-')
                        nc.update(range(int(lo), int(hi or lo) + 1))
                local = [c for c in cores if c in nc]
                if len(local) >= local_world:
                    pool = local
            except Exception:
                pass
        per = max(1, len(pool) // max(1, local_world))
        mine = pool[(local_rank * per) % len(pool):][:per] or pool
        os.sched_setaffinity(0, mine)
        return dict(cores=mine, numa_node=node, cores_visible=len(cores))
    except Exception as e:  # not fatal: placement is an optimisation
        return dict(error=str(e))


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------------
def cpu_port_run(case, seconds=None, steps=None, warmup=0, threads=0, sample_images=None):
    """Times oracle.get_bboxes (C restatement of the reference path, OpenMP over images) on one batch of the
    workload (or its first `sample_images` images). Returns (images_per_s, ms_per_step list, cores, sample)."""
    import workloads
    from oracle import oracle
    if sample_images:
        case = dict(case, batch=min(case['batch'], sample_images))
    p = workloads.build_params(case)
    levels = workloads.host_levels(case, p)
    sf = workloads.scale_factors(case)
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
    big = case['sizes'][0][0] * case['strides'][0] > 700
    ips, ms, cores, sample = cpu_port_run(case, steps=max(1, args.steps), warmup=max(0, args.warmup),
                                          sample_images=16 if big else None)
    line = dict(impl='reference', metric=METRIC, value=ips, unit=UNIT, n_gpus=args.gpus, steps=len(ms),
                warmup=max(0, args.warmup), ms_per_step=statistics.mean(ms), p50_ms=statistics.median(ms),
                higher_is_better=True, scaling='strong' if 'global_batch' in case else 'weak', vs_baseline=None,
                dtype='f32', data='synthetic', config=config_dict(args, case),
                notes=dict(reference_arm='CPU port of the reference path (oracle/oracle.c, OpenMP over images, pinned to the '
                                         "reference's own source by golden vectors); the reference itself is Python + "
                                         'mmcv-full and cannot be installed here (mmcv-full absent)'),
                cpu_baseline=dict(value=ips, unit=UNIT, cores=cores, kind='port', sample=sample),
                e2e=dict(value=ips, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------------
def verify_against_oracle(case, p, levels, out):
    """All images of one batch of the timed region against the CPU oracle (bit-exact). Checker use of oracle/."""
    import numpy as np
    import workloads
    from oracle import oracle
    host = [x.cpu().numpy() for x in levels]
    if p.layout == 1:
        host = [np.ascontiguousarray(x) for x in host]  # logical NCHW values of the channels-last tensors
    orc = oracle.get_bboxes(p, host, workloads.scale_factors(case))
    res = {k: v.cpu().numpy() for k, v in out.items()}
    ok = int(res['status'][0]) == 0 and np.array_equal(res['count'], orc['count']) and \
        np.array_equal(res['num_candidates'], orc['num_candidates'])
    bad = []
    for b in range(p.batch):
        n = int(orc['count'][b])
        same = (res['count'][b] == n and np.array_equal(res['labels'][b, :n], orc['labels'][b])
                and np.array_equal(res['anchors'][b, :n], orc['anchors'][b])
                and np.array_equal(res['dets'][b, :n].view(np.uint32), orc['dets'][b].view(np.uint32)))
        if not same:
            bad.append(b)
    return bool(ok and not bad), bad


def run_yolopp(args):
    import torch
    import workloads
    import yolopp
    from yolopp import _capi as capi
    from yolopp.ops import Session, Pipeline, HostPipeline

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_world = int(os.environ.get('LOCAL_WORLD_SIZE', str(world)))
    if world != args.gpus and world > 1:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    placement = None
    if not args.no_affinity:
        try:
            bus = torch.cuda.get_device_properties(dev).pci_bus_id
            dom = getattr(torch.cuda.get_device_properties(dev), 'pci_domain_id', 0)
            devid = getattr(torch.cuda.get_device_properties(dev), 'pci_device_id', 0)
            pci = f'{dom:04x}:{bus:02x}:{devid:02x}.0'
        except Exception:
            pci = None
        placement = pin_rank_to_cores(local_rank, local_world, pci)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    case = get_case(args.workload)
    strong = 'global_batch' in case
    case['seed'] = case['seed'] + 1000 * rank  # every rank decodes its own shard of the global batch
    p = workloads.build_params(case)
    if strong:
        if case['global_batch'] % (world * case['batch']) != 0:
            raise SystemExit('global batch must be a multiple of gpus x images_per_call')
        calls_per_step = case['global_batch'] // world // case['batch']  # this rank's shard, `batch` images per call
    else:
        calls_per_step = 1
    # weak scaling: two input sets used alternately — consecutive steps never read the same 495 MB (each set alone
    # exceeds L2). strong scaling: the rank's whole shard is resident (distinct tensors per call).
    n_sets = calls_per_step if strong else 2
    inputs = [yolopp.synth.synth_levels(p, case['seed'] + 7 * j, case['dist'], device=dev) for j in range(n_sets)]
    if args.layout == 'nhwc':
        inputs = [[x.contiguous(memory_format=torch.channels_last) for x in lv] for lv in inputs]
        p.layout = capi.LAYOUT_NHWC
    levels = inputs[0]
    depth = max(1, args.pipeline_depth)
    pipe = Pipeline(p, depth, dev)
    sess = Session(p, dev)  # a batch that runs alone (batches_in_flight = 0): latency / per-kernel timing
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
    # ---- warm-up (also creates the plan handles of every (slot, input set) pair) ----
    W = max(3, args.warmup)
    for i in range(max(W, n_sets)):
        sess.run(inputs[i % n_sets], sf, profile=True)
    for i in range(max(W, 2 * depth, n_sets)):
        pipe.submit(inputs[i % n_sets], sf)
    # every slot must have seen every input set it will meet in the timed region
    for j in range(n_sets):
        for s_ in range(depth):
            pipe.sessions[s_]._plan(inputs[j], sf)
    torch.cuda.synchronize(dev)

    # ---- timed region: K steps, device-resident inputs, issued round-robin on `depth` streams (software
    #      pipelining across batches); CUDA events: start on the caller's stream before the first submit, end
    #      after the caller's stream has joined every pipeline stream ----
    K = args.steps
    calls = K * calls_per_step
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {n: 0.0 for n in capi.STAGE_NAMES}
    pipe.n = 0
    barrier()
    ev0.record()
    t_host0 = time.perf_counter()
    last = None
    for i in range(calls):
        last = (i % n_sets, pipe.submit(inputs[i % n_sets], sf, inputs_ready=True))  # (synthetic inputs: synchronised above)
    host_submit_ms = (time.perf_counter() - t_host0) * 1e3 / calls  # host time to issue one call (not a GPU time)
    pipe.join()
    ev1.record()
    torch.cuda.synchronize(dev)
    barrier()
    total_ms = ev0.elapsed_time(ev1)

    # ---- the outputs of the timed region's last batch against the oracle, all images ----
    verified, bad = None, []
    if not args.no_verify and rank == 0:
        big = case['sizes'][0][0] * case['strides'][0] > 700
        if not big or os.environ.get('YOLOPP_VERIFY_BIG'):
            verified, bad = verify_against_oracle(case, pipe.sessions[last[1]].p, inputs[last[0]], pipe.result(last[1]))
        else:
            # 1280^2: the oracle needs ~1 s per image; check the first 8 images of the last batch as a batch of 8
            sub = [x[:8].contiguous(memory_format=torch.channels_last) if args.layout == 'nhwc' else x[:8].contiguous()
                   for x in inputs[last[0]]]
            p8 = workloads.build_params(case, batch=8)
            p8.layout = p.layout
            out8 = {k: v[:8] if v.shape[0] == p.batch else v for k, v in pipe.result(last[1]).items()}
            verified, bad = verify_against_oracle(case, p8, sub, out8)

    # ---- latency of ONE batch (no overlap): calls strictly one after the other on one stream ----
    n_lat = min(max(calls, 10), 50)
    lat_ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_lat + 1)]
    lat_ev[0].record()
    for i in range(n_lat):
        sess.run(inputs[i % n_sets], sf, profile=False)
        lat_ev[i + 1].record()
    torch.cuda.synchronize(dev)
    latency_ms = lat_ev[0].elapsed_time(lat_ev[n_lat]) / n_lat
    lat_each = sorted(lat_ev[i].elapsed_time(lat_ev[i + 1]) for i in range(n_lat))  # per-batch latency distribution

    # ---- per-kernel durations (events between the kernels of each call, same stream), separate loop ----
    n_prof = min(max(calls, 5), 20)
    for i in range(n_prof):
        sess.run(inputs[i % n_sets], sf, profile=True)
        torch.cuda.synchronize(dev)
        for n, v in sess.stage_ms().items():
            stage_acc[n] += v
    stage_ms = {n: v / n_prof for n, v in stage_acc.items()}
    if info.tma_tiles == 0 and p.layout == capi.LAYOUT_NCHW:
        stage_ms['decode_tma'] = None   # no such launch in this configuration: the event gap is not a kernel time
    if info.ldg_blocks == 0 and info.dense_tiles == 0:
        stage_ms['decode_ldg'] = None
    # keep the GPU under the same load until the sampler has a few readings (it polls every 100 ms)
    t_end = time.perf_counter() + 0.6
    while time.perf_counter() < t_end:
        for i in range(20):
            pipe.submit(inputs[i % n_sets], sf)
        torch.cuda.synchronize(dev)
    clocks = sampler.stop()

    if distributed:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms_max = float(t.item())
    else:
        total_ms_max = total_ms
    images = case['batch'] * calls * world
    value = images / (total_ms_max * 1e-3)

    # ---- end to end through the public API with HOST buffers: pinned host tensors -> HostPipeline (H2D copy of the
    #      head tensors, the path, one pinned D2H of the detections), two slots so that copy(i+1) overlaps
    #      compute(i) + D2H(i) ----
    e2e = None
    if not args.no_e2e:
        host = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in inputs[0]]
        for h, x in zip(host, inputs[0]):
            h.copy_(x if x.is_contiguous() else x.contiguous())
        hp_params = workloads.build_params(case)
        hp = HostPipeline(hp_params, depth=2, device=dev)
        Ke = max(3, min(K, 10)) * calls_per_step
        for _ in range(2):
            res = hp.result(hp.submit(host))
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prev = None
        for _ in range(Ke):
            t = hp.submit(host)
            if prev is not None:
                res = hp.result(prev)
            prev = t
        res = hp.result(prev)
        e1.record()
        torch.cuda.synchronize(dev)
        wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_ms = max(e0.elapsed_time(e1), wall_ms)  # the host-side read of the results is inside the region
        my_gbs = hp.h2d_bytes * Ke / (e2e_ms * 1e-3) / 1e9
        rates = [my_gbs]
        if distributed:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
            g = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(g, torch.tensor([my_gbs], dtype=torch.float64, device=dev))
            rates = [float(x.item()) for x in g]
        e2e = dict(value=case['batch'] * world * Ke / (e2e_ms * 1e-3), unit=UNIT,
                   h2d_bytes_per_step=hp.h2d_bytes * calls_per_step, d2h_bytes_per_step=hp.d2h_bytes * calls_per_step,
                   steps=Ke // calls_per_step, ms_per_step=e2e_ms / Ke * calls_per_step,
                   h2d_gbs_per_rank=[round(r, 2) for r in rates],
                   api='yolopp.ops.HostPipeline (pinned host tensors -> H2D copy of the head tensors -> the path -> one pinned '
                       'D2H of the detections; 2 slots: copy(i+1) overlaps compute(i) + D2H(i)); bound by the PCIe H2D copy')
        assert len(res) == case['batch']

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (decode) ----
    peak, peak_src = load_peaks()
    if p.layout == capi.LAYOUT_NHWC:
        kern, alg_bytes = 'decode_rows_kernel', info.ldg_bytes_per_image * p.batch
    else:
        kern, alg_bytes = 'decode_tma_kernel', info.tma_bytes_per_image * p.batch
    dec_ms = stage_ms['decode_tma'] or 0.0
    achieved = alg_bytes / (dec_ms * 1e-3) / 1e9 if dec_ms > 0 else 0.0
    all_dec_bytes = (info.tma_bytes_per_image + info.ldg_bytes_per_image) * p.batch
    all_dec_ms = (stage_ms['decode_tma'] or 0.0) + (stage_ms['decode_ldg'] or 0.0)
    roofline = dict(bound='hbm', kernel=kern, achieved=achieved, peak=peak, unit='GB/s',
                    frac=achieved / peak,
                    traffic=load_traffic() if (args.workload == 'yolov4_608_b64_coco_sparse' and args.layout == 'nchw') else None,
                    peak_source=peak_src,
                    algorithmic_bytes_per_launch=alg_bytes, kernel_ms=dec_ms,
                    decode_all_levels=dict(bytes=all_dec_bytes, ms=all_dec_ms,
                                           achieved=all_dec_bytes / (all_dec_ms * 1e-3) / 1e9 if all_dec_ms > 0 else 0.0,
                                           frac=(all_dec_bytes / (all_dec_ms * 1e-3) / 1e9) / peak if all_dec_ms > 0 else 0.0),
                    stage_ms=stage_ms,
                    measured='CUDA events around each kernel on its stream, batches run one at a time (the pipelined '
                             'region overlaps kernels of different batches, so per-kernel durations are not defined there)')

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:  # (rank 0 at N = 1 only: at N > 1 the ranks share the host cores)
        base_case = get_case(args.workload)
        big = base_case['sizes'][0][0] * base_case['strides'][0] > 700
        ips, ms, cores, sample = cpu_port_run(base_case, seconds=args.cpu_baseline_seconds, sample_images=16 if big else None)
        cpu_baseline = dict(value=ips, unit=UNIT, cores=cores, kind='port', sample=sample,
                            ms_per_batch=statistics.mean(ms))

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W,
                ms_per_step=total_ms_max / K, host_submit_ms_per_step=host_submit_ms, latency_ms=latency_ms,
                p50_ms=statistics.median(lat_each), p90_ms=lat_each[int(0.9 * (n_lat - 1))],
                p99_ms=lat_each[int(0.99 * (n_lat - 1))], higher_is_better=True,
                scaling='strong' if strong else 'weak', vs_baseline=None, dtype='f32', data='synthetic',
                config=config_dict(args, case), verified=verified,
                notes=dict(pipeline_depth=depth, calls_per_step=calls_per_step,
                           pipeline=f'calls issued round-robin on {depth} CUDA streams, one workspace per stream, plan handles '
                                    '(yolopp_plan_run); latency_ms / p50 / p90 / p99 = one batch alone on one stream; '
                                    'host_submit_ms_per_step = host time to issue one call',
                           verified='all images of the last batch of the timed region == CPU oracle, bit-exact'
                                    + (f' — MISMATCH in images {bad[:8]}' if bad else ''),
                           placement=placement),
                clocks=clocks, e2e=e2e, gpu_launches=info.kernel_launches * calls, roofline=roofline,
                cpu_baseline=cpu_baseline)
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
