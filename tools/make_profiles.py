"""Turns the scratch output of one tools_gpu_round.sh run (gpurun_out/*_<tag>.*) into the tracked summaries under
profiles/ (read here, on the CPU box):   python tools/make_profiles.py <tag> [prefix]"""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pre = sys.argv[2] if len(sys.argv) > 2 else 'r1'
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')


def cp(src, dst):
    s = os.path.join(G, src)
    if os.path.isfile(s) and os.path.getsize(s) > 0:
        shutil.copyfile(s, os.path.join(P, dst))
        return True
    print('missing', src)
    return False


cp(f'bench_{tag}.json', f'{pre}_bench_sparse.json')
cp(f'bench_ref_{tag}.json', f'{pre}_bench_reference_arm.json')
cp(f'bench_dense_{tag}.json', f'{pre}_bench_dense.json')
cp(f'bench_depth1_{tag}.json', f'{pre}_bench_sparse_depth1.json')
cp(f'bench_yolov5_640_b128_sparse_{tag}.json', f'{pre}_bench_640_b128_csp.json')
cp(f'bench_yolov3_640_b128_sparse_{tag}.json', f'{pre}_bench_640_b128_v3.json')
cp(f'bench_yolov4_1280_b128_sparse_{tag}.json', f'{pre}_bench_1280_b128.json')
cp(f'bench_driver_{tag}.json', f'{pre}_bench_sparse_driver_cmdline.json')
cp(f'bench_nhwc_{tag}.json', f'{pre}_bench_sparse_nhwc.json')
cp(f'bench_1280_b1024_n1_{tag}.json', f'{pre}_bench_1280_b1024_n1.json')
cp(f'mish_{tag}.json', f'{pre}_mish.json')
cp(f'host_submit_{tag}.txt', f'{pre}_host_submit.txt')
cp(f'timeline_640v3_{tag}.txt', f'{pre}_timeline_640_b128_v3.txt')
cp(f'phases_608_{tag}.txt', f'{pre}_phases_608_b64.txt')
cp(f'smoke_{tag}.log', f'{pre}_smoke.log')
cp(f'launches_{tag}.csv', f'{pre}_launches.csv')
cp(f'pipe_traffic_{tag}.csv', f'{pre}_traffic_in_pipeline.csv')
cp(f'timeline_608_{tag}.txt', f'{pre}_timeline_608_b64.txt')
cp(f'timeline_640_{tag}.txt', f'{pre}_timeline_640_b128.txt')
cp(f'stock_gpu_{tag}.txt', f'{pre}_stock_gpu.txt')
cp(f'pytest_gpu_{tag}.log', f'{pre}_pytest_gpu.log')
for tool in ('memcheck', 'synccheck', 'racecheck'):
    cp(f'sanitizer_{tool}_{tag}.txt', f'{pre}_sanitizer_{tool}.txt')

# per-kernel ncu summaries + hot lines
for k, short in (('decode_tma', 'decode'), ('nms_image', 'nms'), ('select_kernel', 'select')):
    rep = os.path.join(G, f'prof_{k}_{tag}.ncu-rep')
    if not os.path.isfile(rep):
        print('missing', rep)
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_sum.py'), rep], capture_output=True, text=True).stdout
    open(os.path.join(P, f'{pre}_{short}_summary.txt'), 'w').write(out)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_src.py'), rep, k.split('_kernel')[0], '25'],
                         capture_output=True, text=True).stdout
    if out.strip():
        open(os.path.join(P, f'{pre}_{short}_hot_lines.txt'), 'w').write(out)
    if short == 'decode':
        vals = {}
        for ln in open(os.path.join(P, f'{pre}_{short}_summary.txt')):
            parts = ln.split()
            if len(parts) >= 2 and parts[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                v, unit = float(parts[1]), parts[2] if len(parts) > 2 else 'byte'
                vals[parts[0]] = int(v * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(unit, 1))
        if len(vals) == 2:
            rd, wr = vals['dram__bytes_read.sum'], vals['dram__bytes_write.sum']
            json.dump(dict(kernel='decode_tma_kernel<0, 32>',
                           source=f'profiles/{pre}_decode_summary.txt (ncu --set full --clock-control none, one launch, 608^2 batch 64 sparse)',
                           dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_launch=rd + wr,
                           algorithmic_bytes_per_launch=494887680,
                           note=f'same kernel inside the pipeline (single-pass metrics, caches uncontrolled): profiles/{pre}_traffic_in_pipeline.csv'),
                      open(os.path.join(P, 'decode_traffic.json'), 'w'), indent=1)

# launch shares
lp = os.path.join(G, f'launches_{tag}.csv')
if os.path.isfile(lp):
    rows = [r for r in csv.reader(open(lp)) if len(r) > 14 and r[0].isdigit()]
    acc = collections.defaultdict(list)
    for r in rows:
        if r[12] == 'gpu__time_duration.sum':
            v = float(r[14].replace(',', ''))
            v = v / 1e3 if r[13] in ('nsecond', 'ns') else v
            acc[r[4]].append(v)
    with open(os.path.join(P, f'{pre}_launch_shares.txt'), 'w') as f:
        f.write('ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3 (per-launch, serialised, cold caches)\n')
        for k in sorted(acc):
            f.write(f'{k[:70]:70s}   n={len(acc[k])} avg={sum(acc[k]) / len(acc[k]):.1f} us\n')
        own = {k: sum(v) / len(v) for k, v in acc.items() if any(s in k for s in ('select_kernel', 'decode_tma', 'nms_image'))}
        tot = sum(own.values())
        f.write('\nshare of one step (own kernels, one launch each per step):\n')
        for k, v in own.items():
            f.write(f'  {k[:68]:68s} {v:6.1f} us  {100 * v / tot:4.1f} %\n')
        bp = os.path.join(G, f'bench_{tag}.json')
        if os.path.isfile(bp):
            st = json.load(open(bp))['roofline']['stage_ms']
            t2 = st['select'] + st['decode_tma'] + st['nms_image']
            f.write('\nbench.py stage events (batches one at a time, warm): ' + ', '.join(
                f'{n} {st[n] * 1e3:.1f} us ({100 * st[n] / t2:.1f} %)' for n in ('select', 'decode_tma', 'nms_image')) + '\n')
print('done')
