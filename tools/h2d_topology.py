"""Host->device copy bandwidth per GPU when 1, 2, 4, 8 GPUs copy at the same time (one process, one stream and one
pinned buffer per device): which GPUs share a PCIe uplink, and what the host side can feed in total. This bounds the
end-to-end (`e2e`) numbers of bench.py, whose step is dominated by the 495 MB H2D copy of the head tensors.
    python tools/h2d_topology.py [out.json]"""
import json, sys, time
import torch

n = torch.cuda.device_count()
SZ = 512 << 20
host = [torch.empty(SZ, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(SZ, dtype=torch.uint8, device=f'cuda:{i}') for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]


def run(ids, reps=6):
    for i in ids:
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in ids:
            with torch.cuda.stream(streams[i]):
                dev[i].copy_(host[i], non_blocking=True)
    for i in ids:
        streams[i].synchronize()
    dt = time.perf_counter() - t0
    return SZ * reps / dt / 1e9  # GB/s per GPU (all finish together)


run(list(range(n)), 2)
rows = []
subsets = [[0]] + [[0, j] for j in range(1, n)] + [list(range(k)) for k in (4, 8) if k <= n]
if n >= 8:
    subsets += [[0, 2, 4, 6], [0, 1, 4, 5]]
for ids in subsets:
    g = run(ids)
    rows.append(dict(gpus=ids, gbs_per_gpu=round(g, 1), gbs_total=round(g * len(ids), 1)))
    print(rows[-1], flush=True)
if len(sys.argv) > 1:
    json.dump(dict(device_count=n, bytes_per_copy=SZ, rows=rows), open(sys.argv[1], 'w'), indent=1)
