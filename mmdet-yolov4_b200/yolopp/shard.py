"""Multi-GPU: images are independent, so a batch is sharded over the ranks of one node (one process per GPU) with
NO collective on the data path. These helpers are the whole "parallel runtime" the path needs:
the contiguous shard of a rank, and the gather of the per-image detections to rank 0 in image order (the
eval-harness convenience the reference implements with pickles in mmdet/apis/test.py:160-190)."""
import torch
import torch.distributed as dist


def shard_range(num_images, world_size, rank):
    """Contiguous image range [lo, hi) of `rank`: the first (num_images % world_size) ranks take one extra."""
    base, rem = divmod(int(num_images), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_detections(local_results, num_images, group=None):
    """local_results: list over this rank's images of (dets ndarray/tensor (n,5), labels (n,)). Returns on rank 0
    the list over ALL images in image order, None elsewhere. Fixed-size exchange: counts first, then one padded
    block per rank (works with gloo on CPU tensors and nccl on CUDA tensors)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(num_images, world, rank)
    assert len(local_results) == hi - lo
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    counts = torch.tensor([int(len(l)) for _, l in local_results], dtype=torch.int64, device=dev)
    max_imgs = (num_images + world - 1) // world
    cnt_pad = torch.zeros(max_imgs, dtype=torch.int64, device=dev)
    cnt_pad[:counts.numel()] = counts
    all_cnt = [torch.zeros_like(cnt_pad) for _ in range(world)]
    dist.all_gather(all_cnt, cnt_pad, group=group)
    cap = max(int(torch.stack(all_cnt).max()), 1)
    block = torch.zeros((max_imgs, cap, 6), dtype=torch.float32, device=dev)
    for i, (d, l) in enumerate(local_results):
        n = len(l)
        if n:
            block[i, :n, :5] = torch.as_tensor(d, dtype=torch.float32, device=dev)
            block[i, :n, 5] = torch.as_tensor(l, device=dev).to(torch.float32)  # class ids < 2^24: exact
    blocks = [torch.zeros_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    if rank != 0:
        return None
    out = []
    for r in range(world):
        rlo, rhi = shard_range(num_images, world, r)
        for i in range(rhi - rlo):
            n = int(all_cnt[r][i])
            out.append((blocks[r][i, :n, :5].cpu(), blocks[r][i, :n, 5].to(torch.int64).cpu()))
    return out
