"""GOP sharding across the GPUs of one box.

Closed GOPs are independent, so a clip is partitioned into contiguous GOP ranges, one per rank;
each rank encodes its range on its own GPU and the host concatenates the byte streams in rank
order.  There is no data-path collective: the only exchange is a gather of the (small)
bitstreams to rank 0.  `first_gop` carries the clip-level GOP index so idr_pic_id keeps
alternating across shard boundaries and the concatenation equals the unsharded stream.
"""
from __future__ import annotations


def gop_ranges(nframes: int, gop: int, world: int):
    """Contiguous, balanced GOP ranges -> [(first_frame, n_frames, first_gop)] per rank."""
    ngop = (nframes + gop - 1) // gop
    out = []
    base, rem = divmod(ngop, world)
    g0 = 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        f0 = min(g0 * gop, nframes)
        f1 = min((g0 + cnt) * gop, nframes)
        out.append((f0, f1 - f0, g0))
        g0 += cnt
    return out


def gather_streams(local: bytes, rank: int, world: int, group=None):
    """Gather per-rank byte strings on rank 0 (torch.distributed, any backend) and concatenate."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    m = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros(m, dtype=torch.uint8, device=dev)
    if local:
        buf[: len(local)] = torch.frombuffer(bytearray(local), dtype=torch.uint8).to(dev)
    parts = [torch.zeros(m, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    if rank != 0:
        return None
    return b"".join(bytes(parts[r][: int(sizes[r].item())].cpu().numpy()) for r in range(world))
