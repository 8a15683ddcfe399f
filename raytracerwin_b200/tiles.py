"""Screen-space tile ownership for the multi-GPU path (host-side index arithmetic only).

The frame is cut into tile_size x tile_size tiles, numbered row-major; rank r of n owns the tiles with
id % n == r (the geometry sits in the middle of the frame, so interleaving balances it
where contiguous bands would not).  The DENSE order of a rank's pixels — the layout of the buffers that
rt_gpu_pack_owned writes and rt_gpu_unpack_owned reads — is: owned tiles by increasing id, row-major
inside a tile, image-clipped.  The same arithmetic lives in csrc/rt_exchange.cu (tile_copy); a GPU test checks
that the two agree.
"""
import numpy as np


def owned_tiles(width, height, tile_size, tile_count, tile_rank):
    """[(x0, y0, w, h)] of the tiles rank `tile_rank` owns, in dense order."""
    if tile_count <= 1 or tile_size <= 0:
        return [(0, 0, width, height)]
    tiles_x = (width + tile_size - 1) // tile_size
    tiles_y = (height + tile_size - 1) // tile_size
    out = []
    for t in range(tile_rank, tiles_x * tiles_y, tile_count):
        tx, ty = t % tiles_x, t // tiles_x
        x0, y0 = tx * tile_size, ty * tile_size
        out.append((x0, y0, min(tile_size, width - x0), min(tile_size, height - y0)))
    return out


def dense_index(width, height, tile_size, tile_count, tile_rank):
    """Frame pixel index (y*width+x) of every owned pixel, in dense order (int64 array)."""
    parts = []
    for x0, y0, w, h in owned_tiles(width, height, tile_size, tile_count, tile_rank):
        ys, xs = np.mgrid[y0:y0 + h, x0:x0 + w]
        parts.append((ys * width + xs).reshape(-1))
    return np.concatenate(parts) if parts else np.zeros(0, np.int64)


def owned_count(width, height, tile_size, tile_count, tile_rank):
    return sum(w * h for _, _, w, h in owned_tiles(width, height, tile_size, tile_count, tile_rank))


def gather_owned(dist, send, counts, rank, world, dst=0):
    """The one exchange step of the path: every rank contributes its dense buffer `send` (a torch tensor of
    max(counts) rows, its first counts[rank] rows valid) and `dst` receives the list of all of them.
    Works with any torch.distributed backend (NCCL over NVLink on the GPUs, gloo in the CPU tests)."""
    import torch
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    return recv
