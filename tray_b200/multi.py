"""One process per GPU (torchrun): the two ways the path partitions across B200s.

  render_tiles         interleaved rows, rank r renders rows y == r (mod world); no data-path collective,
                       every rank writes its own rows of the caller's (shared) host image.
  render_sample_split  rank r renders samples s == r (mod world) of every pixel into fp64 colour sums that stay on the
                       device; ONE exchange step: a sum-reduce of that buffer to rank 0 over NCCL (NVLink / NVSwitch);
                       rank 0 applies 1/N + sRGB. The summation order differs from the reference's sample order, so
                       the linear values agree to a few ulp and the 8-bit image within 1 LSB.

torch is plumbing here (process group + the NCCL call on the library's device buffer); the arithmetic is libtraycuda's.
The reference's analogue of both is the worker fan-out of Tracer.Render (ray/tracer.go:85-116)."""
import numpy as np

from . import _lib


def sample_subset(spp, rank, world):
    """(offset, stride, count) of the samples rank `rank` of `world` traces in sample-split mode: s == rank (mod world)."""
    count = (spp - rank + world - 1) // world if spp > rank else 0
    return rank, world, count


class _DeviceBuffer:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def render_tiles(tracer, scene, rank, world, out=None):
    """Tile mode across processes. `out`: (h, w, 4) uint8 host image shared by the ranks (rows of other ranks untouched)."""
    tracer.ShardIndex, tracer.ShardCount = (rank, world) if world > 1 else (0, 0)
    if out is not None:
        tracer.imageData = out
    return tracer.Render(scene)


def render_sample_split(tracer, scene, rank, world, group=None, reduce_fn=None):
    """Sample-split mode across processes; returns the image on rank 0 and None elsewhere.
    reduce_fn(tensor) overrides the default torch.distributed.reduce(SUM, dst=0) (tests inject their own)."""
    import torch
    import torch.distributed as dist
    from . import ray
    scene = tracer._prepare(scene)
    ctx = tracer.Context or ray.default_context()
    ctx.upload(scene.flatten())
    off, stride, count = sample_subset(tracer.NumRaysPerPixel, rank, world)
    if tracer.NumRaysPerPixel < world:
        raise _lib.TrayError(_lib.E_INVALID, "sample split needs NumRaysPerPixel >= number of ranks")
    p = tracer._params(0, tracer.height)
    p.shard_index, p.shard_count = 0, 0
    p.sample_offset, p.sample_stride, p.sample_count, p.sums_mode = off, stride, count, _lib.SUMS_OVERWRITE
    tracer.Stats = ctx.render(tracer.to_c(), p, None)
    ptr, n = ctx.device_sums()
    sums = torch.as_tensor(_DeviceBuffer(ptr, n), device=torch.device("cuda", ctx.devices[0]))
    if reduce_fn is not None:
        reduce_fn(sums)
    elif world > 1:
        dist.reduce(sums, dst=0, op=dist.ReduceOp.SUM, group=group)  # the one exchange step (NCCL over NVLink)
    torch.cuda.synchronize(sums.device)
    if rank != 0:
        return None
    ctx.resolve_sums(tracer.NumRaysPerPixel, tracer.imageData)
    return tracer.imageData
