"""N>1 path on CPU: world_size-2 gloo run of the tile-sharding host logic bench.py uses. Each rank renders only
its own interleaved rows (here with the oracle standing in for the device) into a shared host image with no
data-path collective; the union must be bit-identical to the unsharded image, and a MAX all_reduce carries
the timing the way bench.py reports it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, shm, w, h):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tray_b200 import ray
    sc = O.rich_scene(2)
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    p = O.make_params(w, h, spp=2, max_depth=10, seed=2, stream_mode=1)
    img = np.load(shm, mmap_mode="r+")
    rows = ray.shard_rows(0, h, rank, world)
    for y in rows:
        O.render_lines(sc, cam, p, 0, y, y + 1, img)
    img.flush()
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(len(rows))], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    assert t.item() == world and n.item() == h
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tile_sharding(tmp_path, O):
    w, h, world = 40, 37, 2
    shm = str(tmp_path / "img.npy")
    np.lib.format.open_memmap(shm, mode="w+", dtype=np.uint8, shape=(h, w, 4)).flush()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, shm, w, h), nprocs=world, join=True)
    got = np.load(shm)
    full, _, _ = O.render(O.rich_scene(2), O.camera_init(w, h, **O.RICH_CAMERA),
                          O.make_params(w, h, spp=2, max_depth=10, seed=2, num_workers=1, stream_mode=1))
    assert np.array_equal(got, full)


def _split_worker(rank, world, port, out, w, h, spp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tray_b200 import multi
    sc = O.rich_scene(2)
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    p = O.make_params(w, h, spp=spp, max_depth=10, seed=2, stream_mode=1)
    off, stride, count = multi.sample_subset(spp, rank, world)
    sums, st = O.sample_sums(sc, cam, p, 0, h, off, stride, count)   # the oracle stands in for tray_render(sums_mode)
    t = torch.from_numpy(sums)
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)                       # the one exchange step of sample-split mode
    n = torch.tensor([float(st["paths"])], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    assert n.item() == w * h * spp
    if rank == 0:
        np.save(out, O.resolve_sums(t.numpy(), spp))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sample_split(tmp_path, O):
    """Sample-split host logic (tray_b200/multi.py): rank r takes samples s == r (mod world), sums are reduced to rank 0,
    which resolves. Same samples as the one-shot render, different summation order: within 1 LSB."""
    from tray_b200 import multi
    w, h, world, spp = 40, 23, 2, 7
    taken = sorted(off + j * st for r in range(world) for off, st, c in [multi.sample_subset(spp, r, world)] for j in range(c))
    assert taken == list(range(spp))
    out = str(tmp_path / "img.npy")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_split_worker, args=(world, port, out, w, h, spp), nprocs=world, join=True)
    got = np.load(out)
    full, _, _ = O.render(O.rich_scene(2), O.camera_init(w, h, **O.RICH_CAMERA),
                          O.make_params(w, h, spp=spp, max_depth=10, seed=2, num_workers=1, stream_mode=1))
    d = np.abs(got.astype(int) - full.astype(int)).max(axis=2)
    assert d.max() <= 1 and (d == 0).mean() > 0.99
