"""Sample-subset renders through the C ABI (tray_params.sums_mode, tray_device_sums, tray_resolve_sums): the multi-process
sample split (NCCL variant) and progressive refinement. The raw sums must equal the oracle's bit for bit; consecutive
slices must reproduce the one-shot image exactly; strided subsets summed in a different order stay within 1 LSB."""
import os
import sys

import numpy as np
import pytest

from tray_b200 import _lib, multi, rand, ray
from test_gpu_parity import oracle_cam, oracle_flat, tracer

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(ctx, w, h, spp, depth, precision=ray.FP64_STRICT):
    scene = ray.RichScene(rand.New(2))
    t = tracer(w, h, spp, depth, precision=precision)
    t.Context = ctx
    scene = t._prepare(scene)
    ctx.upload(scene.flatten())
    return scene, t


@pytest.mark.parametrize("precision,fma", [(ray.FP64_STRICT, 0), (ray.FP64_FMA, 1)])
def test_subset_sums_bit_exact_vs_oracle(ctx, O, precision, fma):
    w, h, spp, depth = 97, 41, 9, 50
    scene, t = _setup(ctx, w, h, spp, depth, precision)
    osc, ocam = oracle_flat(O, scene), oracle_cam(O, t)
    op = O.make_params(w, h, spp=spp, max_depth=depth, seed=2, stream_mode=1, fma_mode=fma)
    for off, stride, count in [(0, 1, 9), (2, 3, 3), (8, 1, 1), (1, 2, 4)]:
        p = t._params(0, h)
        p.sample_offset, p.sample_stride, p.sample_count, p.sums_mode = off, stride, count, ray.SUMS_OVERWRITE
        st = ctx.render(t.to_c(), p, None)
        want, ost = O.sample_sums(osc, ocam, op, 0, h, off, stride, count)
        assert np.array_equal(ctx.read_hdr(w, h), want)
        assert st["paths"] == w * h * count and st["segments"] == ost["segments"]
        img = np.zeros((h, w, 4), dtype=np.uint8)
        ctx.resolve_sums(count, img)
        assert np.array_equal(img, O.resolve_sums(want, count))


def test_progressive_slices_reproduce_the_one_shot_render(ctx, O):
    w, h, spp, depth = 120, 67, 16, 50
    scene = ray.RichScene(rand.New(2))
    one = tracer(w, h, spp, depth).Render(scene).copy()
    t = tracer(w, h, spp, depth)
    osc = ocam = None
    seen = []
    sums = None
    for done, img in t.RenderProgressive(scene, 5):
        seen.append(done)
        if osc is None:
            osc, ocam = oracle_flat(O, scene), oracle_cam(O, t)
        op = O.make_params(w, h, spp=spp, max_depth=depth, seed=2, stream_mode=1)
        k0 = seen[-2] if len(seen) > 1 else 0
        sums, _ = O.sample_sums(osc, ocam, op, 0, h, k0, 1, done - k0, sums=sums)
        assert np.array_equal(img, O.resolve_sums(sums, done))   # every intermediate frame: mean over the rays so far
    assert seen == [5, 10, 15, 16]
    assert np.array_equal(t.imageData, one)                        # the last one IS Render's image
    assert t.Stats["paths"] == w * h * spp


def test_rows_and_shards_with_sums(ctx, O):
    """Sums cover exactly the rows this context rendered (row range + external shards), row-major in local order."""
    w, h, spp, depth = 64, 40, 4, 12
    scene, t = _setup(ctx, w, h, spp, depth)
    osc, ocam = oracle_flat(O, scene), oracle_cam(O, t)
    op = O.make_params(w, h, spp=spp, max_depth=depth, seed=2, stream_mode=1)
    p = t._params(8, 33)
    p.shard_index, p.shard_count = 1, 2
    p.sample_offset, p.sample_stride, p.sample_count, p.sums_mode = 1, 1, 3, ray.SUMS_OVERWRITE
    ctx.render(t.to_c(), p, None)
    ptr, n = ctx.device_sums()
    rows = ray.shard_rows(8, 33, 1, 2)
    assert ptr and n == len(rows) * w * 3
    hdr = ctx.read_hdr(w, h)
    for y in rows:
        want, _ = O.sample_sums(osc, ocam, op, y, y + 1, 1, 1, 3)
        assert np.array_equal(hdr[y], want[0])


def test_sums_error_paths(ctx):
    w, h = 32, 16
    scene, t = _setup(ctx, w, h, 4, 5)
    p = t._params(0, h)
    p.sample_offset, p.sample_stride, p.sample_count, p.sums_mode = 2, 1, 3, ray.SUMS_OVERWRITE   # 2+2 >= 4
    with pytest.raises(ray.TrayError):
        ctx.render(t.to_c(), p, None)
    p.sample_count, p.sums_mode = 1, 3
    with pytest.raises(ray.TrayError):
        ctx.render(t.to_c(), p, None)
    ctx.render(t.to_c(), t._params(0, h), None)           # a classic render leaves no sums behind
    with pytest.raises(ray.TrayError):
        ctx.device_sums()
    with pytest.raises(ray.TrayError):
        ctx.resolve_sums(4)
    p = t._params(0, h)
    p.sample_offset, p.sample_stride, p.sample_count, p.sums_mode = 0, 1, 1, ray.SUMS_ACCUMULATE  # nothing to continue
    with pytest.raises(ray.TrayError):
        ctx.render(t.to_c(), p, None)


def test_sample_split_host_path_on_one_gpu(ctx, O):
    """tray_b200.multi.render_sample_split with the exchange step injected: 'rank 1' is rendered by a second context on
    the same GPU and added into rank 0's device buffer the way the NCCL reduce would."""
    import torch
    w, h, spp, depth = 96, 54, 8, 50
    scene = ray.RichScene(rand.New(2))
    one = tracer(w, h, spp, depth).Render(scene).copy()
    other = ray.Context([0])
    t1 = tracer(w, h, spp, depth)
    t1.Context = other
    got1 = {}

    def grab(sums):
        got1["sums"] = sums.clone()
    assert multi.render_sample_split(t1, scene, 1, 2, reduce_fn=grab) is None
    t0 = tracer(w, h, spp, depth)
    img = multi.render_sample_split(t0, scene, 0, 2, reduce_fn=lambda sums: sums.add_(got1["sums"]))
    d = np.abs(img.astype(int) - one.astype(int)).max(axis=2)
    assert d.max() <= 1 and (d == 0).mean() > 0.999
    assert t0.Stats["paths"] + t1.Stats["paths"] == w * h * spp
    other.close()


def _nccl_worker(rank, world, port, out, w, h, spp, depth):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tray_b200 import multi as M, rand as R, ray as Y
    t = Y.New(w, h)
    t.Camera = Y.RichSceneCamera()
    t.MaxDepth, t.NumRaysPerPixel, t.Seed = depth, spp, 2
    t.Context = Y.Context([rank])
    img = M.render_sample_split(t, Y.RichScene(R.New(2)), rank, world)
    if rank == 0:
        np.save(out, img)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_sample_split_nccl_reduce(ctx, tmp_path, world):
    """One process per GPU, partial sums reduced to rank 0 with NCCL over NVLink (the north_star's sample-split variant)."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    w, h, spp, depth = 320, 181, 16, 50
    one = tracer(w, h, spp, depth).Render(ray.RichScene(rand.New(2))).copy()
    out = str(tmp_path / "img.npy")
    mp.spawn(_nccl_worker, args=(world, 29800 + os.getpid() % 1000, out, w, h, spp, depth), nprocs=world, join=True)
    img = np.load(out)
    d = np.abs(img.astype(int) - one.astype(int)).max(axis=2)
    assert d.max() <= 1 and (d == 0).mean() > 0.999
