#!/bin/bash
# Memory-safety and race evidence without compute-sanitizer (closed on this GPU pool): (1) the bounds-check build of the library
# (build/variants/libtraycuda_bounds.so, tools/build_variants.sh bounds="-DTRAY_BOUNDS_CHECK") over tools/sanitize_run.py must count
# zero refused accesses, and does count them when the table size is understated (negative control); (2) the race-prone layouts
# (regroup exchange area, wavefront queues, shared-memory pools) must give the same bits on every one of many repeated runs.
mkdir -p gpurun_out
export TRAY_LIB=$PWD/build/variants/libtraycuda_bounds.so
python - <<'PY' 2>&1 | tee gpurun_out/bounds_check.log
import hashlib, os, runpy, sys
sys.path.insert(0, os.getcwd())
from tray_b200 import _lib, rand, ray
assert "bounds" in _lib.library_path(), _lib.library_path()
runpy.run_path("tools/sanitize_run.py", run_name="__main__")
ctx = ray.default_context()
scene = ray.RichScene(rand.New(2))
def render(layout, accel, w=160, h=90, spp=8, depth=50):
    t = ray.New(w, h); t.Camera = ray.RichSceneCamera()
    t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Layout, t.Accel = depth, spp, 2, layout, accel
    img = t.Render(scene)
    return hashlib.sha1(img.tobytes() + ctx.read_hdr(w, h).tobytes()).hexdigest()[:12], t.Stats["bounds_violations"]
big = ray.RichScene(rand.New(2), 50)  # 10 001 spheres: the three-level walk
def render_big():
    t = ray.New(96, 54); t.Camera = ray.RichSceneCamera()
    t.MaxDepth, t.NumRaysPerPixel, t.Seed = 12, 4, 2
    img = t.Render(big)
    return hashlib.sha1(img.tobytes() + ctx.read_hdr(96, 54).tobytes()).hexdigest()[:12], t.Stats["bounds_violations"], t.Stats["box_tests"]
hb = {render_big() for _ in range(10)}
print("10 001 spheres, three-level walk (tables in global memory): 10 runs -> %d distinct image+HDR hashes, refused accesses = %d" % (len({x[0] for x in hb}), sum(int(x[1]) for x in hb)))
assert len(hb) == 1 and all(x[1] == 0 and x[2] > 0 for x in hb)
h0, v = render(ray.LAYOUT_PLAIN, ray.ACCEL_AUTO)
print("bounds-check build over tools/sanitize_run.py + a 160x90x8 render: refused accesses =", int(v))
assert v == 0
# (2) determinism stress of the race-prone layouts
for name, layout, accel in (("plain+clusters", ray.LAYOUT_PLAIN, ray.ACCEL_AUTO), ("regroup+clusters", ray.LAYOUT_REGROUP, ray.ACCEL_AUTO),
                            ("regroup+linear", ray.LAYOUT_REGROUP, ray.ACCEL_BRUTE), ("wavefront", ray.LAYOUT_WAVEFRONT, ray.ACCEL_BRUTE),
                            ("plain+bvh", ray.LAYOUT_PLAIN, ray.ACCEL_BVH)):
    hs = {render(layout, accel)[0] for _ in range(25)}
    print("%-18s 25 runs -> %d distinct image+HDR hashes %s" % (name, len(hs), sorted(hs)))
    assert hs == {h0}, name
# (1b) negative control: understate the staged table size by 4 KB -> the checks must fire (and the image goes wrong)
ctx.configure(99, 1)
h1, v1 = render(ray.LAYOUT_PLAIN, ray.ACCEL_CLUSTER)
ctx.configure(99, 0)
print("negative control (table size understated by 4 KB): refused accesses =", int(v1), "image changed:", h1 != h0)
assert v1 > 0
print("bounds check ok")
PY
