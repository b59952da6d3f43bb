"""One config-2 frame in the wavefront layout (for an ncu launch list of its kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_b200 import rand, ray  # noqa: E402

t = ray.New(1920, 1080)
t.Camera = ray.RichSceneCamera()
t.MaxDepth, t.NumRaysPerPixel, t.Seed = 50, 64, 2
t.Layout = {"wavefront": ray.LAYOUT_WAVEFRONT, "regroup": ray.LAYOUT_REGROUP, "plain": ray.LAYOUT_PLAIN}[sys.argv[1] if len(sys.argv) > 1 else "wavefront"]
t.Render(ray.RichScene(rand.New(2)))
print(t.Stats)
