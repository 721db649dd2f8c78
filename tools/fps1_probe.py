"""Throughput-policy FPS (fps_cull1_kernel) alone at 8 x 40k -> 2048 (SAD_B200_LIB=.../libsad_fpsprof.so built with
-DSAD_FPS_PROFILE prints the per-phase cycles of CTA 0)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa
from sad_b200 import ops
from sad_b200.scenes import make_scenes

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pol = sys.argv[2] if len(sys.argv) > 2 else "throughput"
x = torch.from_numpy(make_scenes(B, 40000, "surface")[0]).cuda()
g = ops.build_scene_grid(x)
ref = ops.furthest_point_sample(x, 2048, g)
for variant in (0, -2):      # 0 = the policy's kernel, -2 = the 16-warp / three-register-set instance
    ts = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); got = ops.furthest_point_sample(x, 2048, g, pol, False, variant); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"B={B} {pol} variant {variant}: {min(ts):.3f} ms  ({1e3 * min(ts) / 2047:.3f} us per pick)  equal to the cluster kernel: {bool((got == ref).all())}", flush=True)
