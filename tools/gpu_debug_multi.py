import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import shocovox_b200 as S
from shocovox_b200 import scenes
tree = scenes.build_tree(scenes.cpu_render_scene(), S.Octree)
cam = scenes.cpu_render_camera()
vp = S.Viewport(cam.origin, cam.direction, cam.frustum, cam.fov)
for devs in ([0], [0, 0], [0, 0, 0]):
    try:
        m = S.MultiGPU(tree, devs, vp, (320, 203))
        f = m.read_root_frame()
        print("multi", devs, "ok hits", int((f["hit_id"] != S.MISS).sum()), flush=True)
        g = m.render_to_host()
        print("  host-assembled equal", all(np.array_equal(f[k].view(np.uint32), g[k].view(np.uint32)) for k in ("hit_id", "albedo", "distance")), flush=True)
        del m
    except S.OctreeError as e:
        print("multi", devs, "FAILED", e, flush=True)
