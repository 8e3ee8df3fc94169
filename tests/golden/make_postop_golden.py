"""Generates tests/golden/postop_*.npz: outputs of the reference's OWN panorama post-ops, `pano_to_lidar` and
`depth_to_normal` (/root/reference/utils/graphics_utils.py:96-149, pure PyTorch, imported unmodified), for pinning
gs_lidar_b200/range_map.py.  Runs only in the build container (needs /root/reference).

    python tests/golden/make_postop_golden.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_graphics_utils", "/root/reference/utils/graphics_utils.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = dict(postop_kitti_half=(34, 257, (-24.9, 2.0), (-90.0, 90.0), 21), postop_kitti_360=(33, 258, (-24.9, 2.0), (-180.0, 180.0), 22),
                 postop_opv2v=(32, 128, (-25.0, 2.0), (-180.0, 180.0), 23))
    for name, (H, W, vfov, hfov, seed) in cases.items():
        g = torch.Generator().manual_seed(seed)
        rng = 2.0 + 6.0 * torch.rand(1, H, W, generator=g)
        rng = rng + 0.5 * torch.sin(torch.arange(W) / 9.0)[None, None, :]          # some structure for the normals
        rng = rng * (torch.rand(1, H, W, generator=g) > 0.2)                       # ray drops: zero range
        pts = ref.pano_to_lidar(rng, vfov, hfov)
        nrm = ref.depth_to_normal(rng, vfov, hfov)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), range_image=rng.numpy(), vfov=np.array(vfov), hfov=np.array(hfov),
                            points=pts.numpy(), normals=nrm.numpy())
        print(name, tuple(rng.shape), "points", tuple(pts.shape))


if __name__ == "__main__":
    main()
