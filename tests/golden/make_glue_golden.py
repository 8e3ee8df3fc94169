"""Generates tests/golden/glue_getters.npz: outputs of the reference's OWN GaussianModel accessors
(scene/gaussian_model.py:139-186, unmodified, imported through oracle/ref_python.py) and of the per-surfel part of its
render() (gaussian_renderer/__init__.py:64-115) on CPU tensors, for pinning tests/glue_oracle.py (the PyTorch restatement the
GPU glue tests check the fused kernels against).  Runs only in the build container (needs /root/reference).

    python tests/golden/make_glue_golden.py
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import glue_oracle as GO
    from oracle import ref_python
    ref_python.stage()
    dummy = types.ModuleType("no_rasterizer")
    dummy.GaussianRasterizationSettings = dummy.GaussianRasterizer = object
    ns = ref_python.load(dummy)
    args = SimpleNamespace(sh_degree=3, time_duration=[-0.5, 0.5], no_time_split=True, t_grad=True, contract=False, t_init=0.1,
                           big_point_threshold=0.1, cycle=0.2, velocity_decay=1.0, random_init_point=0)
    P = 1031
    raw = GO.make_model(P, seed=11)
    pc = ns.GaussianModel(args)
    for n in GO.RAW + ("_features_dc", "_features_rest"):
        setattr(pc, n, getattr(raw, n).detach().clone())
    out = {n: getattr(raw, n).detach().numpy() for n in GO.RAW + ("_features_dc", "_features_rest")}
    ts, shift = 0.11, 0.03
    out.update(timestamp=np.float32(ts), time_shift=np.float32(shift), T=np.float32(pc.T), velocity_decay=np.float32(pc.velocity_decay),
               get_xyz_SHM=pc.get_xyz_SHM(ts).numpy(), get_xyz_SHM_shifted=pc.get_xyz_SHM(ts - shift).numpy(),
               get_inst_velocity=pc.get_inst_velocity.numpy(), get_marginal_t=pc.get_marginal_t(ts).numpy(),
               get_marginal_t_shifted=pc.get_marginal_t(ts - shift).numpy(), get_opacity=pc.get_opacity.numpy(),
               get_scaling=pc.get_scaling.numpy(), get_scaling_t=pc.get_scaling_t.numpy(), get_rotation=pc.get_rotation.numpy(),
               get_features=pc.get_features.numpy())
    np.savez_compressed(os.path.join(HERE, "glue_getters.npz"), **out)
    print("wrote glue_getters.npz:", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.shape})


if __name__ == "__main__":
    main()
