"""Generates tests/golden/toy_*.npz: outputs of the reference's OWN pure-PyTorch panorama splatting
(`surface_splatting`, /root/reference/scripts/compare_2dgs_3dgs.py, unmodified) on CPU, for pinning oracle/toy_splat.py.

Runs only in the build container (needs /root/reference).  The script hard-codes CUDA placement and imports
matplotlib / the training repo's utils, so it is imported through import-time shims -- no source is edited or copied:
  * sys.modules stubs for matplotlib, matplotlib.pyplot and utils.general_utils (seed_everything only),
  * torch.Tensor.cuda -> identity, Tensor.to('cuda') -> identity, device='cuda' keyword of torch.zeros dropped.

    python tests/golden/make_toy_golden.py
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference_script():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    utils = types.ModuleType("utils")
    gu = types.ModuleType("utils.general_utils")
    gu.seed_everything = lambda seed: torch.manual_seed(seed)
    utils.general_utils = gu
    sys.modules["utils"], sys.modules["utils.general_utils"] = utils, gu
    torch.Tensor.cuda = lambda self, *a, **k: self
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
        if k.get("device") is not None and str(k["device"]).startswith("cuda"):
            k.pop("device")
        return _to(self, *a, **k) if (a or k) else self

    torch.Tensor.to = to
    _zeros = torch.zeros

    def zeros(*a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k.pop("device")
        return _zeros(*a, **k)

    torch.zeros = zeros
    sys.path.insert(0, os.path.join(REF, "scripts"))
    import compare_2dgs_3dgs as ref
    return ref


def main():
    ref = import_reference_script()
    from oracle import toy_splat
    torch.set_num_threads(os.cpu_count())
    cases = dict(toy_wide_64=(64, 250, 60, 11), toy_pano_48=(48, 258, 34, 12), toy_dense_300=(300, 128, 32, 13))
    for name, (P, W, H, seed) in cases.items():
        means, scales, quats, colors, opac, intrins, viewmat = toy_splat.make_inputs(P, W, H, seed=seed)
        if name == "toy_wide_64":  # a rotated, translated camera like the script's own get_cameras()
            c2w = torch.tensor([[-8.6086e-01, 3.7950e-01, -3.3896e-01, 0.3], [5.0884e-01, 6.4205e-01, -5.7346e-01, 0.2],
                                [1.0934e-08, -6.6614e-01, -7.4583e-01, -0.4], [0.0, 0.0, 0.0, 1.0]])
            viewmat = torch.linalg.inv(c2w).permute(1, 0).contiguous()
            means = (torch.cat([means, torch.ones(P, 1)], 1) @ c2w.T)[:, :3].contiguous()  # keep the surfels in view
        image, depth, centre, radii, _ = ref.surface_splatting(means, scales, quats, colors, opac, intrins, viewmat, None)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), means3D=means.numpy(), scales=scales.numpy(), quats=quats.numpy(),
                            colors=colors.numpy(), opacities=opac.numpy(), intrins=intrins.numpy(), viewmat=viewmat.numpy(),
                            image=image.numpy().astype(np.float32), depth=depth.numpy().astype(np.float32),
                            centre=centre.numpy(), radii=radii.numpy())
        print(name, "P", P, "WxH", W, H, "covered pixels", int((image.sum(-1) > 0).sum()), "image max", float(image.max()))


if __name__ == "__main__":
    main()
