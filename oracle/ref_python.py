"""oracle/ref_python.py -- TEST INFRASTRUCTURE ONLY.

The reference's own PYTHON callers of the rasterizer, unmodified, as checkers for the drop-in claim
(one import line of gaussian_renderer/__init__.py:10):

    gaussian_renderer/__init__.py   render() :16-155, render_range_map() :158-227
    scene/gaussian_model.py         GaussianModel and its accessors :139-186
    scene/cameras.py                Camera
    utils/{graphics_utils,general_utils,sh_utils}.py

`stage()` (run by __graft_entry__.build() / oracle.build_ref() where /root/reference exists) copies these files --
byte for byte -- into oracle/_ref/py/, which is git-ignored and travels to the GPU box like the compiled reference
kernels; nothing of them is committed.  `load(rasterizer)` imports them with
  * `gaussian_renderer.diff_gaussian_rasterization_2d` pre-bound to the module given (this package's drop-in, or the
    adapter over the reference's compiled CUDA kernels in tests/ref_rasterizer.py) -- the reference wrapper itself
    JIT-builds its extension at import time and cannot be imported here,
  * an EMPTY `scene/__init__.py` (the reference's imports the data loaders: open3d, camtools, ...),
  * stubs for the packages this image lacks and the path never calls: plyfile, simple_knn, matplotlib.cm, open3d, and
    kornia.utils.create_meshgrid (two lines of torch; Camera.__init__ builds a pixel grid with it).
"""
import hashlib
import importlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE = os.path.join(HERE, "_ref", "py")
FILES = ["gaussian_renderer/__init__.py", "scene/gaussian_model.py", "scene/cameras.py", "utils/graphics_utils.py",
         "utils/general_utils.py", "utils/sh_utils.py"]
_MODULES = ("gaussian_renderer", "scene", "scene.gaussian_model", "scene.cameras", "utils", "utils.graphics_utils",
            "utils.general_utils", "utils.sh_utils", "gaussian_renderer.diff_gaussian_rasterization_2d")


def reference_dir():
    return os.environ.get("GSL_REFERENCE_DIR", "/root/reference")


def stage():
    """Copies the reference files into oracle/_ref/py (only where the reference is present).  Returns the directory or None."""
    ref = reference_dir()
    if not os.path.isdir(os.path.join(ref, "gaussian_renderer")):
        return STAGE if available() else None
    for rel in FILES:
        dst = os.path.join(STAGE, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref, rel), dst)
    # package markers written by this repo (not reference sources): empty on purpose
    for pkg in ("scene", "utils"):
        with open(os.path.join(STAGE, pkg, "__init__.py"), "w") as f:
            f.write("")
    with open(os.path.join(STAGE, "MANIFEST"), "w") as f:
        for rel in FILES:
            with open(os.path.join(STAGE, rel), "rb") as g:
                f.write("%s  %s\n" % (hashlib.sha1(g.read()).hexdigest(), rel))
    return STAGE


def available():
    return all(os.path.exists(os.path.join(STAGE, rel)) for rel in FILES)


def _stub(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__dict__["__gsl_stub__"] = True
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _install_stubs():
    import torch

    def _missing(name):
        try:
            importlib.import_module(name)
            return False
        except Exception:
            return True

    if _missing("plyfile"):
        _stub("plyfile", PlyData=object, PlyElement=object)
    if _missing("simple_knn"):
        pkg = _stub("simple_knn")
        pkg._C = _stub("simple_knn._C", distCUDA2=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("simple_knn stub")))
    if _missing("matplotlib"):
        pkg = _stub("matplotlib")
        pkg.cm = _stub("matplotlib.cm")
    if _missing("open3d"):
        _stub("open3d")
    if _missing("kornia"):
        def create_meshgrid(height, width, normalized_coordinates=True, device=None, dtype=None):
            assert not normalized_coordinates
            xs = torch.arange(width, device=device, dtype=dtype or torch.float32)
            ys = torch.arange(height, device=device, dtype=dtype or torch.float32)
            return torch.stack(torch.meshgrid(xs, ys, indexing="xy"), dim=-1).unsqueeze(0)  # (1, H, W, 2) = (x, y)

        pkg = _stub("kornia")
        pkg.utils = _stub("kornia.utils", create_meshgrid=create_meshgrid)


def load(rasterizer_module):
    """Imports the staged reference modules with `rasterizer_module` standing where gaussian_renderer/__init__.py:10
    imports `.diff_gaussian_rasterization_2d` from.  Returns a namespace (render, render_range_map, GaussianModel, Camera,
    graphics_utils).  Every call re-imports, so two rasterizers can be compared in one process."""
    if not available():
        raise ImportError("the reference's Python files are not staged (oracle/_ref/py); run __graft_entry__.build() "
                          "where /root/reference exists")
    _install_stubs()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in _MODULES or k.startswith(("scene.", "utils.", "gaussian_renderer."))}
    sys.path.insert(0, STAGE)
    try:
        sys.modules["gaussian_renderer.diff_gaussian_rasterization_2d"] = rasterizer_module
        gr = importlib.import_module("gaussian_renderer")
        ns = types.SimpleNamespace(
            render=gr.render, render_range_map=gr.render_range_map, module=gr,
            GaussianModel=importlib.import_module("scene.gaussian_model").GaussianModel,
            Camera=importlib.import_module("scene.cameras").Camera,
            graphics_utils=importlib.import_module("utils.graphics_utils"),
            rasterizer=rasterizer_module)
    finally:
        sys.path.remove(STAGE)
        for k in list(sys.modules):
            if k in _MODULES or k.startswith(("scene.", "utils.", "gaussian_renderer.")):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return ns
