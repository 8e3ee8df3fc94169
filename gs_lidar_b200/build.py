"""Builds the C-ABI library libgsl_b200.so (plain nvcc, sm_100a only, no torch headers).

    python -m gs_lidar_b200.build        # or gs_lidar_b200.build.build()

The .so is written in-tree (gs_lidar_b200/libgsl_b200.so) so that it travels with a repo snapshot;
it is git-ignored.  A stamp of the sources + flags skips rebuilds when nothing changed.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libgsl_b200.so")
SOURCES = ["gsl_api.cu", "gsl_preprocess.cu", "gsl_binning.cu", "gsl_sort.cu", "gsl_glue.cu", "gsl_peer.cu", "gsl_chamfer.cu", "gsl_postops.cu", "gsl_render_fwd.cu", "gsl_render_bwd.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-I", INCLUDE,
         "-Wno-deprecated-gpu-targets"]


def _stamp():
    h = hashlib.sha1()
    h.update(" ".join(FLAGS).encode())
    names = sorted(os.listdir(CSRC)) + ["../../include/gsl_b200.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: developer builds (e.g. variant="stats", defines=["-DGSL_STATS"]) written to
    libgsl_b200_<variant>.so and selected at run time with GSL_B200_LIB; the product is the default build."""
    lib = LIB if not variant else os.path.join(HERE, "libgsl_b200_%s.so" % variant)
    stamp_file = os.path.join(HERE, ".build_stamp" + ("_" + variant if variant else ""))
    stamp = _stamp() + "|" + " ".join(defines)
    if not force and os.path.exists(lib) and os.path.exists(stamp_file):
        with open(stamp_file) as f:
            if f.read().strip() == stamp:
                return lib
    objdir = os.path.join(HERE, "build" + ("_" + variant if variant else ""))
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", lib] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return lib


if __name__ == "__main__":
    variant, defines = None, []
    for a in sys.argv[1:]:
        if a.startswith("--variant="):
            variant = a.split("=", 1)[1]
        elif a.startswith("-D"):
            defines.append(a)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=variant, defines=defines))
