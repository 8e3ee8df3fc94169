"""The peer-memory gradient exchange of W emulated ranks (default 2, up to 8) on ONE GPU (all exchange buffers local,
barriers as signal-all-then-wait-all): runs every kernel of the exchange at the headline size so that ncu can capture
them, and prints CUDA-event times of the pieces.  Same calls as tests/test_peer_exchange_gpu.py.

    python scripts/peer_emulate.py [surfels] [iterations] [ranks]"""
import ctypes as C
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gs_lidar_b200 import GaussianRasterizer, synth, parallel
from gs_lidar_b200 import _lib as L

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
W = int(sys.argv[3]) if len(sys.argv) > 3 else 2
SPARSE = len(sys.argv) > 4 and sys.argv[4] == "sparse"  # the fused step's expansion kernel (zero-filled dL_dsh)
RANKS = tuple(range(W))
POSES = [(0.0, (0.0, 0.0, 0.0)), (0.4, (0.01, 0.0, 0.0)), (-0.3, (0.02, 0.0, 0.0)), (0.7, (0.03, 0.0, 0.0)),
         (-0.6, (0.04, 0.0, 0.0)), (0.2, (0.05, 0.0, 0.0)), (-0.1, (0.06, 0.0, 0.0)), (0.5, (0.07, 0.0, 0.0))][:W]
lib = L.load()
base = synth.make_scene(P, seed=0, view_yaw_deg=POSES[0][0], view_shift=POSES[0][1]).to("cuda")
scenes = [base]
for y, sh in POSES[1:]:
    cam = synth.make_scene(16, seed=0, view_yaw_deg=y, view_shift=sh)
    scenes.append(base._replace(viewmatrix=cam.viewmatrix.cuda(), projmatrix=cam.projmatrix.cuda(), campos=cam.campos.cuda()))
cot = {k: v.cuda() for k, v in synth.make_cotangents(scenes[0].H, scenes[0].W, 4, seed=1).items()}
nbytes = lib.gsl_peer_buffer_bytes(P, 4, W)
bufs = []
for _ in RANKS:
    q = C.c_void_p()
    L.check(lib.gsl_peer_alloc(nbytes, C.byref(q), None), "gsl_peer_alloc")
    bufs.append(q.value)


class Rank(parallel.PeerExchange):
    def __init__(self, r):
        super().__init__(sync=False)
        self.r = r

    def world_size(self):
        return W

    def rank(self):
        return self.r

    def prepare(self, P_, S, M, device):
        if self.pkey is None:
            self.setup(P_, S, device, buffers=bufs)
        return super().prepare(P_, S, M, device)


ranks = [Rank(r) for r in RANKS]
st = torch.cuda.current_stream()
sp = C.c_void_p(st.cuda_stream)
d_sh = [torch.zeros((P, 16, 4), device="cuda") for _ in RANKS[:2]]
leaves = []
for sc in scenes:
    lv = dict(means3D=sc.means3D.clone(), means2D=torch.zeros((P, 4), device="cuda"), opacities=sc.opacities.clone(),
              shs=sc.shs.clone(), features=sc.features.clone(), scales=sc.scales.clone(), rotations=sc.rotations.clone())
    for v in lv.values():
        v.requires_grad_(True)
    leaves.append(lv)
rasts = [GaussianRasterizer(synth.settings_for(sc)) for sc in scenes]


def barrier(slot):
    for r in RANKS:
        L.check(lib.gsl_peer_signal(C.byref(ranks[r].ctx), slot, sp), "signal")
    for r in RANKS:
        L.check(lib.gsl_peer_wait(C.byref(ranks[r].ctx), slot, sp), "wait")


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


times = {}
for it in range(ITERS):
    for r in RANKS:
        for v in leaves[r].values():
            v.grad = None
        with ranks[r]:
            contrib, color, feature, depth, alpha, radii = rasts[r](mask=scenes[r].mask, **leaves[r])
            torch.autograd.backward([color, feature, depth, alpha], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])
    for r in RANKS:
        ranks[r].ctx.epoch = ranks[r].epoch
    barrier(0)
    marks = [ev()]
    for r in RANKS:
        ranks[r].launch_reduce(P, 0, P, st)
    marks.append(ev())
    for r in RANKS:
        ranks[r].launch_expand(P, 3, 16, scenes[0].means3D, d_sh[r % 2], 0, P, st, sparse=SPARSE)
    marks.append(ev())
    barrier(2)
    out = [ranks[r].unpack(P) for r in RANKS]; out = None
    marks.append(ev())
    torch.cuda.synchronize()
    if it == ITERS - 1:
        times = dict(reduce_per_rank_ms=marks[0].elapsed_time(marks[1]) / W, expand_per_rank_ms=marks[1].elapsed_time(marks[2]) / W,
                     barrier_plus_unpack_per_rank_ms=marks[2].elapsed_time(marks[3]) / W)
assert all(int(ranks[r]._err[0]) == 0 for r in RANKS)
print(json.dumps(dict(P=P, world_emulated=W, sparse_expand=SPARSE, **times)))
