"""Where the frame-parallel step spends its exchange time: run under torchrun (one rank per GPU).
Times the headline step (1 frame per rank, fwd+bwd, 1M surfels) with parts of the gradient exchange ablated, and the two
NCCL collectives on their own.  Ablated variants compute WRONG gradients -- this is a measurement tool, not a mode."""
import json, os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gs_lidar_b200 import synth, parallel
import gs_lidar_b200.diff_gaussian_rasterization_2d as G

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
P, H, W, S = 1000000, 66, 1030, 4
STEPS = int(os.environ.get("STEPS", "30"))
scene = synth.make_scene(P, H=H, W=W, S=S, seed=0, **bench.make_frame_pose(0))
if rank != 0:
    cam = synth.make_scene(16, H=H, W=W, S=S, seed=0, **bench.make_frame_pose(rank))
    scene = scene._replace(viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos)
scene = scene.to(dev)
cot = {k: v.to(dev) for k, v in synth.make_cotangents(H, W, S, seed=1).items()}
rast = G.GaussianRasterizer(synth.settings_for(scene))
leaves = dict(means3D=scene.means3D.clone(), means2D=torch.zeros((P, 4), device=dev), opacities=scene.opacities.clone(),
              shs=scene.shs.clone(), features=scene.features.clone(), scales=scene.scales.clone(),
              rotations=scene.rotations.clone())
for v in leaves.values():
    v.requires_grad_(True)


def step():
    for v in leaves.values():
        v.grad = None
    contrib, color, feature, depth, alpha, radii = rast(mask=scene.mask, **leaves)
    torch.autograd.backward([color, feature, depth, alpha], [cot["color"], cot["feature"], cot["depth"], cot["alpha"]])


class Ablated(parallel.GradientExchange):
    def __init__(self, reduce=True, gather=True):
        super().__init__()
        self.do_reduce, self.do_gather = reduce, gather

    def world_size(self):
        return world

    def _all_reduce(self, flat):
        return super()._all_reduce(flat) if self.do_reduce else None

    def _all_gather(self, out, local):
        return super()._all_gather(out, local) if self.do_gather else None


def timed(fn, n):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3   # enqueue time per step (no sync inside)
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n, host], dtype=torch.float64, device=dev)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    return dict(ms_max=float(tmax[0]), ms_min=float(tmin[0]), host_ms_max=float(tmax[1]), host_ms_min=float(tmin[1]))


res = {"world": world, "lib": os.environ.get("GSL_B200_LIB", "default")}
PARTS_ONLY = os.environ.get("PARTS_ONLY", "0") == "1"
NCCL_VARIANTS = (("full", {}), ("no_reduce", dict(reduce=False)), ("no_gather", dict(gather=False)),
                 ("no_comm", dict(reduce=False, gather=False)), ("full_again", {}))
if os.environ.get("ONLY_PEER", "0") == "1":
    NCCL_VARIANTS = (("full", {}),)
if PARTS_ONLY:
    NCCL_VARIANTS = ()
else:
    res["no_exchange"] = timed(step, STEPS)
for name, kw in NCCL_VARIANTS:
    ex = Ablated(**kw).enable()
    res[name] = timed(step, STEPS)
    ex.disable()
# the peer-memory exchange (own kernels over NVLink, parallel.PeerExchange)
for chunks in ((8,) if PARTS_ONLY else (1, 2, 4, 8)):
    pex = parallel.PeerExchange(chunks=chunks).enable()
    res["peer_chunks%d" % chunks] = timed(step, STEPS)
    pex.disable()
    if chunks != 8:
        pex.close()
res["peer_chunks8_again"] = timed(lambda: (pex.enable(), step(), pex.disable()), STEPS)
if os.environ.get("PEER_PARTS", "1") == "1":
    st = torch.cuda.current_stream()
    d_sh = torch.empty((P, 16, 4), device=dev)

    def bar():
        pex.epoch += 1
        pex._barrier(0, pex.epoch, st)

    def red():
        pex.epoch += 1
        pex._barrier(1, pex.epoch * 64, st); pex.launch_reduce(P, 0, P, st); pex._barrier(2, pex.epoch, st)

    def exp():
        pex.epoch += 1
        pex._barrier(0, pex.epoch, st); pex.launch_expand(P, 3, 16, leaves["means3D"], d_sh, 0, P, st)

    def unp():
        pex.unpack(P)

    res["peer_barrier_alone"] = timed(bar, STEPS)
    res["peer_reduce_alone(2 barriers)"] = timed(red, STEPS)
    res["peer_expand_alone(1 barrier)"] = timed(exp, STEPS)
    res["peer_unpack_alone"] = timed(unp, STEPS)
    res["peer_error_flag"] = int(pex._err[0])
if os.environ.get("ONLY_PEER", "0") == "1":
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()
    sys.exit(0)
# the collectives alone
flat = torch.zeros(19 * P, device=dev)
loc = torch.zeros(4 * P + 4, device=dev)
gat = torch.zeros(world * (4 * P + 4), device=dev)
res["nccl_all_reduce_76MB"] = timed(lambda: dist.all_reduce(flat), STEPS)
res["nccl_all_gather_16MBxG"] = timed(lambda: dist.all_gather_into_tensor(gat, loc), STEPS)
half = flat[: 19 * P // 2]
res["nccl_all_reduce_38MB"] = timed(lambda: dist.all_reduce(half), STEPS)
q = flat[: 19 * P // 4]
res["nccl_all_reduce_19MB"] = timed(lambda: dist.all_reduce(q), STEPS)
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
