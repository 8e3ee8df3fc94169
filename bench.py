#!/usr/bin/env python
"""bench.py -- panoramas/s, forward+backward, 1M synthetic surfels on a 66x1030 KITTI-360-shaped
panorama (BASELINE.json metric; workload = configs[2], the config the metric is quoted on).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One process per GPU.  For N > 1 launch with torchrun (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* from the
env): every rank renders its own LiDAR frame (camera pose) of the replicated surfel set, and the
per-surfel gradients are summed inside the timed step by this repo's peer-memory exchange (own kernels
storing over NVLink, gs_lidar_b200.parallel.PeerExchange; `--exchange nccl` selects the NCCL baseline)
-- "weak" scaling: per-GPU work is fixed, value = frames of all ranks / max-over-ranks time.  Before
anything is timed at N > 1 the exchanged gradients are checked against a torch.distributed all-reduce
of the dense per-rank gradients (`exchange_parity`); a failed check prints no `value`.

A "step" is one forward+backward pass of the rasterizer op for one frame through the public
GaussianRasterizer API.  Rank 0 prints ONE JSON line (see the module-level keys at the end).
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: NCCL's version / debug banner goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"  # the VERSION banner is printf'ed to stdout

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "panoramas/sec fwd+bwd at 1M surfels 66x1030"
UNIT = "panoramas/s"
METRIC_C4 = "panoramas/sec fwd+bwd, dynamic training step through render() at 1M surfels 66x1030 (C4)"
WORKLOAD = "KITTI-360 seq 1908-shaped static training step: 1M synthetic surfels, fwd+bwd 66x1030 (hfov +-180, SH deg 3, S=4)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--surfels", type=int, default=1000000)
    ap.add_argument("--height", type=int, default=66)
    ap.add_argument("--width", type=int, default=1030)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--toy-only", action="store_true", help="only time the pure-PyTorch toy splat on the host cores (no GPU needed)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient exchange by this repo's own kernels over NVLink peer memory (default) or by "
                         "NCCL collectives (all-reduce + all-gather)")
    ap.add_argument("--exchange-chunks", type=int, default=1)
    ap.add_argument("--exchange-schedule", default="default", choices=["default", "early-low", "early-high", "late"],
                    help="N > 1, peer exchange: when the SH factors leave and at which priority the SH expansion runs "
                         "(gs_lidar_b200.parallel.PeerExchange.set_schedule); default = the library's")
    ap.add_argument("--wrap-azimuth", action="store_true",
                    help="opt-in extension, NOT the reference's semantics and not the headline: periodic panorama")
    ap.add_argument("--cpu-sample-surfels", type=int, default=0, help="0 = pick from a quick calibration")
    ap.add_argument("--clock-sample-ms", type=int, default=20, help="period of the clock sampler on rank 0 (0 = off)")
    ap.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi", "off"],
                    help="nvml: a side process polling two NVML queries (default); smi: an `nvidia-smi --query-gpu -lms` loop")
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c4", "c5"],
                    help="c3 (default, the headline): fwd+bwd, 1M surfels, 66x1030.  c4: the DYNAMIC training step (BASELINE.json "
                         "configs[3]): 51 frames with their own poses and timestamps through gs_lidar_b200.renderer.render() "
                         "(fused glue: SHM motion model, marginal, activations) + the rasterizer, frame-parallel with the fused "
                         "exchange carrying the glue's frame-dependent VJP.  c2: forward-only render of 100k surfels, 66x1030, "
                         "one GPU.  c5: stress inference, 4M surfels, 128x2048 OPV2V-style panoramas, --frames frames sharded by "
                         "frame over the GPUs (BASELINE.json configs[1] / configs[4]); both forward only")
    ap.add_argument("--frames", type=int, default=512, help="c5: frames of the batch (all ranks together)")
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="CUDA-graph replay of the forward / backward pass (gs_lidar_b200.set_cuda_graphs); the line always "
                         "carries the other mode's device-timed number as well")
    return ap.parse_args()


def init_dist(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


_NVML_SAMPLER = r"""
import sys, time
import pynvml as N
idx, period = int(sys.argv[1]), float(sys.argv[2]) / 1e3
N.nvmlInit()
h = N.nvmlDeviceGetHandleByIndex(idx)
mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
while True:
    sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
    try:
        r = N.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        r = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    sys.stdout.write("%d,%d,%d\n" % (sm, mx, r))
    sys.stdout.flush()
    time.sleep(period)
"""


class ClockSampler:
    """SM clock / clock-event reasons sampled DURING the timed region, on rank 0, by a separate process: two NVML queries per
    sample (`nvidia-smi --query-gpu=... -lms 20` was measured to cost 2-4 % of the step it samples, and at N > 1 every rank
    waits for rank 0 at the exchange barriers; the NVML loop was measured at no visible cost).  `--clock-sampler smi` keeps
    the nvidia-smi loop."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # NVML clock-event reason bits (nvml.h)
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, period_ms=20, kind="nvml"):
        self.index = index
        self.period_ms = int(period_ms)
        self.kind = kind
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        if self.period_ms <= 0 or self.kind == "off":
            return
        try:
            if self.kind == "nvml":
                self.p = subprocess.Popen([sys.executable, "-c", _NVML_SAMPLER, str(self.index), str(self.period_ms)],
                                          stdout=self.f, stderr=subprocess.DEVNULL)
            else:
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                           "--format=csv,noheader,nounits", "-lms", str(self.period_ms)], stdout=self.f,
                                          stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], sampler=self.kind, period_ms=self.period_ms)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            try:
                if self.kind == "nvml":
                    if len(parts) < 3:
                        continue
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                    bits = int(parts[2])
                    for nme, b in self.BITS.items():
                        if bits & b:
                            reasons.add(nme)
                else:
                    if len(parts) < 9:
                        continue
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                    for nme, v in zip(names, parts[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(nme)
            except ValueError:
                continue
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = max(mx)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def algorithmic_bytes(P, V, R, N, S, K, M):
    """algorithmic bytes of one fwd+bwd panorama."""
    fwd = 45 * P + V * (16 * K + 4 * S) + 4 * P + N * (8 + 16 + 4 * (S + 3) + 16 + 4)
    bwd = N * (16 + 4 * (S + 3) + 16 + 4) + V * (44 + 16 * K + 4 * S) + P * (12 + 16 + 16 + 4 * S + 4 + 24 + 16 * M + 12 + 16)
    binning = 40 * R
    return fwd, bwd, binning


def render_bwd_algorithmic_bytes(V, N, S):
    """Dominant kernel (k_render_bwd), DESIGN.md 'kernels': cotangents + forward state per pixel,
    one 64-B record + 16-B colour read and one packed gradient record written per visible surfel."""
    per_px = 4 * (4 + (S + 3) + 4 + 1) + 12 + 8
    per_surfel = 64 + 16 + 4 * (20 + ((S + 3) // 4) * 4)
    return N * per_px + V * per_surfel


def shared_config(P, H, W, S, M, D, V, R, world):
    """The `config` object of the JSON line: identical keys and values in both arms (ours / reference) for the same run."""
    return {"workload": WORKLOAD, "surfels": P, "height": H, "width": W, "sh_degree": D, "sh_coefficients": M,
            "feature_channels": S, "visible_surfels": V, "tile_instances": R,
            "parallelism": "frame-parallel dp%d, 1 frame/rank/step" % world,
            "l2": "inputs (%.0f MB of surfel parameters per step) exceed the 126 MB L2; no explicit flush" % (
                (45 * P + 16 * M * P + 4 * S * P) / 1e6)}


def make_frame_pose(rank):
    # frame k of a drive: yaw N(0, 0.5 deg)-like fixed offsets and +0.1*sf*k along x
    yaw = [0.0, 0.4, -0.3, 0.7, -0.6, 0.2, -0.1, 0.5][rank % 8]
    return dict(view_yaw_deg=yaw, view_shift=(0.01 * rank, 0.0, 0.0))


def check_exchange_parity(step, leaves, names, exchange, dev, tol=1e-4):
    """N > 1, outside the timed region: one step with the fused exchange against the same step without it followed by a
    torch.distributed all-reduce (sum) of every dense gradient tensor."""
    step()
    torch.cuda.synchronize(dev)
    got = {k: leaves[k].grad.detach().clone() for k in names}
    exchange.disable()
    step()
    torch.cuda.synchronize(dev)
    want = {k: leaves[k].grad.detach().clone() for k in names}
    exchange.enable()
    max_err, worst, identical = 0.0, None, True
    for k in names:
        dist.all_reduce(want[k], op=dist.ReduceOp.SUM)
        err = float((got[k].double() - want[k].double()).norm() / (want[k].double().norm() + 1e-30))
        if not (err <= max_err):  # also catches NaN
            max_err, worst = err, k
        ref0 = got[k].clone()
        dist.broadcast(ref0, src=0)
        identical = identical and bool(torch.equal(ref0.view(torch.int32), got[k].view(torch.int32)))
    flags = torch.tensor([max_err if max_err == max_err else float("inf"), 0.0 if identical else 1.0],
                         dtype=torch.float64, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    max_err, identical = float(flags[0]), float(flags[1]) == 0.0
    return {"max_err": max_err, "worst_tensor": worst, "bit_identical": identical, "tol": tol,
            "against": "torch.distributed all-reduce (sum) of the dense per-rank gradients, norm-wise per tensor",
            "ok": bool(max_err <= tol and identical)}


def ncu_traffic_from_profiles():
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of this repo's kernels from the newest
    profiles/rNN_*ncu_full*.md (written by scripts/ncu_summary.py from an `ncu --set full` capture of this workload)."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full*.md")))
    out, src = {}, None
    for f in files:  # later rounds overwrite earlier ones
        try:
            lines = open(f).read().splitlines()
        except OSError:
            continue
        hdr = None
        for ln in lines:
            cells = [c.strip() for c in ln.strip().strip("|").split("|")]
            if "kernel" in cells and "dram_rd" in cells:
                hdr = cells
                continue
            if hdr and len(cells) == len(hdr) and cells[0].startswith("k_"):
                row = dict(zip(hdr, cells))
                try:
                    mb = float(row["dram_rd"].split()[0]) + float(row["dram_wr"].split()[0])
                except (ValueError, IndexError):
                    continue
                out[re.sub(r"<.*", "", row["kernel"])] = mb * 1e6
                src = os.path.basename(f)
    return out, src


def run_ours(args, rank, world, local):
    from gs_lidar_b200 import synth, parallel
    from gs_lidar_b200 import _lib as L
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    dev = torch.device("cuda", local)
    P, H, W, S = args.surfels, args.height, args.width, 4
    if args.wrap_azimuth:
        G.set_wrap_azimuth(True)
    # replicated surfels (the world-space set of frame 0 on every rank), one camera pose per rank
    scene = synth.make_scene(P, H=H, W=W, S=S, seed=0, **make_frame_pose(0))
    if rank != 0:
        cam = synth.make_scene(16, H=H, W=W, S=S, seed=0, **make_frame_pose(rank))
        scene = scene._replace(viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos)
    scene = scene.to(dev)
    cot_cpu = synth.make_cotangents(H, W, S, seed=1)
    cot = {k: v.to(dev) for k, v in cot_cpu.items()}
    settings = synth.settings_for(scene)
    rast = G.GaussianRasterizer(settings)
    names = ["means3D", "means2D", "opacities", "shs", "features", "scales", "rotations"]
    leaves = dict(means3D=scene.means3D.clone(), means2D=torch.zeros((P, 4), device=dev), opacities=scene.opacities.clone(),
                  shs=scene.shs.clone(), features=scene.features.clone(), scales=scene.scales.clone(),
                  rotations=scene.rotations.clone())
    for v in leaves.values():
        v.requires_grad_(True)
    # N > 1: the gradient exchange is fused into the backward pass (parallel.PeerExchange: packed rows + SH factors pushed
    # over NVLink peer memory by this repo's kernels; parallel.GradientExchange is the NCCL baseline)
    exchange = None
    if world > 1 and args.exchange_schedule != "default":
        parallel.PeerExchange.set_schedule(early_factors=args.exchange_schedule != "late",
                                           expand_low_priority=args.exchange_schedule == "early-low")
    if world > 1:
        exchange = (parallel.PeerExchange(chunks=args.exchange_chunks) if args.exchange == "peer"
                    else parallel.GradientExchange()).enable()
    last = {}

    def step(rasterizer=rast, cots=cot):
        for v in leaves.values():
            v.grad = None
        contrib, color, feature, depth, alpha, radii = rasterizer(
            means3D=leaves["means3D"], means2D=leaves["means2D"], opacities=leaves["opacities"], shs=leaves["shs"],
            features=leaves["features"], scales=leaves["scales"], rotations=leaves["rotations"], mask=scene.mask)
        torch.autograd.backward([color, feature, depth, alpha],
                                [cots["color"], cots["feature"], cots["depth"], cots["alpha"]])
        last.update(color=color, feature=feature, depth=depth, alpha=alpha, radii=radii, contrib=contrib)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up (the clock sampler starts here: the timed region alone is tens of milliseconds, shorter than
    # nvidia-smi's sampling period, so the record covers warm-up + timed region, all under the same load)
    sampler = ClockSampler(local, args.clock_sample_ms, args.clock_sampler)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    # a box that has just been handed over is cold (first touches of the driver, allocator growth, clocks still ramping:
    # measured up to 20 % on the first dozen steps): keep stepping for about a second before anything is timed
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 1.0:
        for _ in range(10):
            step()
        torch.cuda.synchronize(dev)
    barrier()
    # realised V and R (for the roofline arithmetic)
    V = int((last["radii"] > 0).sum())
    R = int(last["color"].grad_fn.num_rendered) if last["color"].grad_fn is not None else 0

    # ---- N > 1: parity gate, outside the timed region.  The gradients the fused exchange returned must equal a
    # torch.distributed all-reduce (sum) of the dense gradients each rank computes alone, norm-wise to 1e-4 per tensor, and
    # must be bit-identical on all ranks.  A timing without this check is void, so a failure prints no value.
    exchange_parity = None
    if world > 1:
        exchange_parity = check_exchange_parity(step, leaves, names, exchange, dev)
        if not exchange_parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "impl": "ours", "n_gpus": world, "error": "exchange parity check failed",
                                  "exchange_parity": exchange_parity}))
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(1)
        for _ in range(2):
            step()
        barrier()

    # ---- N > 1 diagnostic, outside the timed region: every rank alone (exchange off, no barrier between the ranks inside
    # the loop) -- the spread of these times is the rank skew the in-step barriers turn into waiting
    per_rank_alone = None
    if world > 1:
        exchange.disable()
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(args.steps):
            step()
        eb.record()
        torch.cuda.synchronize(dev)
        mine = torch.tensor([ea.elapsed_time(eb) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank_alone = [float(x) for x in allr]
        exchange.enable()
        for _ in range(2):
            step()
        barrier()

    # ---- timed region: device-resident inputs, per-kernel profiling OFF ---------------------------
    L.load().gsl_profile_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(graph_on):
        G.set_cuda_graphs(graph_on, deferred_count=True)
        for _ in range(4):  # a new call signature runs once un-graphed, is captured on its second call
            step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    use_graph = args.graph == "on"
    ms_other = timed(not use_graph)
    ms = timed(use_graph)  # the headline mode runs last and stays on for the e2e leg
    clocks = sampler.stop() if rank == 0 else {}
    G.set_cuda_graphs(False)  # the per-kernel pass below needs individual launches

    # ---- separate pass: per-kernel CUDA-event times (two event records per kernel, so NOT part of the timed region)
    # and the per-step spread (one event per step boundary)
    L.load().gsl_profile_read(None, None, 1)
    L.load().gsl_profile_enable(1)
    prof_steps = min(args.steps, 20)
    for _ in range(prof_steps):
        step()
    barrier()
    L.load().gsl_profile_enable(0)
    kms = (C.c_double * L.GSL_K_COUNT)()
    kn = (C.c_int64 * L.GSL_K_COUNT)()
    L.load().gsl_profile_read(kms, kn, 1)
    spread = None
    if args.steps >= 50:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            step()
            evs[i + 1].record()
        barrier()
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps))
        q = lambda f: per[min(len(per) - 1, int(f * len(per)))]
        spread = {"p10_ms": q(0.10), "median_ms": q(0.50), "p90_ms": q(0.90), "steps": args.steps,
                  "note": "per-step device time between step-boundary events on rank 0, separate pass"}

    G.set_cuda_graphs(use_graph, deferred_count=True)
    # ---- e2e: per-step host->device copy of the frame's inputs, device->host read of the result --
    # Per-step inputs of the op are the camera (view matrix, camera centre) and the cotangent maps the
    # loss produced from the host-side ground-truth panoramas; the surfel parameters are resident model
    # state, exactly as scene/gaussian_model.py keeps them on the GPU in the reference's training loop.
    # One pinned staging buffer per direction (a loader hands over one block per frame; 3 copies per step instead of 11):
    #   in : [view matrix 16 | camera centre 3 | pad] on the main stream, [cotangent planes color 4, feature S+3, depth 4, alpha 1]
    #        on a copy stream;   out: the rendered planes in the same order + a gradient checksum
    n_pl = 4 + (S + 3) + 4 + 1
    h_cam = torch.zeros(20).pin_memory()
    h_cam[:16] = scene.viewmatrix.cpu().reshape(-1)
    h_cam[16:19] = scene.campos.cpu()
    d_camv = torch.empty(20, device=dev)
    d_cam = dict(viewmatrix=d_camv[:16].view(4, 4), campos=d_camv[16:19])
    h_cotv = torch.cat([cot_cpu[k].reshape(-1, H, W) for k in ("color", "feature", "depth", "alpha")]).pin_memory()
    d_cotv = torch.empty_like(h_cotv, device=dev)
    d_cot = dict(color=d_cotv[0:4], feature=d_cotv[4:4 + S + 3], depth=d_cotv[4 + S + 3:8 + S + 3], alpha=d_cotv[8 + S + 3:])
    h_out = dict(maps=torch.empty((n_pl, H, W)).pin_memory(), gradsum=torch.empty((3,)).pin_memory())
    h2d = (h_cam.numel() + h_cotv.numel()) * 4
    d2h = sum(v.numel() * 4 for v in h_out.values())

    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_in, ev_fwd = torch.cuda.Event(), torch.cuda.Event()
    ev_bwd, ev_maps = torch.cuda.Event(), torch.cuda.Event()
    h_outs = [h_out, {k: torch.empty_like(v).pin_memory() for k, v in h_out.items()}]
    ev_d2h = [torch.cuda.Event(), torch.cuda.Event()]
    gsum = torch.empty((3,), device=dev)
    consumed = []

    def e2e_step(k, pipelined):
        """One step with its host traffic inside: camera + cotangent maps from pinned host memory, all rendered maps and a
        gradient checksum back to pinned host memory.  The cotangent upload rides a copy stream under the forward pass (the
        backward waits for it), the maps are read back on a second copy stream under the backward pass.
        pipelined=False: the step ends with a full synchronisation of all three streams (nothing overlaps across steps).
        pipelined=True : what a training loop with a prefetching loader and asynchronous logging does -- the host never
        blocks on the step it has just issued; it consumes the results of step k-2 (waits for THEIR device->host event, long
        complete) before issuing step k, and the two pinned result buffers alternate."""
        main = torch.cuda.current_stream(dev)
        ho = h_outs[k % 2] if pipelined else h_out
        if pipelined and k >= 2:
            # the host reads the results of step k-2 (they sit in the pinned buffers step k is about to reuse) before it issues
            # step k: it is never more than two steps ahead and never waits for the step it has just issued
            ev_d2h[k % 2].synchronize()
            consumed.append(float(h_outs[k % 2]["gradsum"][0]))
        d_camv.copy_(h_cam, non_blocking=True)
        if pipelined:
            s_in.wait_event(ev_bwd)    # the previous backward has read the cotangent buffers
            main.wait_event(ev_maps)   # the previous maps have left the (static, graph-mode) output buffers
        else:
            s_in.wait_stream(main)
        with torch.cuda.stream(s_in):
            d_cotv.copy_(h_cotv, non_blocking=True)
            ev_in.record(s_in)
        st = settings._replace(viewmatrix=d_cam["viewmatrix"], campos=d_cam["campos"])
        for v in leaves.values():
            v.grad = None
        contrib, color, feature, depth, alpha, radii = G.GaussianRasterizer(st)(
            means3D=leaves["means3D"], means2D=leaves["means2D"], opacities=leaves["opacities"], shs=leaves["shs"],
            features=leaves["features"], scales=leaves["scales"], rotations=leaves["rotations"], mask=scene.mask)
        ev_fwd.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_fwd)
            # colour, feature, depth and alpha planes are consecutive planes of ONE allocation of the op: one copy
            planes = torch.as_strided(color.detach(), (n_pl, H, W), (H * W, W, 1), color.storage_offset())
            ho["maps"].copy_(planes, non_blocking=True)
            ev_maps.record(s_out)
        main.wait_event(ev_in)
        torch.autograd.backward([color, feature, depth, alpha],
                                [d_cot["color"], d_cot["feature"], d_cot["depth"], d_cot["alpha"]])
        torch.sum(leaves["means3D"].grad, dim=0, out=gsum)
        ev_bwd.record(main)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_bwd)
            ho["gradsum"].copy_(gsum, non_blocking=True)
            ev_d2h[k % 2].record(s_out)
        if not pipelined:
            s_out.synchronize()
            main.synchronize()

    def e2e_run(pipelined, n):
        ev_bwd.record(torch.cuda.current_stream(dev))
        ev_maps.record(torch.cuda.current_stream(dev))
        for k in range(3):
            e2e_step(k, pipelined)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for k in range(n):
            e2e_step(k, pipelined)
        if pipelined:
            for k in (n - 2, n - 1):
                if k >= 0:
                    ev_d2h[k % 2].synchronize()
                    consumed.append(float(h_outs[k % 2]["gradsum"][0]))
        torch.cuda.current_stream(dev).wait_stream(s_out)
        e1.record()
        barrier()
        ms_ = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
        tt = torch.tensor([ms_], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_steps = max(5, args.steps // 2)
    e2e_sync_ms = e2e_run(False, e2e_steps)
    e2e_ms = e2e_run(True, e2e_steps)

    # ---- M2: the reference-faithful pair of half panoramas, two calls of H x W/2 with hfov +-90 and the
    # front / back view matrices of scene/kitti360_loader.py:215-218, gradients summed; reported next to the headline
    m2 = None
    if world == 1 and W % 2 == 0 and not args.wrap_azimuth:
        flip = torch.diag(torch.tensor([-1.0, 1.0, -1.0, 1.0], device=dev))
        half = dict(image_width=W // 2, hfov=(-90.0, 90.0))
        st_front = settings._replace(**half)
        vm_back = (flip @ scene.viewmatrix.t()).t().contiguous()  # world->camera of the back camera, transposed
        st_back = settings._replace(viewmatrix=vm_back, projmatrix=vm_back, **half)
        r_front, r_back = G.GaussianRasterizer(st_front), G.GaussianRasterizer(st_back)
        cot_h = {k: v[..., :W // 2].contiguous() for k, v in cot.items()}

        def m2_step():
            for v in leaves.values():
                v.grad = None
            for r in (r_front, r_back):
                o = r(means3D=leaves["means3D"], means2D=leaves["means2D"], opacities=leaves["opacities"], shs=leaves["shs"],
                      features=leaves["features"], scales=leaves["scales"], rotations=leaves["rotations"], mask=scene.mask)
                torch.autograd.backward([o[1], o[2], o[3], o[4]], [cot_h["color"], cot_h["feature"], cot_h["depth"], cot_h["alpha"]])

        for _ in range(3):
            m2_step()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(max(5, args.steps // 2)):
            m2_step()
        e1.record()
        torch.cuda.synchronize(dev)
        m2_ms = e0.elapsed_time(e1) / max(5, args.steps // 2)
        m2 = {"value": 1e3 / m2_ms, "unit": UNIT, "ms_per_panorama": m2_ms,
              "what": "M2: two half panoramas %dx%d (hfov +-90, front + back camera), fwd+bwd each, gradients accumulated" % (H, W // 2)}

    if rank != 0:
        return None
    N = H * W
    K = (scene.sh_degree + 1) ** 2
    M = scene.shs.shape[1]
    fwd_b, bwd_b, bin_b = algorithmic_bytes(P, V, R, N, S, K, M)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    per_kernel = {}
    for i in range(L.GSL_K_COUNT):
        if kn[i] > 0:
            per_kernel[L.load().gsl_kernel_name(i).decode()] = dict(ms_per_launch=kms[i] / kn[i], launches=int(kn[i]))
    # algorithmic bytes of each of this repo's kernels per launch and, where a full ncu capture
    # exists (profiles/r01_*_ncu_full.md), the DRAM traffic ncu measured for it at this workload
    kern_bytes = {
        0: 45 * P + V * (16 * K) + P * (64 + 16 + 8 + 8 + 4 + 4 + 8),                 # k_preprocess_fwd
        1: 12 * P + 3 * 4 * (((W + 15) // 16) * ((H + 15) // 16)) * ((P + 255) // 256),   # k_bin_count/scan/bases (order + rect in, hist out/in/out)
        2: 12 * P + 4 * R,                                                             # k_bin_scatter
        3: 16 * P,                                                                     # depth keys + surfel sort: means3D in, order out (side stream)
        4: 12 * R,                                                                     # k_tile_blists (id + box in, entry out)
        5: N * (8 + 16 + 4 * (S + 3) + 16 + 4 + 12) + V * (64 + 16 + 4 * S),           # k_render_fwd
        6: render_bwd_algorithmic_bytes(V, N, S),                                      # k_render_bwd
        # k_preprocess_bwd: one `touched` byte per surfel; per surfel a pixel composited (measured: half of the visible ones) the
        # accumulator read and re-zeroed + radii, record, parameters and SH in, dense gradient rows out; the zero rows are a memset
        # on the side stream
        7: P * 1 + (V // 2) * (96 + 96 + 4) + (V // 2) * (48 + 45 + 16 * M) + (V // 2) * (12 + 16 + 16 + 4 * S + 4 + 16 * M + 12 + 16),
    }
    # DRAM traffic per launch as ncu measured it (read from the newest profiles/*ncu_full*.md, not hard-coded)
    md, traffic_src = ncu_traffic_from_profiles()
    # (k_bin_bases: a separate kernel in profiles of earlier builds, now the last CTA of k_bin_scan)
    groups = {0: ["k_preprocess_fwd"], 1: ["k_bin_count", "k_bin_scan", "k_bin_bases"], 2: ["k_bin_scatter"],
              3: ["k_depth_keys", "k_sort_hist", "k_sort_scan", "k_sort_scatter", "k_sort_buckets", "k_sort_rank", "k_sort_big"],
              4: ["k_tile_blists"], 5: ["k_render_fwd"], 6: ["k_render_bwd"], 7: ["k_preprocess_bwd"]}
    ncu_traffic = {i: sum(md[k] for k in ks if k in md) for i, ks in groups.items() if any(k in md for k in ks)}
    # the dominant kernel = the largest share of the step among ALL of this repo's kernel groups
    dom_id = max((i for i in kern_bytes if kn[i] > 0), key=lambda i: kms[i])
    dom_ms = kms[dom_id] / max(kn[dom_id], 1)
    dom_name = L.load().gsl_kernel_name(dom_id).decode()
    dom_bytes = kern_bytes[dom_id]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = ncu_traffic.get(dom_id)
    per_kernel_roofline = {}
    for i, b in kern_bytes.items():
        if kn[i] > 0 and kms[i] > 0:
            gbs = b / (kms[i] / kn[i] * 1e-3) / 1e9
            per_kernel_roofline[L.load().gsl_kernel_name(i).decode()] = dict(
                algorithmic_bytes=int(b), achieved_gbs=gbs, frac=gbs / hbm_peak, traffic=ncu_traffic.get(i))
    step_bytes = fwd_b + bwd_b + bin_b
    step_ms = ms / args.steps
    value = world * args.steps / (ms * 1e-3)
    res = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "ours",
        "config": shared_config(P, H, W, S, M, scene.sh_degree, V, R, world),
        "run": {"azimuth_wrap_around": bool(args.wrap_azimuth),
                "grad_exchange": None if exchange is None else (
                    "fp32, inside the step: packed rows + SH factors pushed over NVLink peer memory by the backward kernel, summed by "
                    "the tile owners (own kernels, no collective library)" if args.exchange == "peer" else
                    "fp32, inside the step: NCCL all-reduce of the non-SH gradients + all-gather of the SH factors"),
                "exchange_schedule": args.exchange_schedule if world > 1 else None,
                "grad_exchange_bytes": 0 if exchange is None else (exchange.flat_nbytes + (0 if exchange.packed else exchange.local.numel() * 4)),
                "profiling": "timed region runs with per-kernel events OFF; `kernels` come from a separate pass"},
        "cuda_graph": {"mode": args.graph, "ms_per_step_graph_on": (ms if use_graph else ms_other) / args.steps,
                       "ms_per_step_graph_off": (ms_other if use_graph else ms) / args.steps,
                       "note": "value / e2e are measured in `mode`; graph replay = one launch per pass into static buffers, the "
                               "instance count of a forward read at the next call (gs_lidar_b200.set_cuda_graphs(True, "
                               "deferred_count=True))"},
        "exchange_parity": exchange_parity,
        "per_rank_ms_without_exchange": per_rank_alone,
        "step_spread": spread,
        "clocks": clocks,
        "e2e": {"value": world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "value_synchronous_steps": world * e2e_steps / (e2e_sync_ms * 1e-3),
                "note": "per-step camera + cotangent maps from pinned host memory, rendered maps + a gradient checksum read back to pinned "
                        "host memory; surfel parameters stay resident like model weights.  value: the host consumes the results of step "
                        "k-2 before issuing step k (prefetching loader, asynchronous logging); value_synchronous_steps: a full "
                        "synchronisation of all streams at the end of every step"},
        "gpu_launches": (L.OWN_LAUNCHES_FWD + L.OWN_LAUNCHES_BWD +
                         (0 if exchange is None or not exchange.packed else L.OWN_LAUNCHES_PEER(len(exchange.ranges(P)), args.exchange_schedule))) * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                     "note": "the two compositors are FP32-issue / L2-reduction bound at this shape, not HBM-bound (ncu: profiles/); "
                             "the streaming kernels' fractions are in kernel_rooflines"},
        "kernel_rooflines": per_kernel_roofline,
        "m2_half_panorama_pair": m2,
        "step_roofline": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (step_ms * 1e-3) / 1e9, "peak": hbm_peak,
                          "unit": "GB/s", "frac": step_bytes / (step_ms * 1e-3) / 1e9 / hbm_peak},
        "kernels": per_kernel,
    }
    return res


INFER = {
    "c2": dict(P=100000, H=66, W=1030, vfov=None, metric="panoramas/sec forward-only at 100k surfels 66x1030",
               workload="forward-only render: 100k synthetic surfels -> 66x1030 KITTI-360 range/intensity/raydrop panorama, 1 B200"),
    "c5": dict(P=4000000, H=128, W=2048, vfov="opv2v", metric="panoramas/sec forward-only at 4M surfels 128x2048, frames sharded over the GPUs",
               workload="stress inference: 4M surfels, 128x2048 OPV2V-style panoramas, 512-frame batch sharded by frame across 8 B200"),
}


def run_inference(args, rank, world, local):
    """--config c2 / c5: forward-only frames through the public API (gs_lidar_b200.batch.render_frames: frames on two
    alternating streams), one camera pose per frame, frames sharded frame_id % world, no collective on the data path."""
    from gs_lidar_b200 import synth, parallel, batch
    from gs_lidar_b200 import _lib as L
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    cfg = INFER[args.config]
    dev = torch.device("cuda", local)
    P, H, W, S = cfg["P"], cfg["H"], cfg["W"], 4
    if args.surfels != 1000000:
        P = args.surfels
    vfov = synth.OPV2V_VFOV if cfg["vfov"] == "opv2v" else synth.KITTI_VFOV
    scene = synth.make_scene(P, H=H, W=W, S=S, vfov=vfov, seed=0).to(dev)
    total = args.frames if args.config == "c5" else max(args.steps, 1) * world
    mine = parallel.shard_frames(total, rank, world, equal_steps=True)
    # frame k of a drive: +0.1 sf k along x, small yaw; only the 4x4 view matrix and camera centre differ
    base = synth.settings_for(scene)
    cams = []
    for k in mine:
        c = synth.make_scene(16, H=H, W=W, S=S, vfov=vfov, seed=0, view_yaw_deg=0.5 * math.sin(0.7 * k),
                             view_shift=(0.1 * scene.scale_factor * (k % 64), 0.0, 0.0))
        cams.append(base._replace(viewmatrix=c.viewmatrix.to(dev), projmatrix=c.projmatrix.to(dev), campos=c.campos.to(dev)))
    surfels = dict(means3D=scene.means3D, opacities=scene.opacities, shs=scene.shs, features=scene.features,
                   scales=scene.scales, rotations=scene.rotations, mask=scene.mask)
    stats = {}

    def consume(i, out):
        stats["V"] = out[5]
        stats["last"] = out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local, args.clock_sample_ms, args.clock_sampler)
    if rank == 0:
        sampler.start()
    warm = cams[:max(args.warmup, 3)] if len(cams) >= 3 else cams
    batch.render_frames(warm, consume=consume, **surfels)
    batch.render_frames(warm, consume=consume, **surfels)
    barrier()
    V = int((stats["V"] > 0).sum())
    r_hint = max([v for k, v in G._pool.r_hint.items() if k[1] == P] + [0])
    R = int(r_hint / 1.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    batch.render_frames(cams, consume=consume, **surfels)
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    ms = e0.elapsed_time(e1)
    # the same frames on two alternating streams (the default of the first half of round 2)
    e0.record()
    batch.render_frames(cams, consume=consume, streams=2, **surfels)
    e1.record()
    barrier()
    ms_1s = e0.elapsed_time(e1)
    # separate pass: per-kernel CUDA-event times (one stream, profiling on -- not part of any timed region)
    Lb = L.load()
    Lb.gsl_profile_read(None, None, 1)
    Lb.gsl_profile_enable(1)
    batch.render_frames(cams[:min(len(cams), 16)], consume=consume, streams=1, **surfels)
    barrier()
    Lb.gsl_profile_enable(0)
    kms = (C.c_double * L.GSL_K_COUNT)()
    kn = (C.c_int64 * L.GSL_K_COUNT)()
    Lb.gsl_profile_read(kms, kn, 1)
    per_kernel = {Lb.gsl_kernel_name(i).decode(): dict(ms_per_launch=kms[i] / kn[i], launches=int(kn[i]))
                  for i in range(L.GSL_K_COUNT) if kn[i] > 0}
    # e2e: camera matrices from pinned host memory per frame, rendered maps back to pinned host memory per frame
    n_e2e = min(len(cams), 64)
    h_cam = [(c.viewmatrix.cpu().pin_memory(), c.campos.cpu().pin_memory()) for c in cams[:n_e2e]]
    h_out = [torch.empty((4 + S + 3 + 4 + 1, H, W)).pin_memory() for _ in range(2)]
    d_cam = [(torch.empty(4, 4, device=dev), torch.empty(3, device=dev)) for _ in range(n_e2e)]
    h2d = 16 * 4 + 3 * 4
    d2h = h_out[0].numel() * 4

    def e2e_pass():
        sets = []
        for i in range(n_e2e):
            d_cam[i][0].copy_(h_cam[i][0], non_blocking=True)
            d_cam[i][1].copy_(h_cam[i][1], non_blocking=True)
            sets.append(base._replace(viewmatrix=d_cam[i][0], projmatrix=d_cam[i][0], campos=d_cam[i][1]))

        def to_host(i, out):
            buf = h_out[i % 2]
            o = 0
            for t in (out[1], out[2], out[3], out[4]):
                n = t.shape[0]
                buf[o:o + n].copy_(t, non_blocking=True)
                o += n

        batch.render_frames(sets, consume=to_host, **surfels)
        torch.cuda.synchronize(dev)

    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    e2e_pass()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else {}
    t = torch.tensor([ms, max(ms, wall), ms_1s, e2e_wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, _, ms_1s, e2e_wall = [float(x) for x in t]
    if rank != 0:
        return None
    frames = len(cams)
    N, K, M = H * W, (scene.sh_degree + 1) ** 2, scene.shs.shape[1]
    fwd_b, _, bin_b = algorithmic_bytes(P, V, R, N, S, K, M)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ms_frame = ms / frames
    gbs = (fwd_b + bin_b) / (ms_frame * 1e-3) / 1e9
    return {
        "metric": cfg["metric"], "value": world * frames / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": frames,
        "warmup": 2 * len(warm), "ms_per_step": ms_frame, "higher_is_better": True, "scaling": "weak" if args.config == "c2" else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
        "config": {"workload": cfg["workload"], "surfels": P, "height": H, "width": W, "sh_degree": scene.sh_degree,
                   "feature_channels": S, "visible_surfels": V, "tile_instances": R, "frames_total": frames * world,
                   "parallelism": "frames sharded frame_id %% %d, no collective" % world,
                   "l2": "inputs (%.0f MB of surfel parameters per frame) exceed the 126 MB L2; no explicit flush" % ((45 * P + 16 * M * P + 4 * S * P) / 1e6)},
        "run": {"streams": "auto (one stream)", "ms_per_frame_two_streams": ms_1s / frames,
                "note": "gs_lidar_b200.batch.render_frames; ms_per_frame_two_streams = the same frames issued again on two alternating streams (streams=2)"},
        "clocks": clocks,
        "e2e": {"value": world * n_e2e / (e2e_wall * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e,
                "note": "per frame: view matrix + camera centre from pinned host memory, all rendered maps back to pinned host memory; wall clock incl. the final synchronisation"},
        "gpu_launches": L.OWN_LAUNCHES_FWD * frames,
        "roofline": {"bound": "hbm", "kernel": "whole forward frame", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "traffic": None, "algorithmic_bytes_per_launch": int(fwd_b + bin_b), "ms_per_launch": ms_frame,
                     "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
                     "note": "SURVEY.md 8(d) forward bytes (FWD_IO + BIN) of one frame over the frame time"},
        "kernels": per_kernel,
    }


def run_dynamic(args, rank, world, local):
    """--config c4: one training step of a DYNAMIC scene per rank and step -- render() (fused glue + rasterizer) forward and
    backward down to the RAW GaussianModel parameters; frame (step * world + rank) % 51 of a drive with its own pose and
    timestamp.  N > 1: the fused peer exchange sums the raw-parameter gradients of the ranks' frames, the
    frame-dependent part of the glue's VJP folded into the rows (parallel.PeerExchange.set_glue); checked, outside the timed
    region, against an all-reduce of the raw-parameter gradients every rank's own autograd graph gives."""
    from types import SimpleNamespace
    from gs_lidar_b200 import synth, parallel, renderer
    from gs_lidar_b200 import _lib as L
    import gs_lidar_b200.diff_gaussian_rasterization_2d as G
    dev = torch.device("cuda", local)
    P, H, W, S = args.surfels, args.height, args.width, 4
    NF = 51  # KITTI-360 10750-10800 (scene/kitti360_loader.py:154-156)
    scene = synth.make_scene(P, H=H, W=W, S=S, seed=0).to(dev)
    g = torch.Generator().manual_seed(4)
    n_ = lambda *s_: torch.randn(*s_, generator=g)
    pc = SimpleNamespace()  # raw GaussianModel parameters (scene/gaussian_model.py:266-298), surfels of the synthetic scene
    pc._xyz = scene.means3D.clone().requires_grad_(True)
    pc._velocity = (0.01 * n_(P, 3)).to(dev).requires_grad_(True)
    pc._t = (torch.rand(P, 1, generator=g) * 1.2 - 0.6).to(dev).requires_grad_(True)
    pc._scaling_t = (math.log(0.1) + 0.3 * n_(P, 1)).to(dev).requires_grad_(True)
    pc._opacity = torch.logit(scene.opacities.clamp(1e-4, 1 - 1e-4)).clone().requires_grad_(True)
    pc._scaling = scene.scales.log().clone().requires_grad_(True)
    pc._rotation = scene.rotations.clone().requires_grad_(True)
    pc._features_dc = scene.shs[:, :1].clone().requires_grad_(True)
    pc._features_rest = scene.shs[:, 1:].clone().requires_grad_(True)
    pc.T, pc.velocity_decay, pc.active_sh_degree = 0.2, 1.0, scene.sh_degree
    names = ["_xyz", "_velocity", "_t", "_scaling_t", "_opacity", "_scaling", "_rotation", "_features_dc", "_features_rest"]
    pipe = SimpleNamespace(neg_fov=True, debug=False, scale_factor=scene.scale_factor, dynamic=True, median_depth=False,
                           compute_cov3D_python=False, convert_SHs_python=False)
    cams = []
    for k in range(NF):
        c = synth.make_scene(16, H=H, W=W, S=S, seed=0, view_yaw_deg=0.5 * math.sin(0.7 * k),
                             view_shift=(0.1 * scene.scale_factor * k, 0.0, 0.0))
        cams.append(SimpleNamespace(image_height=H, image_width=W, world_view_transform=c.viewmatrix.to(dev),
                                    full_proj_transform=c.projmatrix.to(dev), camera_center=c.campos.to(dev), vfov=scene.vfov,
                                    hfov=scene.hfov, timestamp=-0.5 + k / (NF - 1.0), towards="forward", FoVx=1.0, FoVy=1.0))
    cot = {k: v.to(dev) for k, v in synth.make_cotangents(H, W, S, seed=1).items()}
    exchange = parallel.PeerExchange().enable() if world > 1 else None
    state = {"k": 0, "last": None}

    def step(frame=None):
        for nme in names:
            getattr(pc, nme).grad = None
        k = state["k"] if frame is None else frame
        state["k"] = k + 1
        cam = cams[(k * world + rank) % NF]
        other = [torch.exp(pc._scaling_t).detach(), pc._velocity.detach()]  # train.py:167-169
        pkg = renderer.render(cam, pc, pipe, scene.bg, other=other)
        # fixed cotangents on the maps the loss of train.py reads (depth, intensity, ray-drop, alpha, features, normal),
        # handed to autograd directly: no loss kernels in the step, like the headline's
        torch.autograd.backward([pkg["depth"], pkg["intensity_sh"], pkg["raydrop"], pkg["alpha"], pkg["feature"], pkg["normal"]],
                                [cot["depth"][:1], cot["color"][2:3], cot["color"][3:4], cot["alpha"], cot["feature"][:S],
                                 cot["feature"][S:]])
        state["last"] = pkg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local, args.clock_sample_ms, args.clock_sampler)
    if rank == 0:
        sampler.start()
    # The instance count of a dynamic frame depends on its timestamp (marginal-t mask): from one step of a rank to the next
    # (world frames further down the drive) it can grow by more than the 25 % head room the deferred count of the graph mode
    # relies on (measured at 8 ranks: the library then raises, as designed).  So the count is read in every forward here.
    G.set_cuda_graphs(args.graph == "on", deferred_count=False)
    for _ in range(max(args.warmup, 3)):
        step()
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 1.0:
        for _ in range(10):
            step()
        torch.cuda.synchronize(dev)
    barrier()
    V = int(state["last"]["visibility_filter"].sum())
    exchange_parity = None
    if world > 1:
        leaves = {nme: getattr(pc, nme) for nme in names}
        exchange_parity = check_exchange_parity(lambda: step(frame=7), leaves, names, exchange, dev)
        if not exchange_parity["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC_C4, "impl": "ours", "n_gpus": world, "error": "exchange parity check failed",
                                  "exchange_parity": exchange_parity}))
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(1)
        for _ in range(2):
            step()
        barrier()
    L.load().gsl_profile_enable(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # separate pass: per-kernel CUDA-event times of rank 0
    G.set_cuda_graphs(False)
    Lb = L.load()
    Lb.gsl_profile_read(None, None, 1)
    Lb.gsl_profile_enable(1)
    for _ in range(min(args.steps, 10)):
        step()
    barrier()
    Lb.gsl_profile_enable(0)
    kms = (C.c_double * L.GSL_K_COUNT)()
    kn = (C.c_int64 * L.GSL_K_COUNT)()
    Lb.gsl_profile_read(kms, kn, 1)
    per_kernel = {Lb.gsl_kernel_name(i).decode(): dict(ms_per_launch=kms[i] / kn[i], launches=int(kn[i]))
                  for i in range(L.GSL_K_COUNT) if kn[i] > 0}
    clocks = sampler.stop() if rank == 0 else {}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    if exchange is not None:
        exchange.disable()
        exchange.close()
    if rank != 0:
        return None
    return {
        "metric": METRIC_C4, "value": world * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
        "config": {"workload": "C4: dynamic KITTI-360-shaped training step through render(): fused glue (SHM motion model, "
                               "marginal, activations) + rasterizer, fwd+bwd to the raw GaussianModel parameters, 51 frames with "
                               "their own poses and timestamps", "surfels": P, "height": H, "width": W,
                   "sh_degree": scene.sh_degree, "feature_channels": S, "visible_surfels_last_frame": V, "frames": NF,
                   "parallelism": "frame-parallel dp%d, 1 frame/rank/step" % world,
                   "l2": "inputs exceed the 126 MB L2; no explicit flush"},
        "run": {"grad_exchange": None if world == 1 else "peer-memory exchange with the glue's frame-dependent VJP folded into "
                                                        "the rows (96-byte rows: + dL/dvelocity, dL/dt, dL/dscaling_t)",
                "cuda_graph": args.graph + " (forward replayed, instance count read every forward; a backward with the folded glue "
                                           "is issued kernel by kernel)"},
        "exchange_parity": exchange_parity, "clocks": clocks, "kernels": per_kernel,
    }


def cpu_baseline(args, full=True):
    """CPU oracle (plain-C port of the reference algorithm) timed on the host cores on a bounded sample."""
    import oracle
    from gs_lidar_b200 import synth
    o = oracle.CpuOracle()
    H, W, S = args.height, args.width, 4

    def run(P):
        s = synth.make_scene(P, H=H, W=W, S=S, seed=0)
        cot = {k: v.numpy() for k, v in synth.make_cotangents(H, W, S, seed=1).items()}
        p = o.params(P, S, s.sh_degree, s.shs.shape[1], W, H, s.vfov, s.hfov, s.scale_factor, math.tan(-0.5), math.tan(-0.5))
        a = [s.means3D.numpy(), s.scales.numpy(), s.rotations.numpy(), s.opacities.numpy(), s.shs.numpy()]
        t0 = time.perf_counter()
        st = o.forward(p, a[0], a[1], a[2], a[3], a[4], None, s.features.numpy(), s.mask.numpy(), s.viewmatrix.numpy(),
                       s.campos.numpy(), s.bg.numpy())
        o.backward(p, st, a[0], a[1], a[2], a[4], s.features.numpy(), s.viewmatrix.numpy(), s.campos.numpy(), s.bg.numpy(),
                   cot["color"], cot["depth"], cot["alpha"], cot["feature"])
        return time.perf_counter() - t0

    t_small = run(50000)
    P = args.cpu_sample_surfels
    if P <= 0:
        # CPU time grows ~linearly in surfels until pixels saturate; keep the sample within ~25 s
        P = args.surfels if t_small * (args.surfels / 50000.0) < 25.0 else int(50000 * 25.0 / max(t_small, 1e-3))
        P = max(50000, min(P, args.surfels))
    t = run(P)
    reps = 1
    if t < 10.0:  # bounded sample of ~10-15 s of CPU work
        reps = max(1, min(40, int(12.0 / max(t, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(reps):
            run(P)
        t = (time.perf_counter() - t0) / reps
    sample = "%d panorama(s) fwd+bwd, %d of %d surfels, %dx%d" % (reps, P, args.surfels, H, W)
    if P < args.surfels:
        sample += " (value = 1/time of this reduced sample; the full workload is slower)"
    return {"value": 1.0 / t, "unit": UNIT, "cores": o.threads, "kind": "port", "sample": sample, "seconds": t * reps}


def cpu_toy_baseline(args, budget_s=15.0):
    """BASELINE.json configs[0]: the reference's pure-PyTorch panorama surfel splatting (scripts/compare_2dgs_3dgs.py
    `surface_splatting`; restated for CPU tensors in oracle/toy_splat.py and pinned against the reference function's own
    outputs) timed on the host cores at the panorama shape.  Forward only -- the toy has no backward -- and every pixel
    evaluates every surfel, so it is quoted at the surfel counts it can hold; a context number, not a target."""
    from oracle import toy_splat
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, W = args.height, args.width
    out, spent = [], 0.0
    for P in (64, 1024, 4096):
        a = toy_splat.make_inputs(P, W, H, seed=0)
        t0 = time.perf_counter()
        img, _, _, _ = toy_splat.surface_splatting(*a)
        t = time.perf_counter() - t0
        spent += t
        out.append({"surfels": P, "seconds": t, "value": 1.0 / t, "covered_pixels": int((torch.nan_to_num(img).sum(-1) > 0).sum())})
        if spent + 4.5 * t > budget_s:  # the next size costs ~4x
            break
    return {"what": "pure-PyTorch toy panorama splat (compare_2dgs_3dgs.py surface_splatting), forward only, %dx%d, fov +-90 x +-20 deg" % (H, W),
            "unit": "panoramas/s", "cores": cores, "kind": "port", "runs": out}


def run_reference(args, rank, world, local):
    """Reference arm: the UNMODIFIED reference CUDA rasterizer (oracle/_ref) on the same GPU(s), same scene, same
    cotangents.  The reference has no multi-GPU code, so at N > 1 this arm runs N independent replicas, one frame per
    rank and step, with NO gradient exchange (the most favourable reading for the reference); value = frames of all
    ranks / max-over-ranks time.  Falls back to the CPU oracle port (rank 0) when the compiled reference is absent."""
    import oracle
    from gs_lidar_b200 import synth
    if not os.path.exists(oracle.REF_SO):
        if rank != 0:
            return None
        cb = cpu_baseline(args)
        return {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": 1, "steps": 1, "warmup": 0,
                "ms_per_step": cb["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "compiled reference (oracle/_ref) missing: CPU port of the reference algorithm timed instead"}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    dev = torch.device("cuda", local)
    P, H, W, S = args.surfels, args.height, args.width, 4
    scene = synth.make_scene(P, H=H, W=W, S=S, seed=0, **make_frame_pose(0))
    if rank != 0:
        cam = synth.make_scene(16, H=H, W=W, S=S, seed=0, **make_frame_pose(rank))
        scene = scene._replace(viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, campos=cam.campos)
    scene = scene.to(dev)
    cot = {k: v.to(dev) for k, v in synth.make_cotangents(H, W, S, seed=1).items()}
    ref = oracle.RefCuda()
    a = common.ref_args(scene)
    bufs = {}
    last = {}

    def step():
        f = ref.forward(a, zero_fill=True, outs=bufs.get("o"))
        last["R"] = f["R"]
        bufs["o"] = {k: v for k, v in f.items() if k != "R"}
        bufs["g"] = ref.backward(a, f, cot, zero_fill=True, grads=bufs.get("g"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local, args.clock_sample_ms, args.clock_sampler)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    V, R = int((bufs["o"]["radii"][:P] > 0).sum()), int(last["R"])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = max(e0.elapsed_time(e1), 0.0)
    wall = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else {}
    t = torch.tensor([ms, max(ms, wall)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    value = world * args.steps / (ms * 1e-3)
    # M2: the pair of half panoramas the reference's own render_range_map issues (see run_ours)
    m2 = None
    if W % 2 == 0 and world == 1:
        flip = torch.diag(torch.tensor([-1.0, 1.0, -1.0, 1.0], device=dev))
        vm_back = (flip @ scene.viewmatrix.t()).t().contiguous()
        a_f = dict(a, W=W // 2, hfov=(-90.0, 90.0))
        a_b = dict(a_f, viewmatrix=vm_back, projmatrix=vm_back)
        cot_h = {k: v[..., :W // 2].contiguous() for k, v in cot.items()}
        refs = [(oracle.RefCuda(), a_f, {}), (oracle.RefCuda(), a_b, {})]

        def m2_step():
            for r, aa, bb in refs:
                f = r.forward(aa, zero_fill=True, outs=bb.get("o"))
                bb["o"] = {k: v for k, v in f.items() if k != "R"}
                bb["g"] = r.backward(aa, f, cot_h, zero_fill=True, grads=bb.get("g"))

        for _ in range(3):
            m2_step()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(max(3, args.steps // 2)):
            m2_step()
        e1.record()
        torch.cuda.synchronize()
        m2_ms = e0.elapsed_time(e1) / max(3, args.steps // 2)
        m2 = {"value": 1e3 / m2_ms, "unit": UNIT, "ms_per_panorama": m2_ms,
              "what": "M2: two half panoramas %dx%d (hfov +-90, front + back camera), fwd+bwd each" % (H, W // 2)}
    if rank != 0:
        return None
    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": shared_config(P, H, W, S, scene.shs.shape[1], scene.sh_degree, V, R, world),
            "run": {"what": "unmodified reference CUDA rasterizer (diff-gaussian-rasterization-2d) compiled for sm_100a by oracle/build_ref.sh, "
                            "buffers pre-allocated, outputs/gradients zero-filled per step like its torch binding"
                            + ("" if world == 1 else "; %d independent replicas (the reference is single-GPU), one frame per rank and "
                                                     "step, no gradient exchange" % world)},
            "clocks": clocks,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                             "sample": "full workload on the GPU (the reference path is CUDA; it has no CPU implementation)"},
            "m2_half_panorama_pair": m2,
            "e2e": {"value": world * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    args = parse_args()
    if args.toy_only:
        print(json.dumps(cpu_toy_baseline(args)))
        return
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU fallback"}))
        sys.exit(1)
    rank, world, local = init_dist(args)
    if args.impl == "reference":
        res = run_reference(args, rank, world, local)
    elif args.config == "c4":
        res = run_dynamic(args, rank, world, local)
    elif args.config != "c3":
        res = run_inference(args, rank, world, local)
    else:
        res = run_ours(args, rank, world, local)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            try:
                res["cpu_baseline"] = cpu_baseline(args)
            except Exception as ex:  # the baseline must never take the measurement down
                res["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
            try:
                res["cpu_toy_baseline"] = cpu_toy_baseline(args)
            except Exception as ex:
                res["cpu_toy_baseline"] = {"runs": [], "kind": "port", "what": "failed: %r" % (ex,)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0 and res is not None:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
