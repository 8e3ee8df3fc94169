"""Forward-only rendering of many LiDAR frames (cameras) of one surfel set -- the inference side of SURVEY.md 8(e):
"frames sharded frame_id % G; no collective".  Frames are independent: each is issued with its own workspace, without the
blocking copy the reference has in the middle of every forward (rasterizer_impl.cu:314-315) -- the one host wait of a
forward (the instance count) is polled after all of its launches.  With streams=2 consecutive frames alternate between two
CUDA streams, so that the per-surfel stages of frame k+1 run under the compositing of frame k; whether that pays depends on
how much of the GPU a single frame leaves idle (see the note at "auto" below).

    frames = render_frames(settings_list, means3D=..., opacities=..., shs=..., scales=..., rotations=..., features=...)

Each element of the result is the 6-tuple GaussianRasterizer returns.  With torch.distributed, give every rank its own
shard of the camera list (parallel.shard_frames); nothing here communicates.
"""
import torch

from .diff_gaussian_rasterization_2d import GaussianRasterizer

_streams = {}


def _side_streams(device, n):
    key = (device.index if device.index is not None else torch.cuda.current_device(), n)
    if key not in _streams:
        _streams[key] = [torch.cuda.Stream(device) for _ in range(n)]
    return _streams[key]


@torch.no_grad()
def render_frames(settings_list, means3D, opacities, means2D=None, streams="auto", consume=None, **surfels):
    """Renders one frame per element of `settings_list` (GaussianRasterizationSettings: one camera each).  `surfels` are the
    keyword arguments of GaussianRasterizer.forward shared by all frames (shs / colors_precomp, features, scales, rotations,
    mask, shs_rest).  streams: how many alternating streams ("auto": 1, see below).
    `consume(i, outputs)`, if given, is called on the frame's stream right after frame i was issued (to
    copy results out, accumulate metrics ...) and the outputs are not kept; otherwise the list of outputs is returned."""
    dev = means3D.device
    if streams == "auto":
        # small scenes are host-bound (a 100k-surfel frame is 0.2 ms of GPU work, about what issuing it costs): the stream
        # switches then cost more than the overlap returns (measured: 0.37 vs 0.20 ms per frame at 100k surfels).  At 4M
        # surfels two alternating streams used to win (2.49 vs 3.02 ms per frame) while the surfel sort left the GPU half
        # idle; with the warp-per-bucket sort one stream measured faster (2.26 vs 2.43 ms per frame, 64 frames,
        # profiles/r02_bench_c5_n1_64frames.json), so "auto" is one stream at every size now; streams=2 remains selectable
        streams = 1
    if means2D is None:
        means2D = torch.zeros((means3D.shape[0], 4), dtype=means3D.dtype, device=dev)
    main = torch.cuda.current_stream(dev)
    side = _side_streams(dev, max(1, int(streams)))
    start = torch.cuda.Event()
    start.record(main)
    results = []
    for i, settings in enumerate(settings_list):
        s = side[i % len(side)]
        if i < len(side):
            s.wait_event(start)  # the surfel tensors were produced on the caller's stream
        with torch.cuda.stream(s):
            out = GaussianRasterizer(settings)(means3D=means3D, means2D=means2D, opacities=opacities, **surfels)
            if consume is not None:
                consume(i, out)
            else:
                results.append(out)
    for s in side:
        main.wait_stream(s)
    for t in (means3D, opacities, means2D) + tuple(v for v in surfels.values() if isinstance(v, torch.Tensor)):
        for s in side:
            t.record_stream(s)
    return None if consume is not None else results
