"""2+ GPU probe: does torch symmetric memory hand out a multicast (NVLS) address here, and what do multimem.st /
multimem.ld_reduce achieve through it?  torchrun --nproc-per-node N scripts/probe/run_mm_probe.py"""
import ctypes as C
import json
import os
import sys
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmm_probe.so"))
out = {"world": world}
try:
    n_f4 = 4 << 20  # 64 MB
    t = symm.empty(n_f4 * 4 * 2, dtype=torch.float32, device=dev)
    hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
    out["multicast_ptr"] = int(hdl.multicast_ptr)
    out["buffer_ptrs"] = [int(p) for p in hdl.buffer_ptrs]
    out["local_ptr"] = t.data_ptr()
    try:
        out["has_multicast_support"] = bool(symm._SymmetricMemory.has_multicast_support(DeviceType := torch._C._autograd.DeviceType.CUDA, local))
    except Exception as ex:
        out["has_multicast_support"] = repr(ex)
    mc = hdl.multicast_ptr
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if mc:
        A = t[: n_f4 * 4]
        B = t[n_f4 * 4:]
        A.fill_(float(rank + 1))
        B.zero_()
        torch.cuda.synchronize(); dist.barrier()
        res = torch.empty(n_f4 * 4, device=dev)
        lib.mm_ld.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        lib.mm_st.argtypes = [C.c_void_p, C.c_float, C.c_long, C.c_void_p]
        lib.mm_p2p.argtypes = [C.c_void_p, C.c_float, C.c_long, C.c_void_p]
        lib.mm_or.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p]
        # ld_reduce over 1/world of the rows (what a tile owner does)
        per = n_f4 // world
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            lib.mm_ld(C.c_void_p(mc + rank * per * 16), C.c_void_p(res.data_ptr()), per, st)
            e1.record(); torch.cuda.synchronize()
        out["ld_reduce_ms"] = e0.elapsed_time(e1)
        out["ld_reduce_bytes_out"] = per * 16
        out["ld_reduce_ok"] = bool((res[: per * 4] == world * (world + 1) / 2).all())
        # multimem.st of my 1/world slice into B on every rank
        for it in range(3):
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            lib.mm_st(C.c_void_p(mc + (n_f4 + rank * per) * 16), float(rank + 1), per, st)
            e1.record(); torch.cuda.synchronize()
        out["mm_st_ms"] = e0.elapsed_time(e1)
        dist.barrier(); torch.cuda.synchronize()
        want = torch.arange(1, world + 1, device=dev, dtype=torch.float32).repeat_interleave(per * 4)
        out["mm_st_ok"] = bool((B[: per * 4 * world] == want).all())
        # plain P2P stores of the same slice to every peer, one after the other (what the exchange does today)
        for it in range(3):
            dist.barrier(); torch.cuda.synchronize()
            e0.record()
            for g in range(world):
                lib.mm_p2p(C.c_void_p(out["buffer_ptrs"][g] + (n_f4 + rank * per) * 16), float(rank + 1), per, st)
            e1.record(); torch.cuda.synchronize()
        out["p2p_st_all_ms"] = e0.elapsed_time(e1)
        # variants: weak qualifiers, grid sizes, fused ld_reduce + st
        for name, fn in (("st_weak", "mm_st_weak"), ("ld_weak", "mm_ld_weak"), ("ld_st", "mm_ld_st")):
            for blocks in (148, 148 * 4, 148 * 16):
                f = getattr(lib, fn)
                f.argtypes = [C.c_void_p, C.c_float if fn == "mm_st_weak" else C.c_void_p, C.c_long, C.c_int, C.c_void_p]
                for it in range(3):
                    dist.barrier(); torch.cuda.synchronize()
                    e0.record()
                    if fn == "mm_st_weak":
                        f(C.c_void_p(mc + (n_f4 + rank * per) * 16), float(rank + 1), per, blocks, st)
                    elif fn == "mm_ld_weak":
                        f(C.c_void_p(mc + rank * per * 16), C.c_void_p(res.data_ptr()), per, blocks, st)
                    else:
                        f(C.c_void_p(mc + rank * per * 16), C.c_void_p(mc + (n_f4 + rank * per) * 16), per, blocks, st)
                    e1.record(); torch.cuda.synchronize()
                out["%s_b%d_ms" % (name, blocks)] = e0.elapsed_time(e1)
        dist.barrier(); torch.cuda.synchronize()
        want = torch.full((per * 4 * world,), world * (world + 1) / 2, device=dev)
        out["ld_st_ok"] = bool((B[: per * 4 * world] == want).all())
        out["slice_bytes"] = per * 16
        # multimem.red.or
        B.zero_(); torch.cuda.synchronize(); dist.barrier()
        lib.mm_or(C.c_void_p(mc + n_f4 * 16), C.c_uint32(1 << rank), 1024, st)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        out["mm_or_ok"] = bool((B[:1024].view(torch.int32) == (1 << world) - 1).all())
except Exception as ex:
    out["error"] = repr(ex)
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    print(json.dumps(gathered[0]))
    for g in gathered[1:]:
        print(json.dumps({k: g.get(k) for k in ("ld_reduce_ms", "mm_st_ms", "p2p_st_all_ms", "ld_reduce_ok", "mm_st_ok", "mm_or_ok", "error")}))
dist.barrier()
dist.destroy_process_group()
