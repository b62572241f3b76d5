"""All-reduce hook for rlpt_set_allreduce built on torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU
tests). The library hands over a raw device pointer on its own stream; the hook wraps it as a tensor without copying
and runs the collective in that stream's order, so no host synchronisation is needed between tracing, all-reduce and
the merge kernel. PyTorch is plumbing here: the reduction itself is NCCL's.
"""
import numpy as np


class _DevicePtr:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def torch_allreduce_hook(device_index, group=None):
    """hook(ptr, count, dtype, stream) for Context.set_allreduce; dtype 0 = float32, 1 = uint32 (summed as int32)."""
    import torch
    import torch.distributed as dist
    cache = {}

    def hook(ptr, count, dtype, stream):
        key = (ptr, count, dtype)
        t = cache.get(key)
        if t is None:
            t = torch.as_tensor(_DevicePtr(ptr, count, "<f4" if dtype == 0 else "<i4"), device="cuda:%d" % device_index)
            cache[key] = t
        with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device_index)):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return hook


def host_allreduce_hook(group=None):
    """The same contract over host memory (gloo), for the world_size-2 CPU tests of the merge logic: `ptr` is a host
    address of `count` elements."""
    import ctypes
    import torch
    import torch.distributed as dist

    def hook(ptr, count, dtype, stream):
        ctype = ctypes.c_float if dtype == 0 else ctypes.c_int32
        a = np.ctypeslib.as_array((ctype * count).from_address(ptr))
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return hook


def p2p_setup(ctx, group=None):
    """Peer-memory exchange for the ranks of one node (Context.p2p_export / p2p_import): gathers every rank's CUDA IPC
    handles with torch.distributed and opens them. After this the per-frame Q exchange needs no collective call.
    All ranks agree on the outcome: if any rank cannot open its peers' buffers, every rank stays on the all-reduce hook.
    Returns True when the peer-memory exchange is active."""
    import torch
    import torch.distributed as dist
    ok = 1
    try:
        blob = ctx.p2p_export()
    except Exception as e:  # noqa: BLE001
        print("rlpt p2p export failed:", e); blob = None; ok = 0
    blobs = [None] * dist.get_world_size(group)
    dist.all_gather_object(blobs, blob, group=group)
    if ok and all(b is not None for b in blobs):
        try:
            ctx.p2p_import(blobs)
        except Exception as e:  # noqa: BLE001
            print("rlpt p2p import failed:", e); ok = 0
    else:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        ctx.p2p_close()
        return False
    dist.barrier(group=group)
    return True
