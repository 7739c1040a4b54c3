"""One process per GPU: torch.distributed carries the NCCL unique id from rank 0 to the other ranks (plumbing); the
communicator itself lives in libitsolv_b200.so and is what the kernels' all-reduces run on."""
from __future__ import annotations

import os

from .api import Context


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def bind_host_to_device(local_rank: int) -> dict:
    """Pin the calling process to the CPU cores that are closest to its GPU (NVML's ideal affinity: the cores of the
    NUMA node the GPU's PCIe root hangs on). Host buffers allocated afterwards - the pinned CSR arrays and solution vectors
    of the end-to-end path - are then first-touched on that node, so that with one process per GPU the eight uploads do
    not all cross the socket interconnect. Returns what was done (for the bench record); never raises."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = local_rank
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                phys = int(ids[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"bound": True, "cpus": f"{allowed[0]}-{allowed[-1]} ({len(allowed)})"}
    except Exception as e:  # no NVML, no permission: run unbound
        info["why"] = repr(e)[:120]
    return info


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """broadcast a byte string over the default torch.distributed group (works on gloo and nccl)"""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def attach_communicator(ctx: Context, peer_memory: bool = True) -> None:
    """Give the context a communicator spanning the torch.distributed world (no-op for a single process)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        ctx.init_comm(0, 1, b"\0" * 128)
        return
    uid = ctx.unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_bytes(uid, 128, 0)
    ctx.init_comm(dist.get_rank(), dist.get_world_size(), uid)
    if peer_memory and dist.get_world_size() <= 8 and dist.get_backend() == "nccl":
        # the fused all-reduce of the Gram kernels writes into the other ranks' exchange buffers over NVLink. That needs
        # all ranks on one host with peer access between their GPUs: compare host names first, then gather every rank's
        # CUDA IPC handle (64 bytes) and map them; if any rank fails, every rank unmaps and the communicator stays on
        # its ncclAllReduce path.
        import socket
        import torch
        world = dist.get_world_size()
        names = [None] * world
        dist.all_gather_object(names, socket.gethostname())
        if len(set(names)) != 1:
            return
        mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).cuda()
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        ok = ctx.p2p_import(b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh))
        status = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(status, op=dist.ReduceOp.MIN)
        if int(status.item()) == 0:
            ctx.p2p_disable()
        dist.barrier()
