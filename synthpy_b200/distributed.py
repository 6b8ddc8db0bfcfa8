"""Multi-GPU plumbing for the ray path: one process per GPU (torchrun), rays sharded by rank, field replicated,
detector images combined by ONE all-reduce (SUM) -- the counterpart of the reference's
``comm.reduce(sh.H, root=0, op=MPI.SUM)`` (examples/jobs/run_scripts/interference_MPI.py:189) and of the
``+=`` loop in pvti_trace_multiprocess.py:129-134.  No collective sits inside the data path.

torch.distributed is plumbing here (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            local = int(os.environ.get("LOCAL_RANK", "0"))
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world


def shard(n_total, rank, world):
    """Contiguous, balanced partition of global ray indices: returns (offset, count) of ``rank``.
    Ray i is generated from (seed, i), so the union over ranks is the same bundle for every ``world``."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def combine_images(images, root=None):
    """Sum detector images over ranks: to every rank (``root=None``, one all-reduce per image) or to ``root`` only (one
    reduce, the reference's ``comm.reduce(H, root=0, op=MPI.SUM)``).  uint64 counts (held as int64 tensors) are exact and
    order-independent, and so are interferogram planes: int64 fixed-point sums (2^-40 units, include/synthpy_b200.h).

    The sum is taken OUT OF PLACE for ``engine.ImageBuffer``s: the per-rank accumulator that ``solve_and_image`` keeps
    adding to is left alone and the global image is attached with ``set_global`` (``ImageBuffer.result()`` then returns
    it).  Calling this after every batch of a loop therefore re-sums the accumulators and never counts a ray twice.
    Objects that only offer ``tensors()`` (or bare tensors) are summed in place."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    rank = dist.get_rank()
    for img in images:
        keep = hasattr(img, "set_global")
        outs = []
        for t in (img.tensors() if hasattr(img, "tensors") else [img]):
            g = t.clone() if keep else t
            if root is None:
                dist.all_reduce(g, op=dist.ReduceOp.SUM)
            else:
                dist.reduce(g, dst=root, op=dist.ReduceOp.SUM)
            outs.append(g)
        if keep:
            img.set_global(outs if (root is None or rank == root) else None)


def allreduce_images(images):
    """``combine_images(images, root=None)``: every rank ends up with the full images."""
    combine_images(images, root=None)


def bind_to_local_numa(local_rank):
    """Best effort: restrict this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers are
    allocated, so that first-touch places them next to the GPU (8 ranks staging rays through one node's memory cost 9 %
    of the end-to-end rate in round 1).  Returns a small dict describing what was done, or None."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return {"gpu": bus, "node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"gpu": bus, "node": node, "bound": False}
        os.sched_setaffinity(0, allowed)
        return {"gpu": bus, "node": node, "bound": True, "cpus": len(allowed)}
    except Exception as e:                                   # no sysfs, no permission, old torch: run unbound
        return {"bound": False, "why": type(e).__name__}


def solve_and_image_sharded(domain, beam, probing_depth, diagnostics, n_total=None, root=None, **kw):
    """Each rank traces its shard of a device ``Beam``; the images are then summed over ranks (``combine_images``: to
    every rank, or to ``root``).  The per-rank accumulators are not modified by the sum, so the call can be repeated
    (batches of one beam) without double counting.  Returns this rank's stats tensor (device) -- sum them for global
    counters."""
    from . import propagator
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    off, cnt = shard(beam.Np if n_total is None else n_total, rank, world)
    stats, _ = propagator.solve_and_image(domain, beam, probing_depth, diagnostics, n_rays=cnt, ray_offset=off,
                                          sync=False, **kw)
    combine_images([d.image for d in diagnostics], root=root)
    return stats
