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


def allreduce_images(images):
    """Sum detector images over ranks in place.  uint64 counts (held as int64 tensors) are exact and
    order-independent; interferogram planes are float64 sums (order-dependent at 1e-16, inside the 1e-3 L1 budget)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for img in images:
        for t in (img.tensors() if hasattr(img, "tensors") else [img]):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)


def solve_and_image_sharded(domain, beam, probing_depth, diagnostics, n_total=None, **kw):
    """Each rank traces its shard of a device ``Beam`` and the images are all-reduced; every rank ends up with
    the full images.  Returns this rank's stats tensor (device) -- sum them for global counters."""
    from . import propagator
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    off, cnt = shard(beam.Np if n_total is None else n_total, rank, world)
    stats, _ = propagator.solve_and_image(domain, beam, probing_depth, diagnostics, n_rays=cnt, ray_offset=off,
                                          sync=False, **kw)
    allreduce_images([d.image for d in diagnostics])
    return stats
