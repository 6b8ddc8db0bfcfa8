"""Host-side multi-rank logic on CPU with the gloo backend (world_size 2 and 3): ray sharding and the image
all-reduce that replaces the reference's MPI.SUM reduce.  The GPU kernels are not involved."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Img:
    def __init__(self, t):
        self.t = t

    def tensors(self):
        return [self.t]


class _Acc(_Img):
    """Stand-in for engine.ImageBuffer's accumulate / combine protocol (the real one needs a GPU)."""
    glob = None

    def set_global(self, tensors):
        self.glob = None if tensors is None else tensors[0]


def _worker_batches(rank, world, port, n_total, out):
    """A batched loop that combines after EVERY batch (an in-place all-reduce of the live accumulator
    re-sums earlier global totals and over-counts): two sharded batches, then a reduce to rank 0."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from synthpy_b200 import distributed as D
    D.init(backend="gloo")
    acc = _Acc(torch.zeros(64, dtype=torch.int64))
    seen = []
    for b in range(2):                                     # batch b = global rays [b * n_total, (b + 1) * n_total)
        off, cnt = D.shard(n_total, rank, world)
        idx = torch.arange(b * n_total + off, b * n_total + off + cnt, dtype=torch.int64)
        acc.t.index_add_(0, (idx * 7919) % 64, torch.ones_like(idx))
        D.combine_images([acc])
        seen.append(acc.glob.numpy().copy())
    D.combine_images([acc], root=0)
    out[rank] = (seen[0], seen[1], None if acc.glob is None else acc.glob.numpy().copy(), acc.t.numpy().copy())
    dist.destroy_process_group()


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from synthpy_b200 import distributed as D
    r, w = D.init(backend="gloo")
    assert (r, w) == (rank, world)
    off, cnt = D.shard(n_total, r, w)
    # stand-in for the traced shard: ray i lands in pixel (i * 7919) % 64 -- any partition must give the same image
    idx = torch.arange(off, off + cnt, dtype=torch.int64)
    counts = torch.zeros(64, dtype=torch.int64)
    counts.index_add_(0, (idx * 7919) % 64, torch.ones_like(idx))
    planes = torch.zeros(4, 64, dtype=torch.float64)
    planes[0].index_add_(0, (idx * 7919) % 64, torch.cos(idx.double()))
    D.allreduce_images([_Img(counts), _Img(planes)])
    out[rank] = (off, cnt, counts.numpy().copy(), planes.numpy().copy())
    dist.destroy_process_group()


def _run(world, n_total, fn=None):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(fn or _worker, args=(world, port, n_total, out), nprocs=world, join=True)
    return dict(out)


def test_shard_partition_properties():
    from synthpy_b200.distributed import shard
    for n in (0, 1, 7, 1000, 10 ** 9 + 7):
        for w in (1, 2, 3, 8):
            parts = [shard(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (o1, c1), (o2, _) in zip(parts, parts[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_allreduce_images_world2_and_3_match_single_rank():
    n = 10007
    idx = np.arange(n)
    ref_counts = np.bincount((idx * 7919) % 64, minlength=64)
    ref_plane = np.bincount((idx * 7919) % 64, weights=np.cos(idx.astype(float)), minlength=64)
    for world in (2, 3):
        out = _run(world, n)
        assert sorted(out) == list(range(world))
        assert sum(v[1] for v in out.values()) == n
        for r in range(world):
            assert np.array_equal(out[r][2], ref_counts)                      # integer sums: exact on every rank
            assert np.allclose(out[r][3][0], ref_plane, rtol=0, atol=1e-9)


def test_combine_after_every_batch_does_not_double_count():
    n = 5003
    one = np.bincount((np.arange(n) * 7919) % 64, minlength=64)
    two = np.bincount((np.arange(2 * n) * 7919) % 64, minlength=64)
    out = _run(2, n, _worker_batches)
    for r in range(2):
        first, second, rooted, local = out[r]
        assert np.array_equal(first, one) and np.array_equal(second, two)     # second combine: both batches, each ray once
        assert (rooted is not None) == (r == 0)                               # reduce(root=0): only rank 0 holds the sum
        assert local.sum() < two.sum()                                        # the per-rank accumulator was left alone
    assert np.array_equal(out[0][2], two)
    assert np.array_equal(out[0][3] + out[1][3], two)
