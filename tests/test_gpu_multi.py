"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped otherwise): rays sharded by rank, field replicated, one
NCCL all-reduce of the detector images.  The all-reduced images must equal the single-GPU images of the same
1e6-ray bundle bit for bit (uint64 counts; device Philox rays are a function of the global ray index only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup(n_grid=96):
    sys.path.insert(0, ROOT)
    from synthpy_b200 import beam as B, diagnostics as D, domain as Dm, field_generator as FG
    ne = FG.turbulent_ne(n_grid // 2, noise="torch", seed=7)
    dom = Dm.ScalarDomain((10e-3, 10e-3, 20e-3), n_grid)
    dom.external_ne(ne)
    beam = B.Beam(int(1e6), 5e-3, 5e-5, 10e-3, device=True, seed=3)
    specs = [D.spec("shadow_two", bin_scale=4), D.spec("schlieren_DF", bin_scale=4, R_stop=1)]
    return dom, beam, specs


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    from synthpy_b200 import distributed as Dist, engine
    Dist.init(backend="nccl")
    dom, beam, specs = _setup()
    stats = Dist.solve_and_image_sharded(dom, beam, 10e-3, specs, lwl=1064e-9)
    torch.cuda.synchronize()
    tot = stats.clone()
    torch.distributed.all_reduce(tot)
    if rank == 0:
        np.savez(out_path, a=specs[0].image.total().cpu().numpy(), b=specs[1].image.total().cpu().numpy(), stats=tot.cpu().numpy())
    torch.distributed.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_images_equal_one_gpu_images(tmp_path):
    import torch.multiprocessing as mp
    from synthpy_b200 import engine, propagator as P
    out = str(tmp_path / "multi.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    dom, beam, specs = _setup()
    stats, _ = P.solve_and_image(dom, beam, 10e-3, specs, lwl=1064e-9)
    assert np.array_equal(got["a"], specs[0].image.counts.cpu().numpy())
    assert np.array_equal(got["b"], specs[1].image.counts.cpu().numpy())
    assert int(got["stats"][0]) == stats["ray_steps"] and int(got["stats"][3]) == stats["rays_binned"]
