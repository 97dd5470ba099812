"""Host-side logic of the multi-GPU path on CPU: sample slicing and the single sum-reduce, with world_size 2 over gloo."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mg():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module("go-raytracing_b200.multigpu")


def test_slices_partition_the_samples():
    mg = _mg()
    for spp in [0, 1, 3, 10, 100, 500, 1024]:
        for world in [1, 2, 4, 8]:
            seen = []
            for r in range(world):
                base, count = mg.slice_samples(spp, r, world)
                seen.extend(range(base, base + count))
            assert seen == list(range(spp)), (spp, world)
            counts = [mg.slice_samples(spp, r, world)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        mg.slice_samples(8, 2, 2)


def _fake_pass(npix, base, count):
    """Stand-in for a render pass: a deterministic per-(pixel, global sample) contribution, like the Philox-keyed paths."""
    acc = np.zeros((npix, 4), dtype=np.float32)
    pix = np.arange(npix, dtype=np.float64)
    for s in range(base, base + count):
        v = np.modf(np.sin(pix * 12.9898 + s * 78.233) * 43758.5453)[0]
        acc[:, 0] += v.astype(np.float32)
        acc[:, 1] += (v * v).astype(np.float32)
        acc[:, 2] += np.float32(0.5)
        acc[:, 3] += 1
    return acc.reshape(-1)


def _worker(rank, world, port, spp, npix, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    mg = _mg()
    r, w, _ = mg.init_from_env("gloo")
    assert (r, w) == (rank, world)
    base, count = mg.slice_samples(spp, rank, world)
    t = torch.from_numpy(_fake_pass(npix, base, count))
    dist.barrier()
    mg.reduce_sum(t, dst=0)
    if rank == 0:
        np.save(out, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("spp", [1, 7, 16])
def test_two_rank_reduce_equals_single_rank(tmp_path, spp):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    npix, out = 257, str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, port, spp, npix, out), nprocs=2, join=True)
    got = np.load(out)
    want = _fake_pass(npix, 0, spp)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6)
    assert np.all(got.reshape(-1, 4)[:, 3] == spp)   # every pixel received all its samples exactly once
