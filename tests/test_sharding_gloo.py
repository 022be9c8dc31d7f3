"""N > 1 host logic on CPU: two gloo ranks shard a clip batch with shard_range, "compute" their block (here the CPU
oracle stands in for the per-rank device plan, because this container has no GPU), and gather_to_rank0 reassembles the
batch in clip order. This is the off-hot-path collective of DESIGN.md section 6; the hot path itself has none."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_clips, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from spectrograms_b200.sharding import gather_to_rank0, shard_range
    rng = np.random.default_rng(7)
    clips = rng.standard_normal((n_clips, 4000)).astype(np.float32)          # every rank builds the same batch
    lo, hi = shard_range(n_clips, rank, world)
    desc = oracle.Desc(dtype="f32", n_fft=400, hop=160, mapping="mel", n_bands=16, f_min=0.0, f_max=8000.0, amp="db", floor_db=-80.0)
    local = oracle.compute_batch(desc, clips[lo:hi], 1) if hi > lo else np.zeros((0, 16, 26), np.float32)
    full = gather_to_rank0(torch.from_numpy(local), n_clips)
    if rank == 0:
        ref = oracle.compute_batch(desc, clips, 1)
        np.save(out_path, np.array([float(np.abs(full.numpy() - ref).max()), full.shape[0]]))
    else:
        assert full is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [8, 5])
def test_two_rank_shard_and_gather(tmp_path, n_clips):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_clips, out), nprocs=2, join=True)
    err, n = np.load(out)
    assert n == n_clips and err == 0.0


def test_shard_range_partitions_every_clip_once():
    from spectrograms_b200.sharding import shard_range
    for n in (1, 5, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))          # contiguous, no overlap, no gap
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_bind_host_to_device_degrades_without_nvml():
    # On a box without NVML (this container) or without the affinity call it must return False and leave the process
    # affinity untouched; on a GPU box it returns True and the affinity is whatever NVML reports for that GPU.
    from spectrograms_b200.sharding import bind_host_to_device
    before = os.sched_getaffinity(0)
    ok = bind_host_to_device(0)
    assert isinstance(ok, bool)
    if not ok:
        assert os.sched_getaffinity(0) == before
