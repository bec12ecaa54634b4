"""world_size-2 gloo test of the utterance-shard plumbing (no GPU)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dc_tts_b200.parallel import gather_spectrograms, shard_bounds
from dc_tts_b200.params import synthetic_text


def test_shard_bounds_cover_and_balance():
    for total in (1, 7, 32, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_text_is_shard_invariant():
    full = synthetic_text(10, 50, seed=3)
    lo, hi = shard_bounds(10, 1, 3)
    assert np.array_equal(synthetic_text(hi - lo, 50, seed=3, first_index=lo), full[lo:hi])


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None].expand(-1, 4, 5).contiguous() * 1.0
        out = gather_spectrograms(local, total, dst=0)
        if rank == 0:
            want = torch.arange(total, dtype=torch.float32)[:, None, None].expand(-1, 4, 5)
            q.put(bool(torch.equal(out, want)))
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_gather_two_ranks_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300 + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


# ---- data-parallel training (BASELINE config 5): average of the per-rank gradients == gradient of the global batch ----
def _train_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dc_tts_b200.hyperparams import Hyperparams as hp
        from dc_tts_b200.parallel import allreduce_mean_
        from dc_tts_b200.params import init_params
        from oracle import ref_train as rtr
        P = init_params(0, "perturbed")
        L = synthetic_text(world, 30, seed=2)
        mels = np.random.default_rng(5).uniform(0, 1, (world, hp.max_T, hp.n_mels)).astype(np.float32)
        names = rtr.text2mel_names()
        lo, hi = shard_bounds(world, rank, world)
        _, _, info = rtr.train_step(P, L[lo:hi], mels[lo:hi], rate=0.0)
        flat = torch.cat([torch.from_numpy(info["grads"][n]).reshape(-1) for n in names])
        allreduce_mean_(flat)
        if rank == 0:
            _, _, full = rtr.train_step(P, L, mels, rate=0.0)
            want = torch.cat([torch.from_numpy(full["grads"][n]).reshape(-1) for n in names])
            scale = float(want.abs().max())
            q.put((float((flat - want).abs().max()) / scale, flat.numel()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_average_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    err, n = q.get(timeout=5)
    assert n == 23970288 and err < 1e-5


# ---- overlapped chunked gather: same result as the plain gather, ragged shards included ----
def _og_worker(rank, world, port, total, q, chunks=4):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dc_tts_b200.parallel import OverlappedGather
        og = OverlappedGather(total, (3, 5), torch.float32, "cpu", chunks=chunks)
        lo, hi = shard_bounds(total, rank, world)
        for step in range(2):                                   # the receive buffer is reused across steps
            og.begin()
            for c in og.chunks():
                rows = torch.arange(lo + c.lo, lo + c.hi, dtype=torch.float32)[:, None, None].expand(-1, 3, 5) + 100.0 * step
                view = og.local_view(c)
                if view is not None:
                    view.copy_(rows); og.send(c, view)
                else:
                    og.send(c, rows.contiguous())
            out = og.finish()
            if rank == 0:
                want = torch.arange(total, dtype=torch.float32)[:, None, None].expand(-1, 3, 5) + 100.0 * step
                q.put(bool(torch.equal(out, want)))
            else:
                assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,chunks", [(16, 4), (7, 4), (2, 4), (16, 1)])      # one chunk = bench.py's choice up to 4 GPUs
def test_overlapped_gather_two_ranks_gloo(total, chunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 300 + total + 40 * chunks
    procs = [ctx.Process(target=_og_worker, args=(r, 2, port, total, q, chunks)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True and q.get(timeout=5) is True
