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
