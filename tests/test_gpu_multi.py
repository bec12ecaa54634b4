"""Two-GPU shard equivalence (skipped with fewer than 2 GPUs): the utterance-sharded run with ONE NCCL gather (chunked,
overlapped with the SSRN: dc_tts_b200.parallel.OverlappedGather) reproduces the single-GPU results BIT FOR BIT when the
per-rank batch size is the one the single-GPU run uses -- utterances never interact (networks.py:140-153,
synthesize.py:54), and sharding / gathering does no arithmetic."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
STEPS = 40
PER_RANK = 4


def _worker(rank, world, port, L, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dc_tts_b200.engine import Engine
        from dc_tts_b200.hyperparams import Hyperparams as hp
        from dc_tts_b200.parallel import OverlappedGather, gather_spectrograms, shard_bounds
        from dc_tts_b200.params import init_params
        dev = torch.device("cuda", rank)
        e = Engine(rank)
        e.load_params(init_params(0, "perturbed"))
        lo, hi = shard_bounds(len(L), rank, world)
        og = OverlappedGather(len(L), (hp.max_T * hp.r, 1 + hp.n_fft // 2), torch.float32, dev, chunks=2)
        for _ in range(2):                                          # twice: the receive buffer is reused across steps
            Y, P, _, _ = e.text2mel_generate(L[lo:hi], steps=STEPS)
            # the whole shard goes through the SSRN in ONE call (the kernel variants of the single-GPU run of the shard);
            # the hand-over to rank 0 is chunked and runs on the side stream
            _, Zfull = e.ssrn(Y, want_logits=False)
            og.begin()
            for c in og.chunks():
                og.send(c, Zfull[c.lo:c.hi])
            Zall = og.finish()
        Yall = gather_spectrograms(Y, len(L), dst=0)
        Pall = gather_spectrograms(P.to(torch.float32), len(L), dst=0)
        if rank == 0:
            q.put((Yall.cpu().numpy(), Zall.cpu().numpy(), Pall.cpu().numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_gpu_shards_equal_single_gpu_bit_for_bit(engine):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from dc_tts_b200.params import synthetic_text
    engine.set_tensor_path(1); engine.set_option("decode_mode", 1)
    L = synthetic_text(2 * PER_RANK, 70, seed=11)
    ref_Y, ref_Z, ref_P = [], [], []
    for lo in (0, PER_RANK):                                        # the single-GPU runs, at the per-rank batch size
        Y1, P1, _, _ = engine.text2mel_generate(L[lo:lo + PER_RANK], steps=STEPS)
        _, Z1 = engine.ssrn(Y1, want_logits=False)
        ref_Y.append(Y1.cpu().numpy()); ref_Z.append(Z1.cpu().numpy()); ref_P.append(P1.cpu().numpy())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, L, q)) for r in range(2)]
    for p in procs:
        p.start()
    Y2, Z2, P2 = q.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert np.array_equal(P2.astype(np.int32), np.concatenate(ref_P))
    assert np.array_equal(Y2, np.concatenate(ref_Y))               # bit for bit
    assert np.array_equal(Z2, np.concatenate(ref_Z))
