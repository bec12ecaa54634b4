"""BASELINE config 5: Text2Mel training step, B = 32 per GPU, fixed N = 180 / T = 210, synthetic batch, dropout on,
data-parallel over the launched ranks (gradient all-reduce of the flat arena over NCCL).  Prints one JSON line.
    python tools/bench_train.py [--steps 10 --warmup 3 --batch 32]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from dc_tts_b200.engine import Engine
from dc_tts_b200.hyperparams import Hyperparams as hp
from dc_tts_b200.params import init_params, synthetic_text

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--probe", type=int, default=0, help="measurement only: 1 = the tcgen05 GEMMs fetch their operands but issue no MMA (results are garbage)")
ap.add_argument("--net", type=int, default=1, choices=[1, 2], help="1 = Text2Mel trainer (BASELINE config 5), 2 = SSRN trainer (train.py num=2) at T = 210")
ap.add_argument("--train-tc", type=int, default=7, help="bit mask: 1 forward conv, 2 data gradient, 4 weight gradient on tcgen05 (default 7 = all), 0 = fp32 CUDA-core kernels")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local)
eng.load_params(init_params(0))
B = a.batch
eng.set_option("train_tc", a.train_tc)
eng.set_option("train_probe", a.probe)
if a.net == 1:
    eng.train_init(B)
else:
    eng.train_init_ssrn(B, hp.max_T)
L = torch.from_numpy(synthetic_text(B, 100, seed=rank)).cuda()
mels = torch.from_numpy(np.random.default_rng(rank).uniform(0, 1, (B, hp.max_T, hp.n_mels)).astype(np.float32)).cuda()
mags = torch.from_numpy(np.random.default_rng(rank + 100).uniform(0, 1, (B, hp.max_T * hp.r, 1 + hp.n_fft // 2)).astype(np.float32)).cuda() if a.net == 2 else None
grads = eng.train_grads()


def step(i):
    if a.net == 2:
        out = eng.train_step_ssrn(mels, mags, global_step=4000 + i, seed=i * world + rank, apply=(world == 1))
    else:
        out = eng.train_step(L, mels, global_step=4000 + i, seed=i * world + rank, apply=(world == 1))
    if world > 1:
        dist.all_reduce(grads)
        grads.mul_(1.0 / world)
        eng.train_apply(4000 + i)
    return out


for i in range(a.warmup):
    first = step(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = eng.launch_count()
t0.record()
for i in range(a.steps):
    last = step(a.warmup + i)
t1.record()
torch.cuda.synchronize()
ms = torch.tensor([t0.elapsed_time(t1) / a.steps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(ms)
    flops = 3 * 2 * B * (hp.max_N * 17.10e6 + hp.max_T * (4.08e6 + 2.71e6 + 0.09e6))          # fwd MACs x 2 x 3
    if a.net == 2:
        flops = 3 * 2 * B * hp.max_T * 93.66e6                                                    # SSRN: 93.7 MMAC per mel frame
    print(json.dumps({"metric": "train_steps_per_sec", "value": 1e3 / ms, "unit": "steps/s", "n_gpus": world, "ms_per_step": ms,
                      "mel_frames_per_sec": world * B * hp.max_T * 1e3 / ms, "steps": a.steps, "warmup": a.warmup,
                      "config": {"workload": ("BASELINE config 5: Text2Mel train step (fwd + bwd + clip + Adam), B=%d per GPU, N=180, T=210, dropout %.2f" if a.net == 1 else
                                              "SSRN train step (train.py num=2: fwd + bwd + clip + Adam), B=%d per GPU, T=210 -> 840 frames x 1025 bins, dropout %.2f") % (B, hp.dropout_rate),
                                 "parallelism": "dp%d (all-reduce of %d gradients)" % (world, grads.numel())},
                      "dtype": ("f32 tensors; GEMMs as split-fp16 x3 on tcgen05, fp32 accumulate" if a.train_tc else "f32 (CUDA-core kernels)"), "data": "synthetic",
                      "achieved_tflops": world * flops / (ms * 1e-3) / 1e12, "gpu_launches_per_step": (eng.launch_count() - n0) // a.steps,
                      "loss_first": first["loss"], "loss_last": last["loss"]}))
if world > 1:
    dist.destroy_process_group()
