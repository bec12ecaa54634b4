"""Times the finished-spectrogram gather alone (run under torchrun on >= 2 GPUs)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from dc_tts_b200.parallel import gather_spectrograms  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B = 32
Z = torch.full((B, 840, 1025), float(rank), device=dev)
for name, fn in (("p2p gather", lambda: gather_spectrograms(Z, B * world, dst=0)),
                 ("all_gather", lambda: dist.all_gather_into_tensor(torch.empty((B * world, 840, 1025), device=dev), Z))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        out = fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    if rank == 0:
        print("%s: %.3f ms per gather of %d x %.1f MB (%.1f GB/s into rank 0)" % (name, dt * 1e3, world - 1, Z.numel() * 4 / 1e6,
                                                                          (world - 1) * Z.numel() * 4 / dt / 1e9), flush=True)
if rank == 0:
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
