"""Times whole networks on the selected kernel set (CUDA events, warm): SSRN and TextEnc at batch B."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from dc_tts_b200.engine import Engine  # noqa: E402
from dc_tts_b200.params import init_params, synthetic_text  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
e = Engine(0)
e.load_params(init_params(0, "perturbed"))
Y = torch.from_numpy(np.random.default_rng(0).uniform(0, 1, (B, 210, 80)).astype(np.float32)).cuda()
L = synthetic_text(B, 100, 0)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("B=%d  SSRN %.3f ms   TextEnc %.3f ms   env: %s" % (B, timeit(lambda: e.ssrn(Y, want_logits=False)), timeit(lambda: e.textenc(L)),
      {k: v for k, v in os.environ.items() if k.startswith("DCTTS_")}))
