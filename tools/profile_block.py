"""Runs one block (default SSRN/HC_11 at B=32, L=840) a few times -- target of `ncu --set full`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_tts_b200.engine import Engine  # noqa: E402
from dc_tts_b200.params import init_params  # noqa: E402

scope = sys.argv[1] if len(sys.argv) > 1 else "SSRN/HC_11"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L = int(sys.argv[3]) if len(sys.argv) > 3 else 840
path = int(sys.argv[4]) if len(sys.argv) > 4 else 1
e = Engine(0)
e.load_params(init_params(0, "perturbed"))
e.set_tensor_path(path)
print(scope, B, L, e.bench_block(scope, B, L, iters=3, warmup=1))
