"""Short, profiler-friendly slice of the hot path (run under ncu on the GPU box):
B utterances, TextEnc + `steps` AR steps + SSRN, once warm and once measured."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dc_tts_b200.engine import Engine  # noqa: E402
from dc_tts_b200.params import init_params, synthetic_text  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--tensor-path", type=int, default=1)
ap.add_argument("--no-ssrn", action="store_true")
ap.add_argument("--decode-mode", type=int, default=1, help="1 persistent cluster decode kernel, 0 one CUDA graph per frame")
a = ap.parse_args()
e = Engine(0)
e.load_params(init_params(0, "perturbed"))
e.set_tensor_path(a.tensor_path)
e.set_option("decode_mode", a.decode_mode)
L = synthetic_text(a.batch, 100, seed=0)
for it in range(2):
    torch.cuda.nvtx.range_push("pass%d" % it)
    Y, _, _, _ = e.text2mel_generate(L, steps=a.steps)
    if not a.no_ssrn:
        e.ssrn(Y, want_logits=False)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("launches", e.launch_count())
