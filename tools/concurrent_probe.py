"""How much does running G independent groups of utterances on G streams (one handle each) help the
latency-bound decode loop?   python tools/concurrent_probe.py 32 1 2 4"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dc_tts_b200.engine import Engine
from dc_tts_b200.params import init_params, synthetic_text

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
groups = [int(x) for x in sys.argv[2:]] or [1, 2, 4]
P = init_params(0, "perturbed")
L = synthetic_text(B, 100, seed=0)
engines, streams = [], []
for G in groups:
    while len(engines) < G:
        e = Engine(0); e.load_params(P); engines.append(e); streams.append(torch.cuda.Stream())
    per = B // G
    def run():
        outs = []
        for g in range(G):
            with torch.cuda.stream(streams[g]):
                outs.append(engines[g].text2mel_generate(L[g * per:(g + 1) * per])[0])
        return outs
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    n = 4
    t0 = time.perf_counter()
    for _ in range(n):
        outs = run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("B=%d in %d concurrent groups of %d: %.2f ms  (%.0f mel-frames/s)  checksum %.4f" %
          (B, G, per, dt * 1e3, B * 210 / dt, sum(float(o.double().sum()) for o in outs)))
