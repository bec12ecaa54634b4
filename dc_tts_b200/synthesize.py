"""Drop-in for the reference's synthesize.py (/root/reference/synthesize.py:21-64).

Same flow: load text -> build Graph(mode="synthesize") -> restore parameters -> mel
loop (:45-54) -> SSRN (:57) -> write one output per sentence.  Differences, all forced
by what exists offline: parameters are restored from the TF checkpoints under
hp.logdir-1 / hp.logdir-2 when they exist (dc_tts_b200/checkpoint.py reads the tensor
bundles without TensorFlow), else come from a name->array dict or the seeded initialiser.  The Griffin-Lim vocoder
(utils.py:67-114) runs on the GPU (dc_tts_b200/utils.py) and the wavs are written with
scipy.io.wavfile like the reference does.
"""
import os

import numpy as np

from .checkpoint import latest_checkpoint
from .data_load import load_data
from .engine import get_engine
from .hyperparams import Hyperparams as hp
from .params import init_params
from .train import Graph, Session


def synthesize(params=None, sentences=None, fast=True, write=True, seed=0, vocoder=True, allow_random_init=False):
    """`params`: a name -> array dict to use instead of the checkpoints.  Without it the latest checkpoints of
    hp.logdir-1 (Text2Mel) and hp.logdir-2 (SSRN) are restored, and a missing one RAISES like the reference's
    `saver.restore(sess, None)` does (synthesize.py:33,39) -- seeded random weights are used only when the caller asks
    for them (`allow_random_init=True`, benchmarks and smoke tests) or has already loaded parameters into the engine."""
    # Load data
    L = load_data("synthesize", sentences)

    # Load graph
    g = Graph(mode="synthesize")
    print("Graph loaded")

    with Session() as sess:
        # Restore parameters (synthesize.py:31-41)
        if params is not None:
            g.engine.load_params(params)
            print("Parameters loaded from the caller's dictionary")
        elif g.engine.params_loaded:
            print("Using the parameters already committed to the engine")
        else:
            ck1, ck2 = latest_checkpoint(hp.logdir + "-1"), latest_checkpoint(hp.logdir + "-2")
            if ck1 and ck2:
                g.engine.restore(hp.logdir + "-1", hp.logdir + "-2")
                print("Text2Mel Restored!")
                print("SSRN Restored!")
            elif allow_random_init:
                g.engine.load_params(init_params(seed))
                print("WARNING: no checkpoint -- seeded random weights (allow_random_init=True); the output is noise")
            else:
                missing = [d for d, c in ((hp.logdir + "-1", ck1), (hp.logdir + "-2", ck2)) if not c]
                raise FileNotFoundError("no checkpoint under %s (reference: Saver.restore(sess, None) fails); pass params=..., "
                                        "or allow_random_init=True for seeded random weights" % " and ".join(missing))

        if fast:
            # the whole loop on the device (CUDA-graph replay), identical results
            Y, _ = g.generate(L)
        else:
            # the reference's loop, verbatim in structure (synthesize.py:45-54)
            Y = np.zeros((len(L), hp.max_T, hp.n_mels), np.float32)
            prev_max_attentions = np.zeros((len(L),), np.int32)
            for j in range(hp.max_T):
                _gs, _Y, _max_attentions, _alignments = \
                    sess.run([g.global_step, g.Y, g.max_attentions, g.alignments],
                             {g.L: L, g.mels: Y, g.prev_max_attentions: prev_max_attentions})
                Y[:, j, :] = _Y[:, j, :]
                prev_max_attentions = _max_attentions[:, j]

        # Get magnitude (synthesize.py:57)
        Z = sess.run(g.Z, {g.Y: Y})

    # Generate wav files (synthesize.py:60-64): Griffin-Lim on the GPU for the whole batch
    if write:
        if not os.path.exists(hp.sampledir):
            os.makedirs(hp.sampledir)
        if vocoder:
            from scipy.io.wavfile import write as write_wav
            from .utils import spectrograms2wavs
            for i, wav in enumerate(spectrograms2wavs(Z)):
                print("Working on file", i + 1)
                write_wav(os.path.join(hp.sampledir, "{}.wav".format(i + 1)), hp.sr, wav)
        else:
            for i, mag in enumerate(Z):
                np.save(os.path.join(hp.sampledir, "{}.mag.npy".format(i + 1)), mag)
    return (Y.cpu().numpy() if hasattr(Y, "cpu") else Y), Z


if __name__ == '__main__':
    synthesize()
    print("Done")
