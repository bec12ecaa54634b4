"""Pre-computes the training targets, like the reference's prepo.py (/root/reference/prepo.py:15-25): for every wav of
the transcript, `mels/<name>.npy` (reduced mel, utils.py:147-162) and `mags/<name>.npy`.  The spectrograms come from the
GPU feature-extraction row (`dc_tts_b200.utils.load_spectrograms` -> `dctts_get_spectrograms`)."""
import os

import numpy as np

from .trainer import load_train_data


def prepo(data_dir=None, out_dir=".", load_spectrograms=None, progress=None):
    if load_spectrograms is None:
        from .utils import load_spectrograms
    fpaths, _, _ = load_train_data(data_dir)
    for sub in ("mels", "mags"):
        os.makedirs(os.path.join(out_dir, sub), exist_ok=True)
    for fpath in (progress(fpaths) if progress else fpaths):
        fname, mel, mag = load_spectrograms(fpath)
        np.save(os.path.join(out_dir, "mels", fname.replace("wav", "npy")), mel)
        np.save(os.path.join(out_dir, "mags", fname.replace("wav", "npy")), mag)
    return len(fpaths)


if __name__ == "__main__":
    print("Done", prepo())
