"""The trainer loop (dc_tts_b200/trainer.py; reference train.py:137-160 + data_load.py:41-56,97-112) with a recording
stand-in for the engine: batching / padding, step counting, checkpoint cadence and names, termination."""
import os

import numpy as np
import pytest

from dc_tts_b200 import trainer
from dc_tts_b200.hyperparams import Hyperparams as hp


class FakeEngine:
    def __init__(self):
        self.calls, self.saved, self.init = [], [], None
        self.applied, self.applies, self.grads = [], [], np.ones(4, np.float32)

    def train_init(self, B):
        self.init = ("t2m", B)

    def train_init_ssrn(self, B, T):
        self.init = ("ssrn", B, T)

    def train_step(self, L, mels, global_step=0, seed=0, apply=True):
        self.calls.append(("t2m", L.shape, mels.shape, global_step, seed))
        self.applied.append(apply)
        return {"loss": 1.0, "loss_mels": 0.3, "loss_bd1": 0.69, "loss_att": 0.01}

    def train_step_ssrn(self, mels, mags, global_step=0, seed=0, apply=True):
        self.calls.append(("ssrn", mels.shape, mags.shape, global_step, seed))
        self.applied.append(apply)
        return {"loss": 1.0, "loss_mags": 0.3, "loss_bd2": 0.7}

    def train_grads(self):
        return self.grads

    def train_apply(self, gs):
        self.applies.append(gs)

    def save_checkpoint(self, prefix, gs, scope):
        self.saved.append((prefix, gs, scope))

    restored_from = None
    resume_at = None

    def restore_training(self, logdir, scope):
        self.restored_from = (logdir, scope)
        return self.resume_at


def _dataset(tmp_path, n=7):
    d = tmp_path / "LJSpeech-1.0"
    (d / "wavs").mkdir(parents=True)
    rng = np.random.default_rng(0)
    lines, lengths = [], []
    for i in range(n):
        text = "Sentence number %d, with Ünïcode & digits 123." % i if i != 3 else "x" * 400          # one too long
        lines.append("LJ%03d|raw|%s" % (i, text))
        t = 40 + 10 * i if i != 5 else hp.max_T + 8                                                   # one too long in time
        lengths.append(t)
    (d / "transcript.csv").write_text("\n".join(lines) + "\n", encoding="utf-8")
    store = {"LJ%03d.wav" % i: (rng.uniform(0, 1, (t, hp.n_mels)).astype(np.float32),
                                rng.uniform(0, 1, (4 * t, 1 + hp.n_fft // 2)).astype(np.float32)) for i, t in enumerate(lengths)}
    loader = lambda fpath: (os.path.basename(fpath),) + store[os.path.basename(fpath)]
    return str(d), loader, store


def test_transcript_parser_and_fixed_size_batches(tmp_path):
    d, loader, store = _dataset(tmp_path)
    fpaths, lens, texts = trainer.load_train_data(d)
    assert len(fpaths) == 7 and fpaths[0].endswith("wavs/LJ000.wav") and lens[3] == 401
    assert texts[0].dtype == np.int32 and texts[0][-1] == hp.vocab.index("E")
    assert "".join(hp.vocab[i] for i in texts[1]) == "sentence number with unicode digits .E"   # digits, comma, & -> spaces, squeezed; accents stripped
    batches = list(trainer.fixed_size_batches(fpaths, texts, B=2, seed=1, loader=loader, epochs=1))
    assert len(batches) == 2                                       # 7 utterances - 2 skipped = 5 -> two full batches, remainder dropped
    seen = []
    for L, mels, mags, names in batches:
        assert L.shape == (2, hp.max_N) and mels.shape == (2, hp.max_T, hp.n_mels) and mags.shape == (2, 4 * hp.max_T, 1025)
        for b, name in enumerate(names):
            mel, mag = store[name]
            assert np.array_equal(mels[b, :len(mel)], mel) and not mels[b, len(mel):].any()
            assert np.array_equal(mags[b, :len(mag)], mag) and not mags[b, len(mag):].any()
            i = int(name[2:5])
            assert np.array_equal(L[b, :lens[i]], texts[i]) and not L[b, lens[i]:].any()
        seen += names
    assert "LJ003.wav" not in seen and "LJ005.wav" not in seen and len(set(seen)) == 4


@pytest.mark.parametrize("num", [1, 2])
def test_loop_cadence_and_termination(tmp_path, num):
    d, loader, _ = _dataset(tmp_path)
    fpaths, _, texts = trainer.load_train_data(d)
    eng = FakeEngine()
    logdir = str(tmp_path / ("LJ01-%d" % num))
    gs = trainer.train(num, eng, trainer.fixed_size_batches(fpaths, texts, B=2, seed=0, loader=loader), num_iterations=7,
                       logdir=logdir, global_step=1996, save_every=1000, log=lambda *_: None)
    # runs until gs > num_iterations is first checked after a step: starting at 1996 the first step already exceeds 7
    assert gs == 1997 and len(eng.calls) == 1
    eng = FakeEngine()
    gs = trainer.train(num, eng, trainer.fixed_size_batches(fpaths, texts, B=2, seed=0, loader=loader), num_iterations=2003,
                       logdir=logdir, global_step=1996, save_every=1000, log=lambda *_: None)
    assert gs == 2004 and [c[3] for c in eng.calls] == list(range(1996, 2004))       # global_step fed BEFORE the increment
    assert [c[4] for c in eng.calls] == list(range(1996, 2004))                      # per-step dropout seed
    assert eng.saved == [(os.path.join(logdir, "model_gs_002k"), 2000, "Text2Mel" if num == 1 else "SSRN")]
    assert eng.init == (("t2m", 2) if num == 1 else ("ssrn", 2, hp.max_T)) and os.path.isdir(logdir)
    assert eng.calls[0][0] == ("t2m" if num == 1 else "ssrn")
    with pytest.raises(ValueError):
        trainer.train(3, eng, [])


def test_prepo_writes_what_the_trainer_reads(tmp_path):
    """prepo.py:15-25 -> data_load.py:105-109: files named after the wav, loadable by the trainer's default loader."""
    from dc_tts_b200 import prepo
    d, loader, store = _dataset(tmp_path, n=3)
    n = prepo.prepo(d, str(tmp_path), load_spectrograms=lambda p: (os.path.basename(p),) + store[os.path.basename(p)])
    assert n == 3 and sorted(os.listdir(tmp_path / "mels")) == ["LJ000.npy", "LJ001.npy", "LJ002.npy"]
    fname, mel, mag = trainer._load_spectrograms_npy(os.path.join(d, "wavs", "LJ001.wav"), str(tmp_path / "mels"), str(tmp_path / "mags"))
    assert fname == "LJ001.wav" and np.array_equal(mel, store[fname][0]) and np.array_equal(mag, store[fname][1])


def test_resume_and_empty_dataset(tmp_path):
    """a logdir that already holds a checkpoint is resumed (tf.train.Supervisor, train.py:144), and a data set
    that can never fill a batch raises instead of spinning."""
    d, loader, _ = _dataset(tmp_path)
    fpaths, _, texts = trainer.load_train_data(d)
    eng = FakeEngine(); eng.resume_at = 4000
    logdir = str(tmp_path / "LJ01-1")
    gs = trainer.train(1, eng, trainer.fixed_size_batches(fpaths, texts, B=2, seed=0, loader=loader), num_iterations=4002,
                       logdir=logdir, log=lambda *_: None)
    assert eng.restored_from == (logdir, "Text2Mel") and [c[3] for c in eng.calls] == [4000, 4001, 4002] and gs == 4003
    eng = FakeEngine(); eng.resume_at = 4000
    gs = trainer.train(1, eng, trainer.fixed_size_batches(fpaths, texts, B=2, seed=0, loader=loader), num_iterations=1,
                       logdir=logdir, global_step=0, log=lambda *_: None)
    assert eng.restored_from is None and gs == 2                     # an explicit global_step starts over
    with pytest.raises(ValueError):
        next(trainer.fixed_size_batches(fpaths, texts, B=6, seed=0, loader=loader))   # only 5 utterances fit


def test_bucketed_batches_follow_the_reference_queue(tmp_path):
    """data_load.py:120-129: bucket boundaries every 20 characters from minlen+1, a batch = B utterances of ONE bucket,
    padded to the longest member of that batch; pad_to_fixed extends the padding to the CUDA step's fixed shapes."""
    rng = np.random.default_rng(1)
    n = 60
    lens = [int(x) for x in rng.integers(12, 150, n)]
    texts = [rng.integers(2, 30, l).astype(np.int32) for l in lens]
    frames = [int(1.2 * l) + 5 for l in lens]
    fpaths = ["wavs/U%03d.wav" % i for i in range(n)]
    store = {os.path.basename(p): (np.full((t, hp.n_mels), i + 1, np.float32), np.full((4 * t, 5), i + 1, np.float32))
             for i, (p, t) in enumerate(zip(fpaths, frames))}
    loader = lambda p: (os.path.basename(p),) + store[os.path.basename(p)]
    bounds = trainer.bucket_boundaries(lens)
    assert bounds == list(range(min(lens) + 1, max(lens) - 1, 20))
    assert trainer.bucket_index(bounds[0] - 1, bounds) == 0 and trainer.bucket_index(bounds[0], bounds) == 1
    assert trainer.bucket_index(10 ** 6, bounds) == len(bounds)
    seen = []
    for L, mels, mags, names, k in trainer.bucketed_batches(fpaths, lens, texts, B=4, seed=3, loader=loader, epochs=2):
        idx = [int(nm[1:4]) for nm in names]
        assert all(trainer.bucket_index(lens[i], bounds) == k for i in idx)             # one bucket per batch
        assert L.shape == (4, max(lens[i] for i in idx))                                  # dynamic_pad: longest member
        assert mels.shape == (4, max(frames[i] for i in idx), hp.n_mels) and mags.shape[1] == 4 * mels.shape[1]
        for b, i in enumerate(idx):
            assert np.array_equal(L[b, :lens[i]], texts[i]) and not L[b, lens[i]:].any()
            assert (mels[b, :frames[i]] == i + 1).all() and not mels[b, frames[i]:].any()
        fixed = trainer.pad_to_fixed(L, mels, mags)
        if L.shape[1] <= hp.max_N and mels.shape[1] <= hp.max_T:
            Lf, mf, gf = fixed
            assert Lf.shape == (4, hp.max_N) and mf.shape == (4, hp.max_T, hp.n_mels) and gf.shape[1] == 4 * hp.max_T
            assert np.array_equal(Lf[:, :L.shape[1]], L) and not Lf[:, L.shape[1]:].any() and not mf[:, mels.shape[1]:].any()
        else:
            assert fixed is None
        seen += idx
    assert len(seen) >= 2 * n - 4 * (len(bounds) + 1)                                     # only partial buckets are left over
    assert max(np.bincount(seen)) <= 2                                                    # each utterance at most once per epoch


def test_graph_train_surface(tmp_path):
    """train.py:22-135 / :148: `Graph(num)` is the training object; `sess.run([g.global_step, g.train_op])` runs one
    optimiser step on the next batch and returns the incremented step; the losses and the Noam rate are fetchable."""
    from dc_tts_b200.train import Graph, Session
    from dc_tts_b200.utils import guided_attention, learning_rate_decay
    from oracle import ref_train as rtr
    d, loader, _ = _dataset(tmp_path)
    fpaths, _, texts = trainer.load_train_data(d)
    for num in (1, 2):
        eng = FakeEngine()
        g = Graph(num=num, engine=eng, batches=trainer.fixed_size_batches(fpaths, texts, B=2, seed=0, loader=loader), global_step=3998)
        with Session() as sess:
            with pytest.raises(ValueError):
                sess.run(g.loss)                                                     # nothing has run yet
            gs, _ = sess.run([g.global_step, g.train_op])
            assert gs == 3999 and eng.calls[-1][3] == 3998 and eng.calls[-1][0] == ("t2m" if num == 1 else "ssrn")
            gs, _, loss, lr = sess.run([g.global_step, g.train_op, g.loss, g.lr])
            assert gs == 4000 and loss == np.float32(1.0) and lr == np.float32(learning_rate_decay(hp.lr, 4000))
            assert sess.run(g.loss_mels if num == 1 else g.loss_mags) == np.float32(0.3)
        assert eng.init == (("t2m", 2) if num == 1 else ("ssrn", 2, hp.max_T))
    with pytest.raises(ValueError):
        Graph(num=3, engine=FakeEngine(), batches=[])
    for gs in (0, 1, 3999, 4000, 123456):
        assert learning_rate_decay(hp.lr, gs) == pytest.approx(rtr.learning_rate(gs), rel=1e-6)
    np.testing.assert_allclose(guided_attention(), rtr.guided_attention(), rtol=0, atol=1e-7)


def test_data_parallel_loop_wiring(tmp_path):
    """BASELINE config 5 in the loop: disjoint per-rank batches, step without apply, all-reduce of the gradient arena,
    identical apply, per-rank dropout seeds, checkpoints from rank 0 only."""
    d, loader, _ = _dataset(tmp_path, n=11)
    fpaths, _, texts = trainer.load_train_data(d)
    names = []
    for rank in (0, 1):
        names.append([n for b in trainer.fixed_size_batches(fpaths, texts, B=2, seed=4, loader=loader, epochs=1, rank=rank, world=2) for n in b[3]])
    assert names[0] and names[1] and not set(names[0]) & set(names[1])
    reduced = []
    for rank in (0, 1):
        eng = FakeEngine()
        gs = trainer.train(1, eng, trainer.fixed_size_batches(fpaths, texts, B=2, seed=4, loader=loader, rank=rank, world=2),
                           num_iterations=1000, logdir=str(tmp_path / "dp"), global_step=998, save_every=1000, log=lambda *_: None,
                           rank=rank, world=2, allreduce=lambda g: reduced.append(g.sum()))
        assert gs == 1001 and eng.applied == [False] * 3 and eng.applies == [998, 999, 1000]
        assert [c[4] for c in eng.calls] == [998 * 2 + rank, 999 * 2 + rank, 1000 * 2 + rank]
        assert len(eng.saved) == (1 if rank == 0 else 0)
    assert len(reduced) == 6
