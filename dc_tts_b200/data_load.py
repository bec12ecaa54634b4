"""Text -> id adaptor of the synthesis path (reference data_load.py:19-31, :79-86).

Only the pure-Python pieces `synthesize.py` needs are provided; the TF queue-runner
training pipeline (`get_batch`, data_load.py:88-131) is out of scope.
"""
import codecs
import re
import unicodedata

import numpy as np

from .hyperparams import Hyperparams as hp


def load_vocab():
    """data_load.py:19-22."""
    char2idx = {ch: i for i, ch in enumerate(hp.vocab)}
    idx2char = dict(enumerate(hp.vocab))
    return char2idx, idx2char


def text_normalize(text):
    """data_load.py:24-31: strip accents, lower-case, out-of-vocab -> space, squeeze spaces."""
    text = "".join(ch for ch in unicodedata.normalize("NFD", text) if unicodedata.category(ch) != "Mn")
    text = re.sub("[^{}]".format(hp.vocab), " ", text.lower())
    return re.sub("[ ]+", " ", text)


def load_data(mode="synthesize", path=None):
    """data_load.py:79-86: sentences file (first line is a header and is dropped, the
    leading "N. " of each line is removed) -> int32 ids (num_sentences, max_N), each
    sentence terminated by E and zero padded."""
    if mode != "synthesize":
        raise NotImplementedError("only the synthesize branch is on the hot path (SURVEY.md 2.1)")
    char2idx, _ = load_vocab()
    lines = codecs.open(path or hp.test_data, "r", "utf-8").readlines()[1:]
    sents = [text_normalize(line.split(" ", 1)[-1]).strip() + "E" for line in lines]
    texts = np.zeros((len(sents), hp.max_N), np.int32)
    for i, sent in enumerate(sents):
        texts[i, :len(sent)] = [char2idx[ch] for ch in sent]
    return texts
