"""oracle/ -- CPU restatements of the reference DC-TTS synthesis path.

TEST INFRASTRUCTURE ONLY.  Nothing under dc_tts_b200/ imports this package; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
PARITY UNPINNED: the reference holds no golden vectors and cannot be executed here
(TensorFlow 1.x unavailable), see ref_torch.py's header and DESIGN.md.
"""
