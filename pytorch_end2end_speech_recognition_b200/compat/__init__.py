"""Import shims: put this directory on sys.path and ``import warpctc_pytorch`` resolves to the
B200 engine (the reference does ``import warpctc_pytorch`` at models/pytorch_v3/ctc/ctc.py:11)."""
