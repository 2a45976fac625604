"""Drop-in module named ``warpctc_pytorch``.

The reference imports ``warpctc_pytorch`` (models/pytorch_v3/ctc/ctc.py:11, models/pytorch/ctc/ctc.py:11)
and uses ``warpctc_pytorch._CTC`` (:30), ``gpu_ctc`` / ``cpu_ctc`` (:35) and ``CTCLoss`` (:69).  Adding
``pytorch_end2end_speech_recognition_b200/compat`` to ``sys.path`` (or installing this directory as the
top-level package ``warpctc_pytorch``) makes those imports resolve to the B200 engine; the binding that
``tools/install_warpctc_pytorch.sh`` used to build is no longer needed.
"""
from pytorch_end2end_speech_recognition_b200.ctc import CTCLoss, _CTC, cpu_ctc, gpu_ctc  # noqa: F401

__all__ = ["CTCLoss", "_CTC", "cpu_ctc", "gpu_ctc"]
