"""hctr_b200 - B200-native (sm_100a) implementation of the HCTR recognition hot path.

Drop-in mirrors of the reference's Python surface (AndrewCullacino/handwritten-chinese-ocr-samples):
  models.handwritten_ctr_model.hctr_model   (reference: models/handwritten_ctr_model.py:156-178)
  utils.ctc_codec.ctc_codec                 (reference: utils/ctc_codec.py:14-285)
  ctc_loss.CTCLoss                          (reference call: main.py:205,406-409)
All arithmetic runs in hand-written CUDA kernels behind the C ABI in include/hctr_b200.h.
"""
from . import native  # noqa: F401
from . import train_engine  # noqa: F401
from . import pipeline  # noqa: F401
from . import ngram_lm  # noqa: F401
from . import checkpoint  # noqa: F401

__all__ = ["native", "hctr_model", "ctc_codec", "CTCLoss", "TrainStep"]


def __getattr__(name):
    if name == "hctr_model":
        from .models.handwritten_ctr_model import hctr_model
        return hctr_model
    if name == "ctc_codec":
        from .utils.ctc_codec import ctc_codec
        return ctc_codec
    if name == "CTCLoss":
        from .ctc_loss import CTCLoss
        return CTCLoss
    if name == "TrainStep":
        from .train_step import TrainStep
        return TrainStep
    raise AttributeError(name)
