"""Backward (training) entry points of the BIN hot path.

Round-1 status: the forward (inference) path is complete; the backward kernels (dgrad / wgrad
implicit GEMMs) are not built yet.  Per the "fail loudly, no fallback" rule a grad-enabled call
raises instead of silently routing through PyTorch autograd."""
from ._lib import BinB200Error

_MSG = ("bin_b200: grad-enabled forward requested, but the sm_100a backward kernels are not built yet. "
        "Run under torch.no_grad() (inference); there is deliberately no PyTorch fallback.")


def backbone_apply(module, frames):
    raise BinB200Error(_MSG)


def window_apply(module, frames):
    raise BinB200Error(_MSG)


def convlstm_apply(module, x, state):
    raise BinB200Error(_MSG)
