"""B200-native (sm_100a) DPRNN-TasNet / DPRNN-Spe / DPRNN-Spe-IRA separation forward path.

Drop-in for the model classes of Aleksashka-i/tss-with-dprnn (``src/models``): same constructors,
``state_dict`` layout and ``forward`` signatures, with every stage of the forward executed by the
hand-written CUDA kernels of ``csrc/`` through the C ABI declared in ``include/dprnn_b200.h``.
"""
from .models import DPRNNTasNet, DPRNNSpeTasNet, DPRNNSpeIRATasNet, DPRNNRawNetTasNet  # noqa: F401
from ._lib import lib  # noqa: F401
from .resample import Resample  # noqa: F401

__all__ = ['DPRNNTasNet', 'DPRNNSpeTasNet', 'DPRNNSpeIRATasNet', 'DPRNNRawNetTasNet', 'Resample', 'lib']
