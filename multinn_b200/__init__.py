"""multinn_b200 -- B200-native (sm_100a) training and sampling hot path of ilya16/MultINN behind the
reference's Encoder / Generator / mode interface. Compute happens only in libmultinn_sm100.so."""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)
from .multinn import MultINN  # noqa: F401

__all__ = ['MultINN']
