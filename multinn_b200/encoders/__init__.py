from .encoder import Encoder  # noqa: F401
from .pass_encoder import PassEncoder  # noqa: F401
