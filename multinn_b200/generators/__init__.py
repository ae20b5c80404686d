from .rnn_estimator import RnnEstimator, RnnEstimatorStateTuple  # noqa: F401
from .rnn_nade import RnnNade  # noqa: F401
from .rnn_multinade import RnnMultiNADE  # noqa: F401
