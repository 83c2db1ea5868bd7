"""bijax_b200 -- B200-native (sm_100a) engine for bijax's ADVI hot path.

    from bijax_b200 import ADVI, distributions as tfd, bijectors as tfb, likelihoods
    from bijax_b200.utils import train

Importing the package does not touch CUDA; the first call into the engine loads libbjx.so (hand-written CUDA behind the
C ABI of include/bjx.h) and raises if it is missing -- there is no CPU fallback.
"""
from . import bijectors, distributions, likelihoods, optim, random  # noqa: F401
from .advi import ADVI  # noqa: F401
from .dataset import PreparedX  # noqa: F401
from .minibatch import MinibatchLoss  # noqa: F401
from .point import MAP, MCMC  # noqa: F401
from .utils import train, train_fn, value_and_grad  # noqa: F401

__version__ = "0.1.0"
