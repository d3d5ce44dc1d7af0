"""ddg_b200 -- B200-native reverse-diffusion sampling hot path of
denaxen/discrete-diffusion-guidance.  Importing this package
loads libddg_b200.so; there is no CPU or PyTorch fallback."""
from . import _lib, config, dist, noise_schedule, postprocess, weights  # noqa: F401
from .classifier import Classifier  # noqa: F401
from .config import Config, compose  # noqa: F401
from .diffusion import Diffusion  # noqa: F401
from .dit import DIT, ModSpec  # noqa: F401

__all__ = ['Diffusion', 'Classifier', 'DIT', 'ModSpec', 'Config', 'compose',
           'config', 'dist', 'noise_schedule', 'postprocess', 'weights']
