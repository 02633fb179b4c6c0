from .conv_models import FiringRateEncoder, FullGaussian2d, Stacked2dCore, conv_core_gaussian_model  # noqa: F401
from .monkey_models import monkey_v1  # noqa: F401
from .simulated_models import simulated, simulated_population  # noqa: F401
