from .conv_models import FiringRateEncoder, FullGaussian2d, Stacked2dCore, conv_core_gaussian_model  # noqa: F401
from .simulated_models import simulated, simulated_population  # noqa: F401


def monkey_v1(*args, **kwargs):
    raise NotImplementedError("laminr_b200: the pretrained monkey-V1 ensembles need a network download and are outside the B200 hot path "
                              "(SURVEY section 2, row 15); any nn.Module response model can be passed to InvarianceManifold instead")
