"""Mirror of the slice of the vendored `featurevis` package the pipeline uses (reference src/featurevis/core.py, ops.py,
utils.py): `gradient_ascent`, `ChangeNorm`, `ChangeStats`, `ClipRange`, `Compose`."""
from .core import gradient_ascent  # noqa: F401
from . import ops  # noqa: F401
from .utils import Compose, varargin  # noqa: F401
