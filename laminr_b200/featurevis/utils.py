"""`varargin` and `Compose` (reference featurevis/utils.py:4-65)."""
import functools
import inspect


def varargin(f):
    """Lets `f` ignore keyword arguments it does not declare (ops receive `iteration=` from gradient_ascent)."""
    params = inspect.signature(f).parameters.values()
    names = {p.name for p in params}
    takes_kwargs = any(p.kind == inspect.Parameter.VAR_KEYWORD for p in params)

    @functools.wraps(f)
    def wrapper(*args, **kwargs):
        if not takes_kwargs:
            kwargs = {k: v for k, v in kwargs.items() if k in names}
        return f(*args, **kwargs)

    return wrapper


class Compose:
    """Chains operations; every operation gets the previous output plus the original keyword arguments."""

    def __init__(self, operations):
        self.operations = operations

    def __call__(self, x, **kwargs):
        out = None
        for i, op in enumerate(self.operations):
            out = op(x if i == 0 else out, **kwargs)
        return out

    def __getitem__(self, item):
        return self.operations[item]
