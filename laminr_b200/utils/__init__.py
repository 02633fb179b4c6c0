from .mei import get_mei_dict  # noqa: F401
