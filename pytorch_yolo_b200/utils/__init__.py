from .utils import non_max_suppression  # noqa: F401
