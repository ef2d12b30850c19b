from .utils import non_max_suppression, scale_coords, dict_from_results  # noqa: F401
