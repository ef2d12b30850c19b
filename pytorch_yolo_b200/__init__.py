"""pytorch_yolo_b200 -- B200-native (sm_100a) YOLO decode + NMS, a drop-in for the detection
hot path of Dipet/pytorch_yolo (YOLOLayer inference decode -> non_max_suppression)."""
from .models.yolo_layer import YOLOLayer, decode_layers, detect_layers   # noqa: F401
from .utils.utils import non_max_suppression                             # noqa: F401
from .detect import detect                                               # noqa: F401

__version__ = "0.1.0"
