from .yolo_layer import YOLOLayer, decode_layers, detect_layers  # noqa: F401
