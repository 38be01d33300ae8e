from .chamfer_distance import ChamferDistance, ChamferDistanceFunction, ChamferIndex, ChamferIndexFunction  # noqa: F401
