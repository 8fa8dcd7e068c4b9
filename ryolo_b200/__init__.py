"""Importable name of the package that lives in ``r-yolov4_b200/`` (a hyphen is not importable).

``import ryolo_b200`` exposes the reference's call surface: Yolo / ComputeCSLLoss /
ComputeKFIoULoss / post_process (+ the north_star aliases Model / compute_loss factory /
non_max_suppression); sub-modules mirror the reference layout (ryolo_b200.lib.general,
ryolo_b200.lib.loss, ryolo_b200.model.yolo, ...).
"""
import os as _os

_PKG = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "r-yolov4_b200")
__path__.insert(0, _PKG)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
