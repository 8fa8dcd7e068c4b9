"""Public call surface: the three Python callables the reference's train.py /
test.py / detect.py use, under their real names plus the aliases BASELINE.json's north_star uses."""
from ._lib import RyoloError, SO_PATH, lib
from .lib.general import (encode_labels, xyxyxyxy2xywha, nms_rotated, non_max_suppression, norm_angle,
                          pairwise_iou_rotated, post_process, post_process_device)
from .lib.loss import ComputeCSLLoss, ComputeKFIoULoss, KFLoss
from .lib.metrics import ap_per_class, calculate_eval_stats, compute_ap, fitness, get_batch_statistics, match_batch
from .model.yololayer import YoloCSLLayer, YoloKFIoULayer
from .schedule import Schedule, one_cycle
from .train_step import TrainStep

try:  # the conv stack is built on top of the kernels above
    from .model.yolo import Yolo
    Model = Yolo
except ImportError:  # pragma: no cover - during bring-up only
    Yolo = Model = None


def compute_loss(model, hyp, mode="csl"):
    """north_star alias: returns the loss callable train.py binds to `compute_loss` (train.py:140,143)."""
    return ComputeCSLLoss(model, hyp) if mode == "csl" else ComputeKFIoULoss(model, hyp)


__all__ = ["Yolo", "Model", "ComputeCSLLoss", "ComputeKFIoULoss", "KFLoss", "compute_loss", "post_process",
           "post_process_device", "non_max_suppression", "nms_rotated", "pairwise_iou_rotated", "norm_angle",
           "encode_labels", "xyxyxyxy2xywha", "YoloCSLLayer", "YoloKFIoULayer", "TrainStep", "Schedule", "one_cycle",
           "get_batch_statistics", "match_batch", "ap_per_class", "compute_ap", "calculate_eval_stats", "fitness", "RyoloError", "SO_PATH",
           "lib"]
