"""Golden fixture for the box-geometry helpers the reference's detect / plot path uses (lib/general.py xywh2xyxy,
xywha2xyxyxyxy; lib/plot.py rescale_boxes), produced by EXECUTING THE REFERENCE'S OWN functions:

    python tests/golden/make_golden_geometry.py      # needs /root/reference; writes tests/golden/geometry.pt
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    for n in ("detectron2", "detectron2.layers", "detectron2.layers.nms"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["detectron2.layers.nms"].nms_rotated = None
    sys.path.insert(0, REF)
    from lib.general import xywh2xyxy, xywha2xyxyxyxy      # reference code
    from lib.plot import rescale_boxes                     # reference code
    g = torch.Generator().manual_seed(5)
    N = 300
    boxes = torch.cat((torch.rand(N, 2, generator=g) * 800, torch.rand(N, 2, generator=g) * 200 + 2,
                       (torch.rand(N, 1, generator=g) - 0.5) * np.pi), 1)
    boxes[:6, 4] = torch.tensor([0.0, np.pi / 2 - 1e-4, -np.pi / 2, np.pi / 4, -np.pi / 4, 1e-3])
    out = dict(boxes=boxes, corners=xywha2xyxyxyxy(boxes.clone()), xyxy=xywh2xyxy(boxes[:, :4].clone()), rescale=[])
    for dim, shape in ((416, (1000, 800)), (800, (600, 1024)), (608, (512, 512))):
        b = torch.cat((boxes.clone() * (dim / 800.0), torch.rand(N, 2, generator=g)), 1)
        out["rescale"].append(dict(dim=dim, shape=shape, inp=b.clone(), out=rescale_boxes(b.clone(), dim, shape)))
    torch.save(out, os.path.join(HERE, "geometry.pt"))
    print("geometry.pt", tuple(out["corners"].shape))


if __name__ == "__main__":
    main()
