"""Golden fixture for the label encoding, produced by EXECUTING THE REFERENCE'S OWN FUNCTIONS
(lib/general.py xyxyxyxy2xywha, datasets/base_dataset.py gaussian_label) in the build container:

    python tests/golden/make_golden_labels.py        # needs /root/reference; writes tests/golden/labels.pt
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def main():
    d2 = types.ModuleType("detectron2")
    layers = types.ModuleType("detectron2.layers")
    nms = types.ModuleType("detectron2.layers.nms")
    nms.nms_rotated = None
    sys.modules.update({"detectron2": d2, "detectron2.layers": layers, "detectron2.layers.nms": nms})
    sys.path.insert(0, REF)
    from lib.general import xyxyxyxy2xywha            # reference code
    import importlib.util                             # `datasets` also names a site-packages module: load by path
    spec = importlib.util.spec_from_file_location("ref_base_dataset", os.path.join(REF, "datasets", "base_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gaussian_label = mod.gaussian_label               # reference code

    g = torch.Generator().manual_seed(7)
    T = 600
    cx, cy = torch.rand(T, generator=g), torch.rand(T, generator=g)
    a = torch.rand(T, generator=g) * 0.2 + 0.01
    b = torch.rand(T, generator=g) * 0.2 + 0.01
    th = (torch.rand(T, generator=g) - 0.5) * 2 * np.pi
    # edge cases: axis-aligned boxes, squares, 45 degrees
    th[:40] = torch.tensor([0.0, np.pi / 2, -np.pi / 2, np.pi, np.pi / 4, -np.pi / 4, 3 * np.pi / 4, 1e-4]).repeat(5)
    b[20:40] = a[20:40]
    c, s = torch.cos(th), torch.sin(th)
    # clockwise in image coordinates (y down): (-a,-b), (a,-b), (a,b), (-a,b) rotated by th
    corners = []
    for sx, sy in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
        corners += [cx + sx * a * c - sy * b * s, cy + sx * a * s + sy * b * c]
    quad = torch.stack(corners, -1).float()
    quad[100:] += (torch.rand(T - 100, 8, generator=g) - 0.5) * 0.004      # annotations are not perfect rectangles
    cls = torch.randint(0, 16, (T,), generator=g).float()
    img = torch.randint(0, 8, (T,), generator=g).float()
    targets = torch.cat((img[:, None], cls[:, None], quad), 1)
    # datasets/base_dataset.py:139-154
    rboxes = xyxyxyxy2xywha(targets[:, 2:].clone())
    csl_labels = []
    for i in range(len(rboxes)):
        angle = rboxes[i, 4] * 180 / np.pi + 90
        csl_labels.append(gaussian_label(label=angle, num_class=180, u=0, sig=6))
    csl_labels = torch.from_numpy(np.stack(csl_labels)).type(torch.float32)
    labels_csl = torch.cat((targets[:, :2], rboxes, csl_labels), -1)
    labels_kf = torch.cat((targets[:, :2], rboxes), -1)
    torch.save(dict(targets=targets, labels_csl=labels_csl, labels_kfiou=labels_kf), os.path.join(HERE, "labels.pt"))
    print("labels.pt", tuple(labels_csl.shape))


if __name__ == "__main__":
    main()
