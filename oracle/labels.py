"""ORACLE (test infrastructure only) — CPU restatement of the reference's label encoding:
`xyxyxyxy2xywha` (lib/general.py:70-104), `norm_angle` (lib/general.py:7-20), `gaussian_label`
(datasets/base_dataset.py:13-31) and the row assembly of BaseDataset.__getitem__ (datasets/base_dataset.py:137-154).
PINNED: tests/golden/labels.pt holds outputs of the reference's own functions (tests/golden/make_golden_labels.py)."""
import numpy as np
import torch


def norm_angle(theta):
    theta = torch.where(theta >= np.pi / 2, theta - np.pi, theta)
    theta = torch.where(theta < -np.pi / 2, theta + np.pi, theta)
    return theta


def xyxyxyxy2xywha(boxes):
    x1, y1, x2, y2, x3, y3, x4, y4 = boxes.unbind(dim=-1)
    x = (x1 + x2 + x3 + x4) / 4
    y = (y1 + y2 + y3 + y4) / 4

    def n2(a, b):
        return torch.sqrt(a * a + b * b)

    w = (n2(x2 - x3, y2 - y3) + n2(x1 - x4, y1 - y4)) / 2
    h = (n2(x1 - x2, y1 - y2) + n2(x4 - x3, y4 - y3)) / 2
    theta = -(torch.atan2(y1 - y2, x1 - x2) + torch.atan2(y4 - y3, x4 - x3)) / 2
    swap = w >= h
    w2, h2 = torch.where(swap, h, w), torch.where(swap, w, h)
    theta = torch.where(swap, torch.where(theta > 0, theta - np.pi / 2, theta + np.pi / 2), theta)
    return torch.stack((x, y, w2, h2, norm_angle(theta)), -1)


def gaussian_label(label, num_class=180, u=0, sig=6.0):
    x = np.arange(-num_class / 2, num_class / 2)
    y_sig = np.exp(-(x - u) ** 2 / (2 * sig ** 2))
    index = int(num_class / 2 - label)
    return np.concatenate([y_sig[index:], y_sig[:index]], axis=0)


def encode_labels(polys, csl=True):
    """polys [T,10] (img, cls, 8 coords) -> [T,187] or [T,7] (base_dataset.py:137-154)."""
    if len(polys) == 0:
        return torch.zeros((0, 187 if csl else 7), dtype=torch.float32)
    rboxes = xyxyxyxy2xywha(polys[:, 2:])
    if not csl:
        return torch.cat((polys[:, :2], rboxes), -1)
    rows = [gaussian_label(rboxes[i, 4] * 180 / np.pi + 90, 180, 0, 6) for i in range(len(rboxes))]
    return torch.cat((polys[:, :2], rboxes, torch.from_numpy(np.stack(rows)).type(torch.float32)), -1)
