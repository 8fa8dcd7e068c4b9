"""Label files -> polygon targets, the host half of the label pipeline.  Mirrors the reference's
datasets: file formats of `DOTADataset.load_files` (datasets/DOTA_dataset.py:17-49: space separated, 8 vertex
coordinates then the class name) and `UCASAODDataset.load_files` (datasets/UCASAOD_dataset.py:20-50: tab separated, class
name then 8 coordinates), `BaseDataset.load_target` (datasets/base_dataset.py:193-237), `filtering` (:343-354),
`normalize` (:357-363) and `collate_fn` (:161-167).  The result feeds `encode_labels` (csrc/labels.cu) on the device."""
import os

import torch


def category_map(class_names):
    """name -> index with spaces turned into dashes (DOTA_dataset.py:13-15)."""
    return {name.replace(" ", "-"): i for i, name in enumerate(class_names)}


def load_label_file(label_path, category, fmt):
    """(polys fp32 [T,8], labels int64 [T]) of one label file; fmt is 'dota' or 'ucas'."""
    if fmt not in ("dota", "ucas"):
        raise NotImplementedError("The specified label format is not implemented.")
    coords, labels = [], []
    for line in open(label_path, "r").readlines():
        if fmt == "dota":
            p = line.split(" ")
            coords.append([float(v) for v in p[0:8]])
            labels.append(category[p[8]])
        else:
            p = line.split("\t")
            coords.append([float(v) for v in p[1:9]])
            labels.append(category[p[0]])
    if not labels:
        return torch.zeros((0, 8), dtype=torch.float32), []
    return torch.tensor(coords).type(torch.float32), torch.tensor(labels)


def filtering(targets, boarder):
    """Drop objects whose centre lies outside (x1, x2, y1, y2) (base_dataset.py:343-354)."""
    x1, x2, y1, y2 = boarder
    x = torch.mean(targets[:, [2, 4, 6, 8]], dim=1)
    y = torch.mean(targets[:, [3, 5, 7, 9]], dim=1)
    return targets[(x > x1) & (x < x2) & (y > y1) & (y < y2)]


def normalize(targets, img_size):
    """Vertex coordinates -> [0, 1] in place (base_dataset.py:357-363)."""
    height, width = img_size
    targets[:, [2, 4, 6, 8]] /= width
    targets[:, [3, 5, 7, 9]] /= height
    return targets


def load_target(label_path, category, fmt, pad, img_size0, img_size, normalized_labels=False, boarder=None):
    """[T,10] rows (0, class, x1, y1, ..., x4, y4) in padded-image pixels (base_dataset.py:193-237)."""
    label_path = label_path.rstrip()
    assert os.path.exists(label_path), "Label file {} not found".format(label_path)
    polys, labels = load_label_file(label_path, category, fmt)
    if not len(labels):
        return torch.zeros((0, 10))
    if not normalized_labels:
        h0, w0 = img_size0
        polys[:, [0, 2, 4, 6]] /= w0
        polys[:, [1, 3, 5, 7]] /= h0
    h_, w_ = img_size
    polys[:, [0, 2, 4, 6]] *= w_
    polys[:, [1, 3, 5, 7]] *= h_
    targets = torch.zeros((len(labels), 10))
    targets[:, 1:] = torch.cat((labels.unsqueeze(-1), polys), -1)
    if boarder is not None:
        targets = filtering(targets, boarder)
    targets[:, [2, 4, 6, 8]] += pad[1]
    targets[:, [3, 5, 7, 9]] += pad[0]
    return targets


def collate_targets(per_image_targets):
    """Stamp the sample index into column 0 and concatenate (collate_fn, base_dataset.py:161-167)."""
    for i, boxes in enumerate(per_image_targets):
        boxes[:, 0] = i
    return torch.cat(list(per_image_targets), 0)
