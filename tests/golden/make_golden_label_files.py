"""Golden fixture for the label-file half of the label pipeline, produced by EXECUTING THE REFERENCE'S OWN dataset code
(datasets/DOTA_dataset.py, UCASAOD_dataset.py, base_dataset.py) on synthetic label files in the build container:

    python tests/golden/make_golden_label_files.py      # needs /root/reference; writes tests/golden/label_files.pt
"""
import importlib.util
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _load(name, path, package=None):
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    for n in ("detectron2", "detectron2.layers", "detectron2.layers.nms"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["detectron2.layers.nms"].nms_rotated = None
    sys.path.insert(0, REF)
    # `datasets` also names a site-packages module: register the reference package by path
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["datasets"] = pkg
    base = _load("datasets.base_dataset", os.path.join(REF, "datasets", "base_dataset.py"))
    dota = _load("datasets.DOTA_dataset", os.path.join(REF, "datasets", "DOTA_dataset.py"))
    ucas = _load("datasets.UCASAOD_dataset", os.path.join(REF, "datasets", "UCASAOD_dataset.py"))

    g = torch.Generator().manual_seed(11)
    names_dota = ["plane", "ship", "storage tank", "large vehicle"]
    names_ucas = ["car", "plane"]
    files = {}
    lines = []
    for _ in range(37):
        c = (torch.rand(8, generator=g) * 1000).tolist()
        cls = names_dota[int(torch.randint(0, 4, (1,), generator=g))].replace(" ", "-")
        lines.append(" ".join(f"{v:.1f}" for v in c) + f" {cls} {int(torch.randint(0, 2, (1,), generator=g))}\n")
    files["dota"] = "".join(lines)
    lines = []
    for _ in range(23):
        c = (torch.rand(8, generator=g) * 600).tolist()
        cls = names_ucas[int(torch.randint(0, 2, (1,), generator=g))]
        lines.append(cls + "\t" + "\t".join(f"{v:.3f}" for v in c) + "\t12.5\t1.0\t2.0\t3.0\t4.0\n")
    files["ucas"] = "".join(lines)
    files["empty"] = ""
    out = dict(files=files, names_dota=names_dota, names_ucas=names_ucas, cases=[])
    hyp = dict(mosaic=0.0)
    with tempfile.TemporaryDirectory() as td:
        for fmt, cls_names, Klass in (("dota", names_dota, dota.DOTADataset), ("ucas", names_ucas, ucas.UCASAODDataset),
                                      ("empty_dota", names_dota, dota.DOTADataset)):
            path = os.path.join(td, fmt + ".txt")
            open(path, "w").write(files[fmt.split("_")[0]])
            ds = Klass.__new__(Klass)
            base.BaseDataset.__init__(ds, hyp, 416, False, True, False)
            ds.category = {n.replace(" ", "-"): i for i, n in enumerate(cls_names)}
            ds.img_files, ds.label_files = ["x.png"], [path]
            polys, labels = ds.load_files(path)
            case = dict(fmt=fmt, polys=polys.clone(), labels=torch.as_tensor(labels).clone())
            for boarder in (None, (0, 300, 0, 280)):
                t = ds.load_target(0, (7, 13), (1000, 800), (416, 333), boarder)
                case["target_" + ("b" if boarder else "nb")] = t.clone()
            t = ds.load_target(0, (7, 13), (1000, 800), (416, 333))
            t = base.BaseDataset.filtering(t, (0, 416, 0, 416))
            t = base.BaseDataset.normalize(t, (416, 416))
            case["final"] = t.clone()
            out["cases"].append(case)
    # collate_fn on two images' targets
    a, b = out["cases"][0]["final"].clone(), out["cases"][1]["final"].clone()
    dsx = dota.DOTADataset.__new__(dota.DOTADataset)
    _, _, cat = dsx.collate_fn([("p0", torch.zeros(3, 4, 4), a), ("p1", torch.zeros(3, 4, 4), b)])
    out["collated"] = cat.clone()
    torch.save(out, os.path.join(HERE, "label_files.pt"))
    print("label_files.pt", [tuple(c["final"].shape) for c in out["cases"]], tuple(cat.shape))


if __name__ == "__main__":
    main()
