"""Pin oracle/model_cpu.py (functional restatement of the reference conv stack) against the model
goldens generated from the reference's own nn.Modules, and check the product's parameter tree keeps
the reference's state_dict contract (keys, order, shapes)."""
import pytest
import torch

from oracle import model_cpu
from tests.util import CFG, det_init, load, rel_err

CASES = [("yolov4", "csl", 2), ("yolov4", "kfiou", 2), ("yolov7", "csl", 16), ("yolov5", "csl", 2)]


def _product_model(ver, mode, nc):
    import ryolo_b200 as R
    return det_init(R.Yolo(nc, CFG, mode, ver))


@pytest.mark.parametrize("ver,mode,nc", CASES)
def test_state_dict_contract_and_init(ver, mode, nc):
    g = load(f"model_{ver}_{mode}_nc{nc}.pt")
    m = _product_model(ver, mode, nc)
    sd = m.state_dict()
    assert list(sd.keys()) == g["keys"]                      # same names, same ORDER (train.py:80-86 relies on it)
    for k, v in g["weight_probe"].items():
        if "running" in k or "num_batches" in k:             # probed after the reference's train-mode forward
            continue
        assert abs(float(sd[k].double().sum()) - v) <= 1e-6 * max(1.0, abs(v)), k
    assert len(sd) == {"yolov4": 648, "yolov7": 564, "yolov5": 612}[ver]


@pytest.mark.parametrize("ver,mode,nc", CASES)
def test_oracle_model_matches_reference(ver, mode, nc):
    g = load(f"model_{ver}_{mode}_nc{nc}.pt")
    sd = {k: v.clone() for k, v in _product_model(ver, mode, nc).state_dict().items()}
    torch.set_num_threads(8)
    levels, _, stats = model_cpu.forward(sd, g["img"], ver, mode, nc, train=True)
    for a, b in zip(levels, g["train_levels"]):
        assert rel_err(a, b) < 1e-3
    for k, v in g["running_after"].items():
        assert torch.allclose(stats[k], v, rtol=1e-3, atol=1e-5), k
    sd.update(stats)
    levels, infer, _ = model_cpu.forward(sd, g["img"], ver, mode, nc, train=False, anchors_cfg=CFG["anchors"],
                                         angles=CFG["angles"])
    for a, b in zip(levels, g["eval_levels"]):
        assert rel_err(a, b) < 1e-3
    assert rel_err(infer[..., :4], g["eval_infer"][..., :4]) < 1e-3


def test_unknown_mode():
    import ryolo_b200 as R
    with pytest.raises(NotImplementedError):
        R.Yolo(2, CFG, "smoothl1", "yolov4")
