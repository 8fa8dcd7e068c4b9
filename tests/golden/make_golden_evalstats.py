"""Golden for calculate_eval_stats (test.py:152-165) and fitness (train.py:41-44): the reference's own function
definitions, compiled from their source text (test.py / train.py cannot be imported here), applied to the
batch statistics already stored in metrics.pt.

    python tests/golden/make_golden_evalstats.py      # needs /root/reference; writes tests/golden/eval_stats.pt
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def _defs(path, names):
    tree = ast.parse(open(path).read())
    return [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]


def main():
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    ns = {"np": np, "torch": torch}
    body = _defs(os.path.join(REF, "test.py"), ("ap_per_class", "compute_ap", "calculate_eval_stats")) + \
        _defs(os.path.join(REF, "train.py"), ("fitness",))
    exec(compile(ast.Module(body=body, type_ignores=[]), "reference_eval", "exec"), ns)
    g = torch.load(os.path.join(HERE, "metrics.pt"), weights_only=False)
    stats = [np.concatenate([np.asarray(x) for x in col], 0) for col in zip(*g["stats"])]
    nt, p, r, ap50, ap, f1, ap_class, mp, mr, map50, map_ = ns["calculate_eval_stats"](stats, 3)
    fit = ns["fitness"](np.array([mp, mr, map50, map_]))            # train.py:237
    empty = ns["calculate_eval_stats"]([], 3)
    torch.save(dict(nt=nt, p=p, r=r, ap50=ap50, ap=ap, f1=f1, ap_class=ap_class, mp=mp, mr=mr, map50=map50, map=map_,
                    fitness=fit, empty=[float(v) if not hasattr(v, "__len__") else list(np.asarray(v).ravel()) for v in empty]),
               os.path.join(HERE, "eval_stats.pt"))
    print("eval_stats.pt", nt, mp, mr, map50, map_, fit)


if __name__ == "__main__":
    main()
