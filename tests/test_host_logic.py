"""CPU tests of host-side logic that needs no kernel: gradient-buffer bookkeeping of the backward executor,
the stem/packing helpers' shapes, patch/bucket helpers."""
import pytest
import torch

from ryolo_b200 import ops
from ryolo_b200.model.backward import GradStore
from ryolo_b200.ops import Act


def _act(C=64, coff=0, Cbuf=128):
    return Act(torch.zeros(1, 2, 2, Cbuf, dtype=torch.bfloat16), C, coff)


def test_gradstore_tracks_channel_ranges_of_concat_buffers():
    G = GradStore()
    buf = torch.zeros(1, 2, 2, 128, dtype=torch.bfloat16)
    whole, lo, hi = Act(buf), Act(buf, 64, 0), Act(buf, 64, 64)
    assert not G.is_init(whole) and not G.is_init(lo)
    gv, acc = G.writable(lo)
    assert not acc and gv.coff == 0 and gv.C == 64 and gv.buf.shape == buf.shape and gv.buf is not buf
    G.mark(lo)
    assert G.is_init(lo) and not G.is_init(hi) and not G.is_init(whole)
    # a consumer that covers the whole buffer while only half is initialised: falls back to zero + accumulate
    gv, acc = G.writable(whole)
    assert acc and G.is_init(whole) and G.is_init(hi)
    # same underlying buffer -> same gradient buffer
    assert G.view(hi).buf is G.view(lo).buf


def test_gradstore_union_of_slices_covers_the_buffer():
    G = GradStore()
    buf = torch.zeros(1, 2, 2, 96, dtype=torch.bfloat16)
    for off in (0, 32, 64):
        G.mark(Act(buf, 32, off))
    assert G.is_init(Act(buf)) and G.is_init(Act(buf, 64, 16))


def test_stem_kpad_and_views():
    assert ops.stem_kpad(3) == 32 and ops.stem_kpad(6) == 128
    a = _act(32, 16, 96)
    s = a.slice(8, 16)
    assert (s.coff, s.C, s.pitch) == (24, 16, 96) and s.ptr == a.buf.data_ptr() + 2 * 24
    assert ops.out_hw(800, 800, 3, 2) == (400, 400) and ops.out_hw(25, 25, 1, 1) == (25, 25)
    assert ops.out_hw(96, 96, 6, 2) == (48, 48)


@pytest.mark.parametrize("epochs,ipe,bs,lr", [(3, 700, 8, 0.01), (2, 40, 32, 0.02), (5, 1000, 64, 0.001), (4, 300, 2, 0.01)])
def test_schedule_matches_reference_loop(epochs, ipe, bs, lr):
    """Schedule reproduces the lr / accumulate / step decisions of the reference loop (train.py:150-163,184-202,219-220)
    executed here with the real torch SGD + LambdaLR objects."""
    import numpy as np
    import torch
    from torch.optim.lr_scheduler import LambdaLR
    import ryolo_b200 as R
    lrf, warmup_prop = 0.1, 0.05
    w = torch.nn.Parameter(torch.zeros(1))
    nbs = 64
    accumulate = max(round(nbs / bs), 1)
    optimizer = torch.optim.SGD([w], lr=lr, momentum=0.937, nesterov=True)
    nw = max(int((epochs * ipe) * warmup_prop), 1000)
    lf = R.one_cycle(1, lrf, int(epochs))
    scheduler = LambdaLR(optimizer, lr_lambda=lf)
    initial_lr = optimizer.param_groups[0]['initial_lr']
    sch = R.Schedule(epochs, ipe, bs, lr, lrf, warmup_prop)
    assert sch.nw == nw
    for epoch in range(epochs):
        for batch in range(ipe):
            global_step = ipe * epoch + batch + 1
            if global_step <= nw:
                xi = [0, nw]
                accumulate = max(1, np.interp(global_step, xi, [1, nbs / bs]).round())
                optimizer.param_groups[0]['lr'] = np.interp(global_step, xi, [0.0, initial_lr * lf(epoch)])
            got_lr, got_acc, got_step = sch.batch(epoch, batch)
            assert got_acc == accumulate and got_step == (global_step % accumulate == 0)
            assert abs(got_lr - optimizer.param_groups[0]['lr']) <= 1e-12
        w.grad = torch.zeros(1)
        optimizer.step()
        scheduler.step()
        assert abs(sch.epoch_end() - optimizer.param_groups[0]['lr']) <= 1e-12


def test_decode_csl_tie_window_is_conservative():
    """csrc/decode.cu only evaluates the sigmoid of angle logits x >= cut(m), m = the largest logit, with
    cut = min(m, 14) - 2^-19 (1 + e^min(m, 14)).  A logit below the cut must be out of reach of a tie: its sigmoid must lie
    more than a handful of fp32 ulps under sigmoid(m), whatever 2-ulp error the device's expf adds on either side."""
    import numpy as np
    for m in np.concatenate((np.linspace(-30, 13.9, 200), np.linspace(14, 40, 60))):
        mc = min(m, 14.0)
        cut = mc - 2.0 ** -19 * (1.0 + np.exp(mc))
        x = np.nextafter(np.float32(cut), np.float32(-np.inf)).astype(np.float64)      # the largest excluded logit
        s_m = 1.0 / (1.0 + np.exp(-np.float64(m)))
        s_x = 1.0 / (1.0 + np.exp(-x))
        ulp = np.spacing(np.float32(s_m)).astype(np.float64)
        assert s_m - s_x > 3.0 * ulp, (m, cut, (s_m - s_x) / ulp)
