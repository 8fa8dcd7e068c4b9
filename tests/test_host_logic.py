"""CPU tests of host-side logic that needs no kernel: gradient-buffer bookkeeping of the backward executor,
the stem/packing helpers' shapes, patch/bucket helpers."""
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
    assert ops.stem_kpad(3) == 64 and ops.stem_kpad(6) == 128
    a = _act(32, 16, 96)
    s = a.slice(8, 16)
    assert (s.coff, s.C, s.pitch) == (24, 16, 96) and s.ptr == a.buf.data_ptr() + 2 * 24
    assert ops.out_hw(800, 800, 3, 2) == (400, 400) and ops.out_hw(25, 25, 1, 1) == (25, 25)
    assert ops.out_hw(96, 96, 6, 2) == (48, 48)
