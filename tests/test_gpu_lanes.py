"""Several fields in flight on one GPU (lossycomp.lanes): whatever the number of lanes, every field's blob is the
blob of a separate compress() call, in input order; a lane is an independent engine, so two threads compressing
different fields at the same time must not see each other's buffers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def _pin(torch, f):
    p = torch.empty(f.shape, dtype=torch.float32, pin_memory=True)
    p.copy_(torch.from_numpy(f))
    return p.numpy()


@pytest.mark.parametrize("lanes", [1, 2, 3])
def test_stream_blobs_equal_single_calls_for_any_lane_count(torch, lanes):
    from lossycomp.compress import compress, compress_stream
    from oracle import lossycomp_oracle as O
    # seven fields of three different shapes and bounds regimes (dense and sparse outliers)
    fields = [O.synthetic_field(16 + 8 * (s % 2), 64 + 16 * (s % 3), 80, seed=s) for s in range(7)]
    for thr in (0.05, 1.0):
        want = [compress(f, thr) for f in fields]
        got = list(compress_stream([_pin(torch, f) for f in fields], thr, lanes=lanes))
        assert got == want
        assert list(compress_stream(fields, thr, lanes=lanes)) == want          # pageable inputs


def test_lanes_are_independent_engines(torch):
    """Two lanes compress two different fields at the same time, many times over: every result equals the
    single-lane one (shared scratch buffers or streams would show up as differing blobs)."""
    from lossycomp.compress import compress_device, dumps_slots
    from lossycomp.lanes import LaneSet
    from oracle import lossycomp_oracle as O
    ls = LaneSet.get(None, "fp16", 0, 2)
    assert ls.lanes[0].eng is not ls.lanes[1].eng and ls.lanes[0].stream != ls.lanes[1].stream
    xs = [torch.from_numpy(O.synthetic_field(32, 96, 144, seed=s)).cuda() for s in (11, 12)]
    torch.cuda.synchronize()
    want = [dumps_slots(compress_device(ls.lanes[0].eng, x, 0.1)[0]) for x in xs]
    futs = [ls.lanes[i % 2].submit(lambda lane, i=i: dumps_slots(compress_device(lane.eng, xs[i % 2], 0.1)[0]))
            for i in range(40)]
    for i, f in enumerate(futs):
        assert f.result() == want[i % 2], i


def test_error_in_one_field_surfaces_and_later_fields_still_work(torch):
    from lossycomp.compress import compress, compress_stream
    from oracle import lossycomp_oracle as O
    good = O.synthetic_field(16, 64, 80, seed=3)
    bad = np.zeros((16, 64, 80, 3), dtype=np.float32)            # wrong channel count (compress.py:55)
    it = compress_stream([good, bad, good], 0.1, lanes=2)
    with pytest.raises(AssertionError):
        list(it)
    assert list(compress_stream([good, good], 0.1, lanes=2)) == [compress(good, 0.1)] * 2


@pytest.mark.parametrize("lanes", [1, 2])
def test_decompress_stream_equals_decompress(torch, lanes):
    from lossycomp.compress import compress
    from lossycomp.decompress import decompress, decompress_stream
    from oracle import lossycomp_oracle as O
    fields = [O.synthetic_field(16 + 8 * (s % 2), 64 + 16 * (s % 3), 80, seed=20 + s) for s in range(5)]
    blobs = [compress(f, 0.1) for f in fields]
    want = [decompress(b) for b in blobs]
    got = list(decompress_stream(blobs, lanes=lanes))
    assert len(got) == len(want)
    for f, a, b in zip(fields, got, want):
        assert a.shape == b.shape == f.shape[:3] + (1,) and np.array_equal(a, b)
        assert np.abs(f[..., :1].astype(np.float64) - a).max() <= 0.1
    outs = [np.empty(w.shape, dtype=np.float32) for w in want]
    res = list(decompress_stream(blobs, lanes=lanes, outs=outs))
    assert all(r is o or r.base is o for r, o in zip(res, outs)) and all(np.array_equal(o, w) for o, w in zip(outs, want))
    assert list(decompress_stream([], lanes=lanes)) == []


def test_two_lanes_under_load_reproduce_the_single_lane_results(torch):
    """Regression test of the stage-ring parity aliasing in sweep_acc_kernel: with the banks of a CTA sharing one ring
    of input stages, a load delayed by the OTHER lane's memory traffic let an issuer warp pass a parity wait one phase
    early and multiply stale data (a few tiles of the last chunks came out a few half-precision ulps off, roughly once
    in 600 fields, only with two fields in flight).  60 compressions and 60 decompressions of a 24 x 336 x 720 field,
    two in flight: latent, reconstruction, mask and codes must equal the single-lane ones every time."""
    from lossycomp import _native as N
    from lossycomp.compress import compress_device
    from lossycomp.decompress import decompress_device
    from lossycomp.lanes import LaneSet
    ls = LaneSet.get(None, "fp16", 0, 2)
    eng = ls.lanes[0].eng
    T, H, W = 24, 336, 720
    x = torch.empty((T, H, W, 2), dtype=torch.float32, device="cuda")
    N.check(eng.lib.lc_synth_field(x.data_ptr(), T, H, W, 0, 0, 99, None), "synth")
    torch.cuda.synchronize()
    slots, rinfo = compress_device(eng, x, 0.1, return_parts=True)
    want = {k: rinfo[k].clone() for k in ("latent", "recon_std", "mask", "q")}
    ref = decompress_device(eng, slots).clone()
    torch.cuda.synchronize()

    def cjob(lane):
        s, info = compress_device(lane.eng, x, 0.1, return_parts=True)
        bad = [k for k in want if info[k].shape != want[k].shape or not bool(torch.equal(info[k], want[k]))]
        return bad + [i for i in (0, 6, 7) if s[i] != slots[i]]

    def djob(lane):
        out = decompress_device(lane.eng, slots)
        ok = bool(torch.equal(out, ref))
        torch.cuda.current_stream(lane.eng.dev).synchronize()
        return [] if ok else ["xhat"]

    jobs = []
    for i in range(60):
        jobs += [ls.submit(cjob), ls.submit(djob)]
    for i, j in enumerate(jobs):
        assert j.result() == [], (i, j.result())
