"""GPU: CUDA StackedHistogram (through the C-ABI) is BIT-EXACT with the reference goldens, the
numpy oracle on seeded inputs, and size-independent properties at the BASELINE size."""
import os

import numpy as np
import pytest
import torch

from oracle import voxel_oracle as vo
from tests.golden_configs import VOXEL_CASES, make_voxel_events
from tests.helpers import GOLD

pytestmark = pytest.mark.gpu


def _construct(case_or_dims, x, y, p, t, fast, cutoff=10, validate=True):
    import rvt_b200
    bins, h, w = case_or_dims
    sh = rvt_b200.StackedHistogram(bins, h, w, cutoff, fast, validate=validate)
    dev = torch.device('cuda:0')
    out = sh.construct(*(torch.from_numpy(a).to(dev) for a in (x, y, p, t)))
    assert out.dtype == torch.uint8 and tuple(out.shape) == (2 * bins, h, w)
    return sh, out.cpu().numpy()


@pytest.mark.parametrize('name', list(VOXEL_CASES))
@pytest.mark.parametrize('fast', [True, False])
def test_voxel_matches_reference_golden(name, fast):
    case = VOXEL_CASES[name]
    gold = np.load(os.path.join(GOLD, f'voxel_{name}.npz'))['fast' if fast else 'slow']
    x, y, p, t = make_voxel_events(case)
    sh, got = _construct((case['bins'], case['height'], case['width']), x, y, p, t, fast, case.get('cutoff', 10))
    assert np.array_equal(got, gold)
    # scratch is left clean: a second call gives the same answer
    out2 = sh.construct(*(torch.from_numpy(a).cuda() for a in (x, y, p, t))).cpu().numpy()
    assert np.array_equal(out2, gold)


@pytest.mark.parametrize('seed,n,h,w,hot', [(21, 1_000_000, 360, 640, 0.0), (22, 2_000_000, 720, 1280, 0.01),
                                            (23, 31, 5, 7, 0.0), (24, 777_777, 240, 304, 0.2)])
def test_voxel_matches_oracle_seeded(seed, n, h, w, hot):
    x, y, p, t = vo.synth_events(seed, n, h, w, hot_fraction=hot)
    for fast in (True, False):
        ref = vo.stacked_histogram(x, y, p, t, 10, h, w, 10, fast)
        _, got = _construct((10, h, w), x, y, p, t, fast)
        assert np.array_equal(got, ref)


def test_voxel_full_size_properties():
    """BASELINE configs[4] size: 50M events -> 2x10x720x1280.  Checked through properties that do
    not need the (slow) CPU oracle: with cutoff 255 and no pixel above 255 counts the histogram
    sums to n; each polarity plane sums to its polarity count; splitting the stream at a bin
    boundary is additive; plus an exact oracle check on a 2M-event prefix window."""
    n, h, w, bins = 50_000_000, 720, 1280, 10
    x, y, p, t = vo.synth_events(0, n, h, w)
    _, full = _construct((bins, h, w), x, y, p, t, True, cutoff=None)
    assert int(full.max()) < 255
    assert int(full.astype(np.int64).sum()) == n
    assert int(full[bins:].astype(np.int64).sum()) == int(p.sum())
    ti = vo.time_bin_index(t, bins)
    per_bin = np.bincount(ti, minlength=bins)
    got_bins = full.reshape(2, bins, h, w).astype(np.int64).sum(axis=(0, 2, 3))
    assert np.array_equal(got_bins, per_bin)
    m = 2_000_000
    ref = vo.stacked_histogram(x[:m], y[:m], p[:m], t[:m], bins, h, w, 10, True)
    _, got = _construct((bins, h, w), x[:m], y[:m], p[:m], t[:m], True)
    assert np.array_equal(got, ref)


def test_voxel_validation_flags():
    x, y, p, t = vo.synth_events(5, 1000, 16, 16)
    bad_p = p.copy(); bad_p[10] = 2
    with pytest.raises(AssertionError):
        _construct((10, 16, 16), x, y, bad_p, t, True)
    bad_x = x.copy(); bad_x[3] = 16
    with pytest.raises(AssertionError):
        _construct((10, 16, 16), bad_x, y, p, t, True)
    with pytest.raises(AssertionError):
        _construct((10, 16, 16), x, y, p, t[::-1].copy(), True)
    # the instance recovers (scratch re-zeroed) after a rejected call
    import rvt_b200
    sh = rvt_b200.StackedHistogram(10, 16, 16, 10)
    dev = torch.device('cuda:0')
    with pytest.raises(AssertionError):
        sh.construct(*(torch.from_numpy(a).to(dev) for a in (x, y, bad_p, t)))
    ok = sh.construct(*(torch.from_numpy(a).to(dev) for a in (x, y, p, t))).cpu().numpy()
    assert np.array_equal(ok, vo.stacked_histogram(x, y, p, t, 10, 16, 16, 10))
