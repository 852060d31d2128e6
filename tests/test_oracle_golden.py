"""CPU: the oracle restatements reproduce the REFERENCE's committed golden outputs
(minted by oracle/make_golden.py from /root/reference)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import backbone_oracle as bo
from oracle import voxel_oracle as vo
from tests.golden_configs import BACKBONE_CASES, VOXEL_CASES, make_voxel_events, spec_of
from tests.helpers import check_against_golden, GOLD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('name', list(BACKBONE_CASES))
def test_backbone_oracle_matches_reference_golden(name):
    case = BACKBONE_CASES[name]
    spec = spec_of(case)
    params = bo.synth_params(spec, case['seed'], case.get('gamma_mode', 'uniform'))

    def step(x, states, mask):
        with torch.no_grad():
            return bo.backbone_forward(x, states, params, spec, mask)

    worst = check_against_golden(name, case, step, tol=2e-5)
    assert worst < 2e-5


def test_param_counts_match_paper():
    """SURVEY.md Appendix A sanity anchors: backbone params B/S/T."""
    for embed, dh, n in ((64, 32, 12_784_768), (48, 24, 7_209_312), (32, 32, 3_220_032)):
        shp = bo.param_shapes(bo.BackboneSpec(embed_dim=embed, dim_head=dh))
        assert sum(int(np.prod(s)) for s in shp.values()) == n


@pytest.fixture(scope='module')
def c_oracle():
    subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'libvoxel_oracle.so'))


@pytest.mark.parametrize('name', list(VOXEL_CASES))
@pytest.mark.parametrize('fast', [True, False])
def test_voxel_oracles_match_reference_golden(name, fast, c_oracle):
    case = VOXEL_CASES[name]
    gold = np.load(os.path.join(GOLD, f'voxel_{name}.npz'))['fast' if fast else 'slow']
    x, y, p, t = make_voxel_events(case)
    got = vo.stacked_histogram(x, y, p, t, case['bins'], case['height'], case['width'],
                               case.get('cutoff', 10), fast)
    assert got.dtype == np.uint8 and np.array_equal(got, gold)
    out = np.zeros(gold.size, np.uint8)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cut = case.get('cutoff', 10)
    rc = c_oracle.rvt_oracle_stacked_histogram(P(x), P(y), P(p), P(t), ctypes.c_int64(len(x)),
                                               case['bins'], case['height'], case['width'],
                                               0 if cut is None else cut, int(fast), P(out))
    assert rc == 0 and np.array_equal(out.reshape(gold.shape), gold)


def test_time_bin_fp32_edges():
    """t == t1 clamps to the last bin; dt == 0 -> max(dt,1) (representations.py:102-109)."""
    t = np.array([0, 4999, 5000, 49999, 50000], np.int64)
    assert vo.time_bin_index(t, 10).tolist() == [0, 0, 1, 9, 9]
    assert vo.time_bin_index(np.array([7, 7, 7], np.int64), 10).tolist() == [0, 0, 0]


@pytest.mark.parametrize('name', ['tiny_p6', 'small_dh24'])
def test_backbone_oracle_gradients_match_reference_golden(name):
    """Training-step pin: autograd through the oracle reproduces the REFERENCE's gradients
    (tests/golden/backbone_grads_*.npz, minted by oracle/make_golden.py)."""
    from tests.helpers import GRAD_CASES, grad_sub, case_inputs, train_loss
    case = BACKBONE_CASES[name]
    spec = spec_of(case)
    gold = np.load(os.path.join(GOLD, f'backbone_grads_{name}.npz'))
    params = {k: v.clone().requires_grad_(True) for k, v in bo.synth_params(spec, case['seed']).items()}
    outs, st = [], None
    for x in case_inputs(case, GRAD_CASES[name]):
        o, st = bo.backbone_forward(x.float(), st, params, spec)
        outs.append(o)
    loss = train_loss(outs, st)
    loss.backward()
    assert abs(float(loss.detach()) - float(gold['loss'])) <= 1e-5 * abs(float(gold['loss']))
    for k, p in params.items():
        ref = torch.from_numpy(gold['g.' + k])
        got = p.grad.contiguous().reshape(-1)[::grad_sub(name)]
        assert float((got - ref).norm()) <= 1e-4 * float(gold['n.' + k]) + 1e-12, k
        assert abs(float(p.grad.double().norm()) - float(gold['n.' + k])) <= 1e-4 * float(gold['n.' + k]) + 1e-12, k


def test_neighbour_oracles_match_reference_golden():
    """downsample_ev_repr, _correct_time, window indices, MixedDensityEventStack restatements vs the
    REFERENCE's outputs (tests/golden/neigh.npz, minted by oracle/make_golden.py from the reference's own code)."""
    import numpy as np
    from oracle import neighbours_oracle as no
    from oracle import voxel_oracle as vo
    from tests.golden_configs import MIXED_DENSITY_CASES, VOXEL_CASES, make_time_glitched, make_voxel_events
    from tests.helpers import GOLD
    gold = np.load(os.path.join(GOLD, 'neigh.npz'))
    c = VOXEL_CASES['uniform']
    x, y, p, t = make_voxel_events(c)
    sh = vo.stacked_histogram(x, y, p, t, c['bins'], c['height'], c['width'], 10, True)
    assert np.array_equal(no.downsample_ev_repr(sh), gold['ds_u8'])
    assert np.array_equal(no.downsample_ev_repr(gold['ds_odd_in']), gold['ds_odd'])
    assert np.array_equal(no.correct_time(make_time_glitched(31, 20000)), gold['ct'])
    ts = np.sort(np.random.RandomState(32).randint(0, 2_000_000, 50000).astype(np.int64))
    q = np.arange(50_000, 2_000_000, 50_000, dtype=np.int64)
    s_d, e_d = no.event_window_indices(ts, q, None, 50)
    s_n, _ = no.event_window_indices(ts, q, 3000, None)
    assert np.array_equal(s_d, gold['win_start_dt']) and np.array_equal(e_d, gold['win_end']) and np.array_equal(s_n, gold['win_start_n'])
    for name, c in MIXED_DENSITY_CASES.items():
        x, y, p, t = make_voxel_events(c)
        assert np.array_equal(no.mixed_density_stack(x, y, p, t, c['bins'], c['height'], c['width'], c['cutoff']), gold[name]), name


def test_detection_oracles_match_reference_golden():
    """PAFPN / YOLOX head / postprocess restatements vs the REFERENCE's outputs (tests/golden/det.npz)."""
    import numpy as np
    from oracle import detection_oracle as do
    from tests.golden_configs import DETECTION_CASES, POSTPROCESS_CASES, detection_inputs
    from tests.helpers import GOLD
    import rvt_b200.detection as det
    gold = np.load(os.path.join(GOLD, 'det.npz'))
    for name, c in DETECTION_CASES.items():
        fpn = det.YOLOPAFPN(depth=c['depth'], in_stages=(2, 3, 4), in_channels=c['in_channels'])
        head = det.YOLOXHead(num_classes=c['num_classes'], strides=(8, 16, 32), in_channels=c['in_channels'])
        sd_f = do.synth_state({k: tuple(v.shape) for k, v in fpn.state_dict().items()}, c['seed'])
        sd_h = do.synth_state({k: tuple(v.shape) for k, v in head.state_dict().items()}, c['seed'] + 1)
        fpn.load_state_dict(sd_f, strict=True)          # the mirror's state_dict keys / shapes are the reference's
        head.load_state_dict(sd_h, strict=True)
        feats = {k: torch.from_numpy(v) for k, v in detection_inputs(c).items()}
        with torch.no_grad():
            o = do.pafpn_forward(feats, sd_f, depth=c['depth'])
            oo = do.head_forward(o, sd_h)
        for i, t in enumerate(o):
            assert float((t - torch.from_numpy(gold[f'{name}_fpn{i}'])).abs().max()) < 1e-5
        ref = torch.from_numpy(gold[f'{name}_head'])
        assert float(((oo - ref).abs() / ref.abs().clamp_min(1.0)).max()) < 1e-5
    for name, c in POSTPROCESS_CASES.items():
        pred = do.synth_predictions(c['seed'], c['batch'], c['anchors'], c['num_classes'])
        for i, d in enumerate(do.postprocess(pred, c['num_classes'], c['conf'], c['nms'])):
            g = gold[f'{name}_img{i}']
            assert (d is None and len(g) == 0) or np.array_equal(d, g)
