"""GPU: the kernel variants behind the runtime switches pass the same per-operator parity test as the defaults.
The switches are read once per process, so every variant set runs `tests/test_gpu_ops.py` for two configurations in a subprocess:

  round1   : every round-2 kernel off (round-1 fused kernels, one-tile-per-CTA stem, register-copy conv loader, per-stage
             register-copy Conv-LSTM at C = 128) and both inference approximations off (exact-erf GELU, ex2/rcp gates)
  midway   : the persistent stem with the shared-memory operand ring, per-thread stem stores, resident-weight MLP only
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    'round1': {'RVT_ATTN_V2': '0', 'RVT_MLP_V2': '0', 'RVT_LSTM_V2': '0', 'RVT_STEM_V2': '0', 'RVT_CONV_TMA': '0',
               'RVT_LSTM_CAST_DIM': '256', 'RVT_GELU_F16X2': '0', 'RVT_FAST_GATES': '0'},
    'midway': {'RVT_STEM_V2': '1', 'RVT_STEM_TMA_STORE': '0', 'RVT_MLP_V2': '1'},
}


@pytest.mark.parametrize('variant', list(VARIANTS))
def test_operator_parity_of_variant(variant):
    env = dict(os.environ)
    env.update(VARIANTS[variant])
    env['RVT_PARITY_TAG'] = '.' + variant
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_ops.py'), '-q', '-x', '-m', 'gpu',
                        '-p', 'no:cacheprovider', '-k', 'rvt_b_1mpx_bs3 or rvt_t_gen1'],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    tail = (r.stdout or '')[-1500:] + (r.stderr or '')[-500:]
    assert r.returncode == 0, f'variant {variant} {VARIANTS[variant]} failed:\n{tail}'
    assert ' passed' in r.stdout and 'failed' not in r.stdout, tail
