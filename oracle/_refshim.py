"""TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Import helper for the *real* reference at /root/reference.  Only
``oracle/make_golden.py`` (run in the build container, never on the GPU box)
uses it, to pin the oracle restatement and to mint tests/golden/*.npz.

The reference needs ``omegaconf`` (absent from this image) solely for an
attribute-dict with ``.get`` and for ``OmegaConf.to_container``; the stub below
provides exactly that.
"""
import sys
import types

REFERENCE_ROOT = '/root/reference'


class DictConfig(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = DictConfig(v) if isinstance(v, dict) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _to_container(c, **_kw):
    if isinstance(c, dict):
        return {k: _to_container(v) for k, v in c.items()}
    if isinstance(c, (list, tuple)):
        return [_to_container(v) for v in c]
    return c


def install():
    if 'omegaconf' not in sys.modules:
        m = types.ModuleType('omegaconf')
        m.DictConfig = DictConfig

        class OmegaConf:
            to_container = staticmethod(_to_container)
            is_config = staticmethod(lambda c: isinstance(c, DictConfig))

        m.OmegaConf = OmegaConf
        m.open_dict = lambda c: c
        m.MISSING = '???'
        sys.modules['omegaconf'] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
