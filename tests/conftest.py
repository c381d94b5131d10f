import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(REPO, 'tests', 'golden')


_SD_CACHE = {}


def omni_sd(wseed, pt_eos_bias):
    """Synthetic OmniParser checkpoint, cached per (seed, bias); only one kept (576 MB each)."""
    from oracle import weights as W
    key = ('omni', wseed, pt_eos_bias)
    if key not in _SD_CACHE:
        for k in [k for k in _SD_CACHE if k[0] == 'omni']:
            del _SD_CACHE[k]
        _SD_CACHE[key] = W.omniparser_state_dict(seed=wseed, pt_eos_bias=pt_eos_bias)
    return _SD_CACHE[key]


def mgp_sd(wseed=0):
    from oracle import weights as W
    key = ('mgp', wseed)
    if key not in _SD_CACHE:
        _SD_CACHE[key] = W.mgpstr_state_dict(seed=wseed)
    return _SD_CACHE[key]
