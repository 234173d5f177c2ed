import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = ('full_default', 'full_adversarial', 'ignore_noaerosol',
                'l30_minimal', 'ragged_adversarial', 'shadow_only',
                'cover_mode', 'guard_band')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    """One committed reference fixture -> (inputs dict, outputs dict)."""
    z = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
    ins = dict(
        bands=[z[f'in_band{k}'] for k in range(6)], fmask=z['in_fmask'],
        dem=z['in_dem'] if 'in_dem' in z else None,
        land=z['in_land'] if 'in_land' in z else None,
        ocean=z['in_ocean'] if 'in_ocean' in z else None,
        sun_azimuth=float(z['in_sun'][0]), sun_elevation=float(z['in_sun'][1]),
        mode=str(z['in_mode']), aerosol=bool(z['in_aerosol']),
        dem_margin=int(z['in_dem_margin']))
    outs = {k[4:]: z[k] for k in z.files if k.startswith('out_')}
    return ins, outs


@pytest.fixture(scope='session')
def golden():
    return load_golden


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope='session')
def pb():
    """The product package on a GPU box; GPU tests fail loudly without it."""
    import proteus_b200
    assert has_cuda(), 'GPU test selected but no CUDA device is visible'
    return proteus_b200
