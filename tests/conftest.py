import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'neural-ode-features_b200')
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def native_lib():
    """The in-tree shared library; built here if stale (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as entry
    entry.build()
    from node_b200 import native
    return native.lib()


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    return load


def odefunc_params(g, device='cpu'):
    """Rebuild the reference-named parameter dict from a golden file."""
    import torch
    return {k[2:]: torch.from_numpy(v).to(device) for k, v in g.items() if k.startswith('p.')}


def load_odefunc(g, device):
    """A node_b200.models.ODEfunc carrying the golden file's weights."""
    import torch
    from node_b200 import models
    f = models.ODEfunc(int(g['p.norm1.weight'].shape[0]))
    f.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith('p.')})
    return f.to(device).eval()
