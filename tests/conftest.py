import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cs():
    import csb200
    return csb200


@pytest.fixture(scope="session")
def oracle():
    import cellscopes_oracle
    return cellscopes_oracle


@pytest.fixture(scope="session")
def ctx(cs):
    """A CUDA context.  GPU tests must run on the CUDA path: a missing library or device is an error,
    never a skip or a fallback."""
    return cs.Context(0)


def make_counts(cs, oracle, n_genes, n_cells, density=0.05, seed=1234, **kw):
    """Synthetic counts (NumPy twin of the device generator), filtered like scRNAObject does."""
    p = cs.synth.make_params(n_genes, n_cells, density=density, seed=seed, **kw)
    m = cs.synth.sample_counts(p)
    m, _, _ = oracle.subset_matrix(m)
    return m


@pytest.fixture(scope="session")
def small_counts(cs, oracle):
    return make_counts(cs, oracle, 1500, 1200, density=0.08, seed=7)


@pytest.fixture(scope="session")
def medium_counts(cs, oracle):
    return make_counts(cs, oracle, 4000, 6000, density=0.05, seed=1234)
