import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fixtures():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def genomes(fixtures):
    """Named genome inputs: the reference's evolved/theory genomes plus seeded random ones (genes 0..10)."""
    rng = np.random.default_rng(0x45766F53)
    return {
        "agg_evolved": np.asarray(fixtures["agg_genome_evolved"], dtype=np.uint8).reshape(-1),
        "agg_theory": np.asarray(fixtures["agg_genome_theory"], dtype=np.uint8).reshape(-1),
        "sep_evolved": np.asarray(fixtures["sep_genome_evolved"], dtype=np.uint8).reshape(-1),
        "sep_theory": np.asarray(fixtures["sep_genome_theory"], dtype=np.uint8).reshape(-1),
        "agg_random": rng.integers(0, 11, size=48, dtype=np.uint8),
        "sep_random": rng.integers(0, 11, size=600, dtype=np.uint8),
        "coat_random": rng.integers(0, 11, size=600, dtype=np.uint8),
        "loco_random": rng.integers(0, 11, size=144, dtype=np.uint8),
    }


@pytest.fixture(scope="session")
def ctx():
    import evosops_b200 as E
    c = E.Context()
    yield c
    c.close()
