import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# joint limits as the env sees them: pirl/launch/pirl.yaml:11-12 ; torso: fetch.urdf:107
LL7 = np.array([-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592])
UL7 = np.array([1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592])


# Fuzzing aid: RLITG_TEST_SEED=k shifts every integer seed handed to numpy.random.default_rng (tests and synth helpers
# alike), so that the same parity suite runs on fresh random data: `RLITG_TEST_SEED=3 python -m pytest tests -m gpu`.
# A few assertions check properties of the default data set (for example "some envs collide"); they can fail under a
# shifted seed without indicating a parity problem.
_SEED_SHIFT = int(os.environ.get("RLITG_TEST_SEED", "0"))
if _SEED_SHIFT:
    _orig_default_rng = np.random.default_rng

    def _shifted_default_rng(seed=None, *args, **kwargs):
        if isinstance(seed, (int, np.integer)):
            seed = int(seed) + 7919 * _SEED_SHIFT
        return _orig_default_rng(seed, *args, **kwargs)

    np.random.default_rng = _shifted_default_rng


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def random_configs(B, D=7, seed=0):
    rng = np.random.default_rng(seed)
    q = rng.uniform(LL7, UL7, size=(B, 7))
    if D == 8:
        q = np.concatenate([rng.uniform(0.0, 0.38615, size=(B, 1)), q], axis=1)
    return q


@pytest.fixture(scope="session")
def fk_pairs():
    return np.load(os.path.join(GOLDEN, "fk_pairs.npz"))


@pytest.fixture(scope="session")
def init_traj():
    return np.load(os.path.join(GOLDEN, "init_traj.npz"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
