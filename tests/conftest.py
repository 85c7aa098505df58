import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODEL_DIR = os.environ.get("SS_MODEL_DIR", "/tmp/ss_models")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def model_path(shape: str, family: str, seed: int = 0) -> str:
    from speaksense_b200 import synth
    p = os.path.join(MODEL_DIR, "ggml-%s-%s-s%d.bin" % (shape, family, seed))
    synth.ensure_model(p, shape=shape, family=family, seed=seed)
    return p


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def tiny_en_peaked():
    return model_path("tiny.en", "peaked", 0)


@pytest.fixture(scope="session")
def micro_v3_random():
    return model_path("micro-v3", "random", 3)


@pytest.fixture(scope="session")
def micro_v3_peaked():
    return model_path("micro-v3", "peaked", 1)


@pytest.fixture(scope="session")
def audio30():
    from speaksense_b200 import synth
    return synth.synth_audio(seed=1234)
