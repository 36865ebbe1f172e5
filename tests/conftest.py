import importlib
import json
import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pkg(sub: str = ""):
    """Import zlib-streams-ts_b200[.sub] (the directory name is not a Python identifier)."""
    name = "zlib-streams-ts_b200" + (("." + sub) if sub else "")
    return importlib.import_module(name)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def fixtures64():
    meta = json.load(open(os.path.join(GOLDEN, "deflate64_fixtures.json")))
    blob = open(os.path.join(GOLDEN, "deflate64_fixtures.bin"), "rb").read()
    out = []
    for f in meta["fixtures"]:
        out.append(dict(f, data=blob[f["offset"]: f["offset"] + f["length"]]))
    return out


def make_text(n: int, seed: int = 1) -> bytes:
    return pkg("corpus").text_numpy(n, seed).tobytes()


def make_mixed(n: int, seed: int = 2) -> bytes:
    return pkg("corpus").mixed_numpy(n, seed, tile=max(4096, min(4 << 20, n // 3 or 4096))).tobytes()


def rand_bytes(n: int, seed: int = 3) -> bytes:
    return random.Random(seed).randbytes(n)


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return pkg("batch").default_context(0)
