import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, {k: z[k] for k in z.files if k != "meta"}


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as fh:
        return json.load(fh)


def golden_case(name):
    """(meta, arrays, state_dict(np), input(np)) for a synthetic-weight golden case."""
    from oracle import synth
    meta, arrs = load_golden(name)
    sd = synth.synth_state_dict(meta["shapes"], meta["wseed"]) if "shapes" in meta else None
    x = None
    if "xshape" in meta:
        x = synth.synth_input(meta["xshape"], meta["xseed"], meta.get("xlo", 0.0), meta.get("xhi", 1.0))
    return meta, arrs, sd, x
