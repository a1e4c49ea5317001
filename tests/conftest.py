import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


GENO = {"0": 0, "1": 1, "2": 2, "N": None}


def decode_genotypes(s):
    return [GENO[c] for c in s]


def decode_raw(s):
    """Golden raw value: "0" is the reference's int sentinel, anything else a float.hex()."""
    return 0 if s == "0" else float.fromhex(s)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "calc_ld_golden.json")) as fh:
        return json.load(fh)
