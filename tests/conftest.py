import glob
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_package():
    """Imports the hyphen-named package directory caps-sa_b200/ as module caps_sa_b200."""
    import __graft_entry__

    return __graft_entry__.load_package()


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def synth():
    spec = importlib.util.spec_from_file_location("capsb_synth", os.path.join(ROOT, "caps-sa_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_cases():
    out = []
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        out.append(os.path.splitext(os.path.basename(path))[0])
    return out


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {"text": z["text"], "subproblems": int(z["subproblems"]), "idx_bytes": int(z["idx_bytes"]),
            "sa": z["sa"], "lcp": z["lcp"]}
