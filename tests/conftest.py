"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` covers the oracle against the golden vectors, the host logic and the C-ABI export list;
`-m gpu` tests are the parity tests proper and need a B200 (they go through the C-ABI library).
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def cls_key(cls):
    return "c" + "_".join(map(str, cls)) if cls else "c"


def key_cls(key):
    body = key[1:]
    return tuple(int(t) for t in body.split("_")) if body else ()


class Goldens:
    def __init__(self):
        self.index = np.load(os.path.join(GOLD, "index_goldens.npz"))
        self.ops = np.load(os.path.join(GOLD, "op_goldens.npz"))
        with open(os.path.join(GOLD, "index_goldens.json")) as f:
            self.index_meta = json.load(f)
        with open(os.path.join(GOLD, "op_goldens.json")) as f:
            self.op_meta = json.load(f)
        with open(os.path.join(GOLD, "config1.json")) as f:
            self.config1 = json.load(f)

    def packed(self, prefix):
        """{class tuple: array} for all npz entries `prefix.c...`."""
        out = {}
        for k in self.ops.files:
            if k.startswith(prefix + "."):
                out[key_cls(k[len(prefix) + 1:])] = self.ops[k]
        return out

    def cases(self, op):
        return [c for c in self.op_meta["cases"] if c["op"] == op]


@pytest.fixture(scope="session")
def goldens():
    return Goldens()
