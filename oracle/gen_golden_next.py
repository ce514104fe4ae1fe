#!/usr/bin/env python
"""Golden fixtures for the rows either side of the hot path, from the UNMODIFIED reference:
add.outer / subtract.outer (symtensor/symalg.py:294-316), construction from a dense array with and without
``symmetrize`` and the symmetry check (symtensor/permcls_symtensor.py:599-618), ``todense`` (:883-887).

TEST INFRASTRUCTURE, build container only (imports /root/reference through oracle/ref_shim, like gen_golden.py).

    python oracle/gen_golden_next.py        # rewrites tests/golden/next_goldens.npz / .json
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("SYMTENSOR_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")

warnings.filterwarnings("ignore")
sys.path[:0] = [os.path.join(HERE, "ref_shim"), REF]

from symtensor import symalg, utils  # noqa: E402
from symtensor.permcls_symtensor import PermClsSymmetricTensor  # noqa: E402

SEED0 = 20261018 + 1000


def key(cls):
    return "c" + "_".join(map(str, cls)) if cls else "c"


def rand_tensor(rng, rank, dim):
    return {cls: rng.standard_normal(utils._get_permclass_size(cls, dim)) for cls in utils._perm_classes(rank) if len(cls) <= dim}


def dump(prefix, data, out):
    for cls, v in data.items():
        if np.size(v):
            out[f"{prefix}.{key(cls)}"] = np.asarray(v)


def main():
    out, cases = {}, []
    c = 0
    for opname in ("add", "subtract"):
        op = getattr(symalg, opname)
        for ra, rb, dim in [(1, 1, 5), (2, 1, 6), (2, 2, 5), (3, 2, 4), (1, 3, 4), (3, 3, 3), (4, 2, 3)]:
            rng = np.random.default_rng(SEED0 + c)
            a, b = rand_tensor(rng, ra, dim), rand_tensor(rng, rb, dim)
            A = PermClsSymmetricTensor(rank=ra, dim=dim, data={q: v.copy() for q, v in a.items()})
            B = PermClsSymmetricTensor(rank=rb, dim=dim, data={q: v.copy() for q, v in b.items()})
            res = op.outer(A, B)
            tag = f"{opname}_outer{c}"
            dump(tag + ".A", a, out)
            dump(tag + ".B", b, out)
            dump(tag + ".out", res._data, out)
            cases.append({"op": opname + "_outer", "tag": tag, "ra": ra, "rb": rb, "dim": dim})
            c += 1
    # dense -> packed (constructor), packed -> dense (todense), symmetrize=True, and the symmetry check
    for rank, dim in [(2, 5), (3, 4), (4, 3), (5, 3), (3, 6)]:
        rng = np.random.default_rng(SEED0 + c)
        a = rand_tensor(rng, rank, dim)
        A = PermClsSymmetricTensor(rank=rank, dim=dim, data={q: v.copy() for q, v in a.items()})
        dense = np.asarray(A.todense())
        B = PermClsSymmetricTensor(rank=rank, dim=dim, data=dense.copy())
        raw = rng.standard_normal((dim,) * rank)
        S = PermClsSymmetricTensor(rank=rank, dim=dim, data=raw.copy(), symmetrize=True)
        rejected = False
        try:
            PermClsSymmetricTensor(rank=rank, dim=dim, data=raw.copy())
        except ValueError as e:
            rejected = "not symmetric" in str(e)
        tag = f"dense{c}"
        dump(tag + ".A", a, out)
        out[tag + ".dense"] = dense
        dump(tag + ".repacked", B._data, out)
        out[tag + ".raw"] = raw
        dump(tag + ".symmetrized", S._data, out)
        cases.append({"op": "dense", "tag": tag, "rank": rank, "dim": dim, "raw_rejected": bool(rejected)})
        c += 1
    np.savez_compressed(os.path.join(GOLD, "next_goldens.npz"), **out)
    with open(os.path.join(GOLD, "next_goldens.json"), "w") as f:
        json.dump({"generator": "oracle/gen_golden_next.py", "seed0": SEED0, "cases": cases}, f, indent=1)
    print("written", len(cases), "cases,", sum(v.nbytes for v in out.values()), "bytes")


if __name__ == "__main__":
    main()
