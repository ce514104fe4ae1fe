#!/usr/bin/env python
"""Generate the golden fixtures in ``tests/golden/`` from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container, where ``/root/reference`` is mounted: the
reference package is imported from there through ``oracle/ref_shim`` (stand-ins for its four
uninstallable dependencies).  The fixtures travel with the repo; the reference does not.

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz, *.json

Seeds are fixed (20261018 + case id) and inputs follow SURVEY.md 8(d): packed values ~ U[0.5,1.5) (headline
distribution) and N(0,1) (stress distribution).
"""
from __future__ import annotations

import itertools
import json
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("SYMTENSOR_REFERENCE", "/root/reference")
GOLD = os.path.join(REPO, "tests", "golden")

warnings.filterwarnings("ignore")
sys.path[:0] = [os.path.join(HERE, "ref_shim"), REF]

import symtensor  # noqa: E402
from symtensor import symalg, utils  # noqa: E402
from symtensor.permcls_symtensor import (PermClsSymmetricTensor, get_index_representative,  # noqa: E402
                                         σindex_iter)
import symtensor.flat_symtensor as fs  # noqa: E402
import symtensor.torch_symtensor as ts  # noqa: E402

SEED0 = 20261018


def key(cls):
    return "c" + "_".join(map(str, cls)) if cls else "c"


def rand_tensor(rng, rank, dim, dist):
    data = {}
    for cls in utils._perm_classes(rank):
        if len(cls) > dim:
            continue
        n = utils._get_permclass_size(cls, dim)
        data[cls] = rng.uniform(0.5, 1.5, n) if dist == "pos" else rng.standard_normal(n)
    return data


def dump_packed(prefix, data, out):
    for cls, v in data.items():
        out[f"{prefix}.{key(cls)}"] = np.asarray(v)


def gen_index_goldens():
    """Integer goldens: class order, sizes, multiplicities, storage order, representatives, flat order."""
    out = {}
    meta = {"perm_classes": {}, "sizes": {}, "multiplicities": {}, "representatives": {}, "permclass_of": {}}
    for r in range(0, 9):
        meta["perm_classes"][str(r)] = [list(c) for c in utils._perm_classes(r)]
        meta["multiplicities"][str(r)] = [int(utils.get_permclass_multiplicity(c)) for c in utils._perm_classes(r)]
        for d in (1, 2, 3, 5, 10, 40, 50, 64, 200, 400, 1000):
            meta["sizes"][f"{r},{d}"] = [int(utils._get_permclass_size(c, d)) if len(c) <= d else 0
                                          for c in utils._perm_classes(r)]
    # full storage order for small (rank, dim): every class
    for r, d in [(1, 4), (2, 3), (2, 5), (3, 3), (3, 4), (3, 6), (4, 4), (4, 6), (5, 5), (6, 6), (6, 7), (7, 4), (8, 5)]:
        for cls in utils._perm_classes(r):
            idx = np.array(list(σindex_iter(cls, d)), dtype=np.int16).reshape(-1, r)
            out[f"sigma.r{r}.d{d}.{key(cls)}"] = idx
    # a mid-size class: a strided sample + checksum of the (1,1,1,1) and (2,1,1) classes at d=50 (config C1)
    for cls in [(2, 1, 1), (1, 1, 1, 1), (3, 1), (2, 2)]:
        idx = np.array(list(σindex_iter(cls, 50)), dtype=np.int16)
        out[f"sigma_sample.r4.d50.{key(cls)}.pos"] = np.arange(0, idx.shape[0], 997, dtype=np.int64)
        out[f"sigma_sample.r4.d50.{key(cls)}.idx"] = idx[::997]
        w = (np.arange(idx.shape[0], dtype=np.int64)[:, None] + 1) * (idx.astype(np.int64) + 1)
        out[f"sigma_sample.r4.d50.{key(cls)}.checksum"] = np.array([int(w.sum() % (2**61 - 1))], dtype=np.int64)
    rng = np.random.default_rng(SEED0)
    reps = []
    for r in range(1, 9):
        for _ in range(12):
            idx = tuple(int(v) for v in rng.integers(0, 7, r))
            reps.append([list(idx), list(get_index_representative(idx)), list(utils._get_permclass(idx))])
    reps.append([[5, 4, 3, 3, 2, 1], list(get_index_representative((5, 4, 3, 3, 2, 1))),
                 list(utils._get_permclass((5, 4, 3, 3, 2, 1)))])
    meta["representatives"] = reps
    # flat order (reference: index_of_multicombination / permutation_count)
    for r, d in [(1, 5), (2, 4), (3, 5), (4, 6), (5, 4), (6, 5)]:
        combos = list(itertools.combinations_with_replacement(range(d), r))
        out[f"flat.r{r}.d{d}.idx"] = np.array(combos, dtype=np.int16)
        out[f"flat.r{r}.d{d}.rank"] = np.array([fs.index_of_multicombination(d, c) for c in combos], dtype=np.int64)
        out[f"flat.r{r}.d{d}.mult"] = np.array([fs.permutation_count(c) for c in combos], dtype=np.int64)
    for r, d, sample in [(4, 200, 4001), (6, 64, 50021), (8, 40, 100003), (3, 1000, 70001)]:
        n = fs.multicomb(d, r)
        rr = np.random.default_rng(SEED0 + r)
        idx = np.sort(rr.integers(0, d, size=(300, r)), axis=1)
        out[f"flat_big.r{r}.d{d}.idx"] = idx.astype(np.int16)
        out[f"flat_big.r{r}.d{d}.rank"] = np.array([fs.index_of_multicombination(d, tuple(int(v) for v in c))
                                                    for c in idx], dtype=np.int64)
        meta.setdefault("flat_sizes", {})[f"{r},{d}"] = int(n)
    np.savez_compressed(os.path.join(GOLD, "index_goldens.npz"), **out)
    with open(os.path.join(GOLD, "index_goldens.json"), "w") as f:
        json.dump(meta, f, indent=0, sort_keys=True)


def gen_op_goldens():
    """Floating goldens: the four ops run by the reference itself (NumPy PermCls backend, fp64)."""
    out = {}
    cases = []
    cid = 0

    def new_rng():
        nonlocal cid
        cid += 1
        return np.random.default_rng(SEED0 + cid), cid

    # --- contract_all_indices_with_vector
    for rank, dim, dist in [(1, 5, "pos"), (2, 6, "pos"), (3, 3, "normal"), (3, 7, "pos"), (4, 5, "pos"), (4, 9, "normal"),
                            (5, 4, "pos"), (6, 4, "pos"), (4, 12, "pos")]:
        rng, c = new_rng()
        data = rand_tensor(rng, rank, dim, dist)
        x = (rng.uniform(0.5, 1.5, dim) / np.sqrt(dim)) if dist == "pos" else rng.standard_normal(dim)
        A = PermClsSymmetricTensor(rank=rank, dim=dim, data={k: v.copy() for k, v in data.items()})
        t0 = time.perf_counter()
        res = symalg.contract_all_indices_with_vector(A, x)
        dt = time.perf_counter() - t0
        assert res.rank == 0
        tag = f"vec{c}"
        dump_packed(tag + ".A", data, out)
        out[tag + ".x"] = x
        out[tag + ".out"] = np.asarray(res._data[()]).reshape(())
        cases.append({"op": "vec", "tag": tag, "rank": rank, "dim": dim, "dist": dist, "ref_seconds": dt})
    # scalar-compressed classes (0-d entries), as in testing/api.py:59-67
    rng, c = new_rng()
    A = PermClsSymmetricTensor(rank=4, dim=6)
    A["iiii"] = rng.standard_normal(6)
    A["iijj"] = rng.standard_normal(15)
    A["iijk"] = 0.75
    x = rng.standard_normal(6)
    res = symalg.contract_all_indices_with_vector(A, x)
    tag = f"vec{c}"
    dump_packed(tag + ".A", A._data, out)
    out[tag + ".x"] = x
    out[tag + ".out"] = np.asarray(res._data[()]).reshape(())
    cases.append({"op": "vec", "tag": tag, "rank": 4, "dim": 6, "dist": "scalar-classes"})

    # --- contract_all_indices_with_matrix
    for rank, dim, dist in [(2, 5, "pos"), (3, 3, "normal"), (3, 6, "pos"), (4, 4, "normal"), (4, 6, "pos"), (5, 3, "pos"),
                            (6, 3, "pos"), (6, 4, "normal")]:
        rng, c = new_rng()
        data = rand_tensor(rng, rank, dim, dist)
        W = (rng.uniform(0.5, 1.5, (dim, dim)) / dim) if dist == "pos" else rng.standard_normal((dim, dim))
        A = PermClsSymmetricTensor(rank=rank, dim=dim, data={k: v.copy() for k, v in data.items()})
        t0 = time.perf_counter()
        res = symalg.contract_all_indices_with_matrix(A, W)
        dt = time.perf_counter() - t0
        tag = f"mat{c}"
        dump_packed(tag + ".A", data, out)
        out[tag + ".W"] = W
        dump_packed(tag + ".out", {k: v for k, v in res._data.items() if np.size(v)}, out)
        cases.append({"op": "mat", "tag": tag, "rank": rank, "dim": dim, "dist": dist, "ref_seconds": dt})

    # --- tensordot
    for ra, rb, k, dim, dist in [(3, 3, 1, 5, "pos"), (3, 3, 2, 6, "normal"), (4, 2, 1, 5, "pos"), (3, 1, 1, 8, "pos"),
                                 (2, 2, 2, 7, "normal"), (3, 3, 3, 4, "pos"), (4, 3, 2, 4, "pos"), (2, 3, 0, 4, "normal"),
                                 (3, 3, 1, 10, "pos")]:
        rng, c = new_rng()
        a, b = rand_tensor(rng, ra, dim, dist), rand_tensor(rng, rb, dim, dist)
        A = PermClsSymmetricTensor(rank=ra, dim=dim, data={q: v.copy() for q, v in a.items()})
        B = PermClsSymmetricTensor(rank=rb, dim=dim, data={q: v.copy() for q, v in b.items()})
        t0 = time.perf_counter()
        res = symalg.tensordot(A, B, axes=k)
        dt = time.perf_counter() - t0
        tag = f"tdot{c}"
        dump_packed(tag + ".A", a, out)
        dump_packed(tag + ".B", b, out)
        dump_packed(tag + ".out", {q: v for q, v in res._data.items() if np.size(v)}, out)
        cases.append({"op": "tensordot", "tag": tag, "ra": ra, "rb": rb, "k": k, "dim": dim, "dist": dist,
                      "out_rank": int(res.rank), "out_dim": int(res.dim), "ref_seconds": dt})

    # --- multiply.outer
    for ra, rb, dim, dist in [(1, 1, 5, "pos"), (2, 1, 6, "normal"), (2, 2, 5, "pos"), (3, 1, 6, "pos"), (3, 2, 4, "normal"),
                              (3, 3, 3, "pos"), (4, 2, 3, "pos"), (4, 4, 2, "pos"), (4, 4, 3, "pos")]:
        rng, c = new_rng()
        a, b = rand_tensor(rng, ra, dim, dist), rand_tensor(rng, rb, dim, dist)
        A = PermClsSymmetricTensor(rank=ra, dim=dim, data={q: v.copy() for q, v in a.items()})
        B = PermClsSymmetricTensor(rank=rb, dim=dim, data={q: v.copy() for q, v in b.items()})
        t0 = time.perf_counter()
        res = symalg.multiply.outer(A, B)
        dt = time.perf_counter() - t0
        tag = f"outer{c}"
        dump_packed(tag + ".A", a, out)
        dump_packed(tag + ".B", b, out)
        dump_packed(tag + ".out", {q: v for q, v in res._data.items() if np.size(v)}, out)
        cases.append({"op": "outer", "tag": tag, "ra": ra, "rb": rb, "dim": dim, "dist": dist, "ref_seconds": dt})

    # --- the e0 (x) e1 example of testing/api.py:497-512
    e0 = PermClsSymmetricTensor(rank=1, dim=2, data={(1,): np.array([1.0, 0.0])})
    e1 = PermClsSymmetricTensor(rank=1, dim=2, data={(1,): np.array([0.0, 1.0])})
    res = symalg.multiply.outer(e0, e1)
    dump_packed("outer_e0e1.out", res._data, out)
    cases.append({"op": "outer_e0e1", "tag": "outer_e0e1"})

    # --- flat format through the same symalg defaults (unpinned by the reference's tests)
    for rank, dim in [(2, 5), (3, 4), (4, 4)]:
        rng, c = new_rng()
        n = fs.multicomb(dim, rank)
        v = rng.uniform(0.5, 1.5, n)
        x = rng.uniform(0.5, 1.5, dim)
        F = fs.FlatSymmetricTensor(rank, dim, v.copy())
        res = symalg.contract_all_indices_with_vector(F, x)
        tag = f"flatvec{c}"
        out[tag + ".A"] = v
        out[tag + ".x"] = x
        out[tag + ".out"] = np.asarray(res._data).reshape(-1)[:1].reshape(())
        cases.append({"op": "flatvec", "tag": tag, "rank": rank, "dim": dim})

    np.savez_compressed(os.path.join(GOLD, "op_goldens.npz"), **out)
    with open(os.path.join(GOLD, "op_goldens.json"), "w") as f:
        json.dump({"numpy": np.__version__, "cases": cases}, f, indent=0, sort_keys=True)


def gen_config1():
    """BASELINE config 1 on the reference itself: rank 4 dim 50 fp64 vector contraction (~9 s)."""
    rng = np.random.default_rng(SEED0 + 1000)
    data = rand_tensor(rng, 4, 50, "pos")
    x = rng.uniform(0.5, 1.5, 50) / np.sqrt(50)
    A = PermClsSymmetricTensor(rank=4, dim=50, data={k: v.copy() for k, v in data.items()})
    t0 = time.perf_counter()
    res = symalg.contract_all_indices_with_vector(A, x)
    dt = time.perf_counter() - t0
    with open(os.path.join(GOLD, "config1.json"), "w") as f:
        json.dump({"seed": SEED0 + 1000, "rank": 4, "dim": 50, "dtype": "float64",
                   "dist": "A~U[0.5,1.5) per class in class order; x~U[0.5,1.5)/sqrt(50)",
                   "result": float(np.asarray(res._data[()])), "ref_seconds": dt,
                   "packed_components": int(A.size), "cpu_count": os.cpu_count()}, f, indent=0, sort_keys=True)
    print("config1", float(np.asarray(res._data[()])), dt)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    gen_index_goldens()
    gen_op_goldens()
    gen_config1()
    print("golden fixtures written to", GOLD)
