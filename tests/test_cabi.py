"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, and its host
functions (class tables, single-index rank / unrank) agree with the oracle and the golden vectors.
No CUDA compute is called here."""
import ctypes
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, key_cls
from oracle import index_oracle as io
from oracle import packed_oracle as po

import symtensor_b200 as st
from symtensor_b200 import _cabi
from symtensor_b200 import combinatorics as comb


def header_functions():
    src = open(os.path.join(ROOT, "include", "symtensor_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_cabi.lib, n), f"{n} declared in include/symtensor_b200.h but not exported"
    assert set(names) == set(_cabi.SIGNATURES), set(names) ^ set(_cabi.SIGNATURES)
    assert _cabi.lib.st_version() >= 100


def test_class_tables_match_oracle(goldens):
    for r in range(0, 9):
        for d in (0, 1, 2, 5, 40, 50, 64, 200, 1000):
            if io.indep_size(r, d) >= 2 ** 62:  # does not fit int64 positions: reported as OverflowError
                with pytest.raises(OverflowError):
                    comb.class_table(r, d)
                continue
            t = comb.class_table(r, d)
            assert list(t.classes) == io.perm_classes(r)
            sizes = [io.permclass_size(c, d) for c in t.classes]
            if True:
                assert list(t.sizes) == sizes
                assert list(t.mults) == [io.permclass_multiplicity(c) for c in t.classes]
                off = 0
                for c, s in enumerate(sizes):
                    assert t.offsets[c] == off and off % 32 == 0
                    off += (s + 31) // 32 * 32
                assert t.total == off
                assert comb.indep_size(r, d) == io.indep_size(r, d)
    # headline sizes
    assert sum(comb.class_table(4, 200).sizes) == 68685050
    assert sum(comb.class_table(6, 64).sizes) == 119877472
    assert sum(comb.class_table(8, 40).sizes) == 314457495


def test_overflow_is_reported():
    with pytest.raises(OverflowError):
        comb.class_table(8, 1000)  # C(1007, 8) ~ 2.5e19 components
    with pytest.raises(OverflowError):
        comb.indep_size(16, 100000)


def test_host_rank_unrank_against_goldens(goldens):
    n = 0
    for k in goldens.index.files:
        if not k.startswith("sigma.r"):
            continue
        _, rs, ds, ck = k.split(".")
        r, d, cls = int(rs[1:]), int(ds[1:]), key_cls(ck)
        ref = goldens.index[k]
        step = max(1, ref.shape[0] // 40)
        for p in range(0, ref.shape[0], step):
            idx = tuple(int(v) for v in ref[p])
            assert comb.index_of(r, d, cls, p) == idx
            assert comb.convert_dense_index(r, d, idx[::-1]) == (cls, p)
            n += 1
    assert n > 500
    for idx, rep, cls in goldens.index_meta["representatives"]:
        assert list(comb.get_index_representative(idx)) == rep
        assert list(comb.get_permclass(idx)) == cls
    for k in goldens.index.files:
        if k.startswith("flat_big.") and k.endswith(".idx"):
            _, rs, ds, _ = k.split(".")
            r, d = int(rs[1:]), int(ds[1:])
            for t, p in zip(goldens.index[k].tolist(), goldens.index[k[:-3] + "rank"].tolist()):
                assert comb.flat_rank(d, t[::-1]) == p
                assert comb.flat_unrank(r, d, p) == tuple(t)


def test_error_codes_map_to_reference_exceptions():
    with pytest.raises(ValueError):
        comb.class_table(17, 3)
    with pytest.raises(IndexError):
        comb.convert_dense_index(3, 4, (0, 1, 4))
    with pytest.raises(IndexError):
        comb.convert_dense_index(3, 4, (0, 1))


def test_host_side_tensor_logic_on_cpu():
    """Packing / indexing logic of the mixin class with host storage (no kernels involved)."""
    rng = np.random.default_rng(0)
    data = {c: rng.standard_normal(io.permclass_size(c, 5)) for c in io.perm_classes(3)}
    A = st.PermClsTorchSymmetricTensor(rank=3, dim=5, data=dict(data), device="host")
    assert A.perm_classes == ["iii", "iij", "ijk"]
    assert A.packed.numel() == comb.class_table(3, 5).total
    # storage positions (symtensor/testing/api.py:308-328)
    assert float(A[0, 0, 3]) == data[(2, 1)][2]
    assert float(A[1, 2, 3]) == data[(1, 1, 1)][6]
    assert float(A[3, 0, 0]) == data[(2, 1)][2]
    dense = A.todense().numpy()
    from oracle import dense_oracle as do
    assert np.array_equal(dense, do.todense(data, 3, 5))
    B = st.PermClsTorchSymmetricTensor(data=dense, device="host")
    assert all(np.array_equal(B.to_numpy_dict()[c], data[c]) for c in data)
    with pytest.raises(ValueError, match="not symmetric"):
        bad = dense.copy()
        bad[0, 1, 2] += 1.0
        st.PermClsTorchSymmetricTensor(data=bad, device="host")
    # scalar-compressed classes are expanded; missing classes are zeros
    C = st.PermClsTorchSymmetricTensor(rank=3, dim=5, data={"iij": 2.0}, device="host")
    assert float(C[1, 1, 0]) == 2.0 and float(C[0, 1, 2]) == 0.0 and C["iij"].shape == (20,)
    C["ijk"] = 1.5
    C[0, 0, 0] = -1.0
    assert float(C[2, 1, 0]) == 1.5 and float(C["iii"][0]) == -1.0
    assert list(A.indep_iter_repindex())[:7] == [(0, 0, 0), (1, 1, 1), (2, 2, 2), (3, 3, 3), (4, 4, 4), (0, 0, 1), (0, 0, 2)]
    F = st.FlatSymmetricTensor(3, 4, np.arange(20.0), device="host")
    assert float(F[1, 2, 3]) == float(po.permcls_to_flat(po.flat_to_permcls(np.arange(20.0), 3, 4), 3, 4)[comb.flat_rank(4, (1, 2, 3))])
    # reference error behaviour that needs no GPU
    with pytest.raises(ValueError, match="must match"):
        st.contract_all_indices_with_vector(A, np.ones(4))
    assert st.contract_all_indices_with_vector(A, np.zeros(5)) == 0
    with pytest.raises(TypeError):  # operands of different dimension: no implementation (symtensor/symalg.py:305-308)
        st.symalg.add.outer(A, st.PermClsTorchSymmetricTensor(rank=2, dim=4, device="host"))


@pytest.fixture(scope="module")
def emu_lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = os.path.join(ROOT, "tests", "emu", "libst_emu.so")
    src = os.path.join(ROOT, "tests", "emu", "emu_vec.cu")
    libdir = os.path.join(ROOT, "symtensor_b200", "lib")
    deps = [src, os.path.join(libdir, "libsymtensor_b200.so")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", out, src, "-L" + libdir,
                        "-lsymtensor_b200", "-Xlinker", "-rpath", "-Xlinker", libdir], check=True, capture_output=True)
    return ctypes.CDLL(out)


def test_tail_kernel_index_walk_emulated_on_cpu(emu_lib):
    """The ring kernel's strategy / chunk / tile / segment / block walk and its per-warp stream cursors (host+device
    code of st_vec_core.cuh), replayed serially on the CPU by tests/emu, against the packed oracle -- all classes,
    forced tail lengths, ring geometries, sharded and ragged ranges."""
    rng = np.random.default_rng(1)
    i64 = ctypes.c_int64

    def pack(data, rank, dim):
        t = comb.class_table(rank, dim)
        buf = np.zeros(t.total)
        for c, s, o in zip(t.classes, t.sizes, t.offsets):
            buf[o:o + s] = data[c]
        return buf

    def emu(rank, dim, buf, x, begin=0, end=None, nwarps=4, grid=3, item=64, bel=16, slots=2, tau=0, use_dir=1, small=0):
        end = len(buf) if end is None else end
        out = ctypes.c_double()
        taus = (ctypes.c_int32 * comb.class_table(rank, dim).ncls)()
        rc = emu_lib.emu_contract_vec_f64(rank, i64(dim), ctypes.c_void_p(buf[begin:].ctypes.data), i64(begin), i64(end),
                                          ctypes.c_void_p(x.ctypes.data), nwarps, grid, i64(item), bel, slots, tau, use_dir, i64(small),
                                          ctypes.byref(out), taus)
        assert rc == 0, rc
        return out.value

    # the table-free pair walk at dimensions that need 8 / 12 column slots and the per-lane fallback (Rt > 384)
    for rank, dim in [(2, 150), (3, 140), (2, 300), (2, 400), (3, 70)]:
        data = {c: rng.uniform(0.5, 1.5, io.permclass_size(c, dim)) for c in io.perm_classes(rank)}
        x = rng.uniform(0.5, 1.5, dim) / np.sqrt(dim)
        ref = po.contract_all_indices_with_vector(data, rank, dim, x)
        buf = pack(data, rank, dim)
        for tau in (-1, -2, -3):
            for nw, grid, item, bel, slots in [(4, 3, 256, 64, 2), (3, 2, 192, 192, 3)]:
                got = emu(rank, dim, buf, x, nwarps=nw, grid=grid, item=item, bel=bel, slots=slots, tau=tau)
                assert abs(got - ref) <= 1e-12 * abs(ref), (rank, dim, tau, nw, grid, item, bel, slots)
    for rank, dim in [(1, 7), (2, 9), (3, 6), (3, 13), (4, 5), (4, 11), (5, 7), (6, 7), (7, 8), (8, 9), (4, 40)]:
        data = {c: rng.uniform(0.5, 1.5, io.permclass_size(c, dim)) for c in io.perm_classes(rank)}
        x = rng.uniform(0.5, 1.5, dim)
        ref = po.contract_all_indices_with_vector(data, rank, dim, x)
        buf = pack(data, rank, dim)
        for tau in (0, 1, 2, 3, -1, -2, -3):  # < 0: the hybrid pair walk (no table / all table / 2 KB suffix table)
            for nw, grid, item, bel, slots in [(4, 3, 64, 16, 2), (12, 5, 1024, 256, 2), (2, 7, 32, 32, 3), (1, 2, 1024, 64, 4)]:
                for use_dir in (0, 1):
                    got = emu(rank, dim, buf, x, nwarps=nw, grid=grid, item=item, bel=bel, slots=slots, tau=tau, use_dir=use_dir)
                    assert abs(got - ref) <= 1e-12 * abs(ref), (rank, dim, tau, nw, grid, item, bel, slots, use_dir)
        for small in (100, 10 ** 9):  # some / all classes through the per-component phase
            assert abs(emu(rank, dim, buf, x, small=small) - ref) <= 1e-12 * abs(ref)
        # dynamic deal (claims served in replay order): costly tail first, groups of g tiles, single tiles last;
        # use_dir = 2 * g + directory bit (the emulation also checks that the deal is a bijection onto the tiles)
        for g in (1, 2, 3, 4):
            for tau in (0, 2, -3):
                for nw, grid, item, bel, slots in [(4, 3, 64, 16, 2), (2, 7, 32, 32, 3)]:
                    got = emu(rank, dim, buf, x, nwarps=nw, grid=grid, item=item, bel=bel, slots=slots, tau=tau, use_dir=2 * g + 1)
                    assert abs(got - ref) <= 1e-12 * abs(ref), (rank, dim, tau, nw, grid, item, bel, slots, g)
        tot = len(buf)
        cut0 = (tot // 2) // 32 * 32
        s = emu(rank, dim, buf, x, begin=0, end=cut0, use_dir=2 * 3 + 1) + emu(rank, dim, buf, x, begin=cut0, end=tot, use_dir=2 * 3 + 1)
        assert abs(s - ref) <= 1e-12 * abs(ref)
        cut = (tot // 3) // 32 * 32
        s = sum(emu(rank, dim, buf, x, begin=b, end=e, small=sm) for sm, (b, e) in zip((0, 50, 0), [(0, cut), (cut, 2 * cut), (2 * cut, tot)]))
        assert abs(s - ref) <= 1e-12 * abs(ref)
        # ragged ends: the last 16-byte unit of a tile is incomplete (fetched outside the bulk copy)
        e1 = min(tot, cut + 37)
        s = emu(rank, dim, buf, x, begin=0, end=e1, item=32, bel=32) + emu(rank, dim, buf, x, begin=e1 // 32 * 32, end=tot, item=128, bel=64)
        s -= emu(rank, dim, buf, x, begin=e1 // 32 * 32, end=e1)
        assert abs(s - ref) <= 1e-12 * abs(ref)
        # fp32 instantiation (4 components per 16-byte vector)
        buf32, x32 = buf.astype(np.float32), x.astype(np.float32)
        out = ctypes.c_double()
        for tau in (0, 2):
            rc = emu_lib.emu_contract_vec_f32(rank, i64(dim), ctypes.c_void_p(buf32.ctypes.data), i64(0), i64(len(buf32)),
                                              ctypes.c_void_p(x32.ctypes.data), 4, 3, i64(4096), 512, 2, tau, 1, i64(0), ctypes.byref(out), None)
            assert rc == 0 and abs(out.value - ref) <= 1e-5 * abs(ref)
