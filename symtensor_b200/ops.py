"""CUDA implementations of the four ``symalg`` ops, registered on the GPU tensor classes exactly the way a
reference backend mixin registers them (``@Cls.implements(symalg.f)``, symtensor/base.py:1057-1063; in-tree
precedent symtensor/decomp_symmtensor.py:1012-1038).  Each function validates arguments with the reference's
error behaviour, then makes ONE call into the C-ABI with raw device pointers on torch's current stream.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import symalg
from ._cabi import VEC_OVERLAP, c_i64, check, lib
from .flat import CudaFlatSymmetricTensor
from .permcls import _NP2TORCH, _TORCH2NP, CudaPermClsSymmetricTensor, _stream_ptr

_WS_BYTES = int(lib.st_contract_vec_workspace_bytes())


def _np_dtype_of(x) -> np.dtype:
    if isinstance(x, torch.Tensor):
        return _TORCH2NP.get(x.dtype, np.dtype("float64"))
    return np.asarray(x).dtype


def _promote(a_dtype: np.dtype, x) -> torch.dtype:
    """NumPy promotion of the operands (fp32 tensor with fp64 x gives fp64 -- SURVEY.md B.2)."""
    xd = _np_dtype_of(x)
    if xd.kind in "iub":
        xd = a_dtype
    res = np.result_type(a_dtype, xd)
    if res not in _NP2TORCH:
        raise TypeError(f"unsupported result dtype {res}")
    return _NP2TORCH[res]


def _all_close_to_zero(x) -> bool:
    if isinstance(x, torch.Tensor):
        return bool(torch.isclose(x, torch.zeros((), dtype=x.dtype, device=x.device)).all().item())
    return bool(np.isclose(np.asarray(x), 0).all())


def contract_vec_device(A, x_dev: torch.Tensor, out: torch.Tensor, ws: torch.Tensor, begin: int = 0, end: int = None,
                        packed: torch.Tensor = None, overlap: bool = False) -> None:
    """Raw launch: partial sum of packed coordinates [begin, end) into the 0-d/1-element ``out``.
    ``packed`` is the local shard starting at coordinate ``begin`` (default: the whole buffer).
    ``overlap`` (``ST_VEC_OVERLAP``): the call belongs to a batch of contractions of resident operands -- the launch starts
    while the previous one on the stream drains its tail; ``ws`` then holds ``2 * st_contract_vec_workspace_bytes()``."""
    packed = A._buf if packed is None else packed
    total = A._buf.numel() if (end is None and packed is A._buf) else end
    f64 = packed.dtype == torch.float64
    if overlap:
        if ws.numel() * ws.element_size() < 2 * _WS_BYTES:
            raise ValueError("overlapping launches need a workspace of 2 * st_contract_vec_workspace_bytes()")
        fn = lib.st_contract_vec_ex_f64 if f64 else lib.st_contract_vec_ex_f32
        check(fn(A.layout, A.rank, c_i64(A.dim), packed.data_ptr(), c_i64(begin), c_i64(total), x_dev.data_ptr(),
                 out.data_ptr(), ws.data_ptr(), VEC_OVERLAP, _stream_ptr(packed.device)))
        return
    fn = lib.st_contract_vec_f64 if f64 else lib.st_contract_vec_f32
    check(fn(A.layout, A.rank, c_i64(A.dim), packed.data_ptr(), c_i64(begin), c_i64(total), x_dev.data_ptr(),
             out.data_ptr(), ws.data_ptr(), _stream_ptr(packed.device)))


def contract_all_indices_with_vectors(symtensor, X):
    """Batch form of ``contract_all_indices_with_vector`` (symtensor/symalg.py:505-527): one resident tensor contracted
    with every row of ``X`` (``n x dim``) -- the polynomial evaluated at n points.  One kernel launch per row, chained with
    ``ST_VEC_OVERLAP`` so that consecutive launches overlap their ramp-up and tail.  Returns a length-n torch tensor."""
    _need_device(symtensor)
    dev = symtensor.device
    Xt = X if isinstance(X, torch.Tensor) else torch.as_tensor(np.asarray(X))
    if Xt.ndim != 2 or Xt.shape[1] != symtensor.dim:
        raise ValueError(f"X must have shape (n, {symtensor.dim}); received {tuple(Xt.shape)}")
    tdt = _promote(symtensor.dtype, Xt)
    buf = symtensor._buf if symtensor._buf.dtype == tdt else symtensor._buf.to(tdt)
    n = Xt.shape[0]
    with torch.cuda.device(dev):
        Xd = Xt.to(device=dev, dtype=tdt).contiguous()
        out = torch.zeros(max(n, 1), dtype=tdt, device=dev)
        ws = _vec_workspace(dev)
        torch.cuda.current_stream(dev).synchronize()  # the operands are complete: the launches below may start early
        fn = lib.st_contract_vec_ex_f64 if tdt == torch.float64 else lib.st_contract_vec_ex_f32
        sp = _stream_ptr(dev)
        for i in range(n):
            check(fn(type(symtensor).layout, symtensor.rank, c_i64(symtensor.dim), buf.data_ptr(), c_i64(0), c_i64(buf.numel()),
                     Xd[i].data_ptr(), out[i:].data_ptr(), ws.data_ptr(), VEC_OVERLAP, sp))
    return out[:n]


_VEC_WS = {}


def _vec_workspace(dev: torch.device) -> torch.Tensor:
    """The vector contraction's 4 MiB partial-sum workspace, one per (device, stream): calls on one stream are ordered, so they can
    share it (a fresh allocation per call cost more than the kernel of a small tensor)."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), _stream_ptr(dev))
    ws = _VEC_WS.get(key)
    if ws is None:
        ws = _VEC_WS[key] = torch.empty(2 * _WS_BYTES // 8, dtype=torch.float64, device=dev)
    return ws


def _contract_all_indices_with_vector(symtensor, x):
    """symtensor/symalg.py:505-527."""
    if len(x) != symtensor.dim:
        raise ValueError("Dimensions of tensor and vector must match; received "
                         f"{symtensor.dim} (tensor) and {len(x)} (vector).")
    if _all_close_to_zero(x):
        return 0
    cls = type(symtensor)
    tdt = _promote(symtensor.dtype, x)
    buf = symtensor._buf if symtensor._buf.dtype == tdt else symtensor._buf.to(tdt)
    if symtensor._host:
        # host buffers: stream the packed data through the GPU in chunks (copies overlap the kernel)
        if not torch.cuda.is_available():
            raise RuntimeError("symtensor_b200 needs a CUDA device; there is no CPU fallback")
        xh = (x.detach().cpu() if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(tdt).contiguous()
        out = torch.zeros(32 if cls.layout == 0 else 1, dtype=tdt)
        fn = lib.st_contract_vec_host_f64 if tdt == torch.float64 else lib.st_contract_vec_host_f32
        check(fn(cls.layout, symtensor.rank, c_i64(symtensor.dim), buf.data_ptr(), c_i64(buf.numel()), xh.data_ptr(),
                 out.data_ptr()))
        return cls.from_packed(0, 1, out)
    dev = symtensor.device
    with torch.cuda.device(dev):
        xd = (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(device=dev, dtype=tdt).contiguous()
        out = torch.zeros(32 if cls.layout == 0 else 1, dtype=tdt, device=dev)  # packed buffer of a rank-0 tensor
        ws = _vec_workspace(dev)
        fn = lib.st_contract_vec_f64 if tdt == torch.float64 else lib.st_contract_vec_f32
        check(fn(cls.layout, symtensor.rank, c_i64(symtensor.dim), buf.data_ptr(), c_i64(0), c_i64(buf.numel()),
                 xd.data_ptr(), out.data_ptr(), ws.data_ptr(), _stream_ptr(dev)))
    return cls.from_packed(0, 1, out)


for _cls in (CudaPermClsSymmetricTensor, CudaFlatSymmetricTensor):
    _cls.implements(symalg.contract_all_indices_with_vector)(_contract_all_indices_with_vector)


# ------------------------------------------------------------------------------------------------------------
# tensor-valued ops: multiply.outer, tensordot, contract_all_indices_with_matrix
# ------------------------------------------------------------------------------------------------------------
_SYM = (CudaPermClsSymmetricTensor, CudaFlatSymmetricTensor)


def _fn(name: str, tdt: torch.dtype):
    return getattr(lib, f"{name}_{'f64' if tdt == torch.float64 else 'f32'}")


def _need_device(t):
    if t._host:
        raise RuntimeError("symtensor_b200: this op needs the tensor on a CUDA device (use .to('cuda')); there is no CPU "
                           "fallback")


def _flat_buffer(t, tdt: torch.dtype) -> torch.Tensor:
    """The components of ``t`` in the flat order (combinations_with_replacement), on its device, as ``tdt``."""
    _need_device(t)
    buf = t._buf if t._buf.dtype == tdt else t._buf.to(tdt)
    if isinstance(t, CudaFlatSymmetricTensor):
        return buf.contiguous()
    out = torch.empty(max(1, t.indep_size), dtype=tdt, device=t.device)
    with torch.cuda.device(t.device):
        check(_fn("st_permcls_to_flat", tdt)(t.rank, c_i64(t.dim), buf.data_ptr(), out.data_ptr(), _stream_ptr(t.device)))
    return out


def _wrap_flat_result(cls, rank: int, dim: int, flat: torch.Tensor):
    """A tensor of class ``cls`` from components in the flat order."""
    if issubclass(cls, CudaFlatSymmetricTensor):
        return cls.from_packed(rank, dim, flat)
    total = comb_total(rank, dim)
    out = torch.empty(total, dtype=flat.dtype, device=flat.device)
    with torch.cuda.device(flat.device):
        check(_fn("st_flat_to_permcls", flat.dtype)(rank, c_i64(dim), flat.data_ptr(), out.data_ptr(), c_i64(0), c_i64(total),
                                                    _stream_ptr(flat.device)))
    return cls.from_packed(rank, dim, out)


def comb_total(rank: int, dim: int) -> int:
    from . import combinatorics as comb
    return comb.class_table(rank, dim).total


def _as_symtensor_operand(x, like):
    """ndarray / torch / scalar operands of tensordot and outer.  Vectors and scalars are symmetric tensors already.  A
    dense array of rank >= 2 must be symmetric: the reference accepts ARBITRARY dense operands there and symmetrizes only
    the result (``Sym(np.tensordot(a.todense(), b, axes))``, symtensor/symalg.py:427-459, 294-316), which for a
    non-symmetric operand is a partially symmetric contraction the packed kernels do not implement -- that case raises
    ``NotImplementedError`` (documented deviation; ``symtensor_b200.plugin`` hands it to the reference's own default).
    Python scalars are weakly typed like in NumPy (fp32 tensor * 2.0 stays fp32)."""
    if isinstance(x, _SYM):
        return x
    if isinstance(x, (bool, int, float)):
        x = torch.tensor(x, dtype=like.torch_dtype)
    arr = x if isinstance(x, torch.Tensor) else np.asarray(x)
    nd = arr.ndim
    if nd == 0:
        return type(like)(rank=0, dim=1, data={(): arr} if isinstance(like, CudaPermClsSymmetricTensor) else arr, device=like.device)
    try:
        if isinstance(like, CudaPermClsSymmetricTensor):
            return CudaPermClsSymmetricTensor(data=arr, device=like.device) if nd > 1 else \
                CudaPermClsSymmetricTensor(rank=1, dim=arr.shape[0], data={(1,): arr}, device=like.device)
        return CudaFlatSymmetricTensor(nd, arr.shape[0], arr, device=like.device)
    except ValueError as e:
        if "not symmetric" in str(e):
            raise NotImplementedError("symtensor_b200: tensordot / outer with a NON-symmetric dense operand of rank >= 2 is not "
                                      "implemented on the packed kernels (the reference symmetrizes only the result); symmetrize the "
                                      "operand or use the reference default") from e
        raise


def _result_dtype(a, b) -> torch.dtype:
    res = np.result_type(a.dtype, b.dtype)
    if res not in _NP2TORCH:
        raise TypeError(f"unsupported result dtype {res}")
    return _NP2TORCH[res]


def outer_device(a, b, out_rank_buf: torch.Tensor, begin: int, end: int, tdt: torch.dtype, af=None, bf=None, op: int = 0):
    """Raw launch: coordinates [begin, end) of the permcls buffer of a (x)_s b into ``out_rank_buf`` (which starts at
    coordinate ``begin``) -- the output range is the sharding axis for multi-GPU runs.  ``op``: 0 multiply (the
    compile-time-rank kernels), 1 add, 2 subtract (``st_outer_op_*``)."""
    af = _flat_buffer(a, tdt) if af is None else af
    bf = _flat_buffer(b, tdt) if bf is None else bf
    with torch.cuda.device(af.device):
        if op == 0:
            check(_fn("st_outer", tdt)(a.rank, b.rank, c_i64(a.dim), af.data_ptr(), bf.data_ptr(), out_rank_buf.data_ptr(),
                                       c_i64(begin), c_i64(end), _stream_ptr(af.device)))
        else:
            check(_fn("st_outer_op", tdt)(op, a.rank, b.rank, c_i64(a.dim), af.data_ptr(), bf.data_ptr(), out_rank_buf.data_ptr(),
                                          c_i64(begin), c_i64(end), _stream_ptr(af.device)))


def _outer(ufunc, a, b, **kwargs):
    """symtensor/symalg.py:294-316 (registered there for add, subtract and multiply alike):
    C_K = mean over the C(n, ra) position splits of A[K_S] (op) B[K_S^c]."""
    op = {symalg.multiply: 0, symalg.add: 1, symalg.subtract: 2}.get(ufunc)
    if op is None:
        return NotImplemented
    like = a if isinstance(a, _SYM) else b
    ranka, rankb = np.ndim(a) if not isinstance(a, _SYM) else a.rank, np.ndim(b) if not isinstance(b, _SYM) else b.rank
    dima = a.dim if isinstance(a, _SYM) else (*np.shape(a), 1)[0]
    dimb = b.dim if isinstance(b, _SYM) else (*np.shape(b), 1)[0]
    if ranka != 0 and rankb != 0 and dima != dimb:
        return NotImplemented
    symargs = tuple(x for x in (a, b) if isinstance(x, _SYM))
    cls = symalg.result_array(*symargs)
    a, b = _as_symtensor_operand(a, like), _as_symtensor_operand(b, like)
    out = kwargs.pop("out", None)
    tdt = _result_dtype(a, b) if out is None else out.torch_dtype
    n = a.rank + b.rank
    dim = a.dim if a.rank else b.dim
    if a.rank == 0 or b.rank == 0:  # scalar (op) tensor: one split, elementwise on the packed buffer
        s, t = (a, b) if a.rank == 0 else (b, a)
        sv, tv = s._buf.to(tdt).reshape(-1)[0], t._buf.to(tdt)
        val = tv * sv if op == 0 else tv + sv if op == 1 else (sv - tv if a.rank == 0 else tv - sv)
        if op and isinstance(t, CudaPermClsSymmetricTensor):  # keep the alignment padding at zero
            tab = t.class_table
            for size, off, nxt in zip(tab.sizes, tab.offsets, list(tab.offsets[1:]) + [tab.total]):
                val[off + size:nxt] = 0
        res = cls.from_packed(t.rank, t.dim, val)
        if out is not None:
            out._buf.copy_(res._buf)
            return out
        return res
    if issubclass(cls, CudaFlatSymmetricTensor):
        tmp = torch.empty(comb_total(n, dim), dtype=tdt, device=a.device)
        outer_device(a, b, tmp, 0, tmp.numel(), tdt, op=op)
        res_p = CudaPermClsSymmetricTensor.from_packed(n, dim, tmp)
        res = cls.from_packed(n, dim, _flat_buffer(res_p, tdt))
    else:
        buf = out._buf if (out is not None and isinstance(out, CudaPermClsSymmetricTensor)) else \
            torch.empty(comb_total(n, dim), dtype=tdt, device=a.device)
        outer_device(a, b, buf, 0, buf.numel(), tdt, op=op)
        res = out if buf is getattr(out, "_buf", None) else cls.from_packed(n, dim, buf)
    if out is not None and res is not out:
        out._buf.copy_(res._buf)
        return out
    return res


def _normalize_axes(axes, ra: int, rb: int) -> int:
    """Only the NUMBER of contracted axes matters for symmetric operands (symtensor/testing/api.py:546-552)."""
    if isinstance(axes, (int, np.integer)):
        k = int(axes)
    else:
        ax_a, ax_b = axes
        ax_a = [ax_a] if isinstance(ax_a, (int, np.integer)) else list(ax_a)
        ax_b = [ax_b] if isinstance(ax_b, (int, np.integer)) else list(ax_b)
        if len(ax_a) != len(ax_b):
            raise ValueError("shape-mismatch for sum")
        k = len(ax_a)
    if k < 0 or k > ra or k > rb:
        raise ValueError(f"cannot contract {k} axes of tensors with ranks {ra} and {rb}")
    return k


def _tensordot(a, b, axes=2):
    """symtensor/symalg.py:427-459: Sym(sum over `axes` contracted index pairs)."""
    like = a if isinstance(a, _SYM) else b
    cls = symalg.result_array(*(x for x in (a, b) if isinstance(x, _SYM)))
    a, b = _as_symtensor_operand(a, like), _as_symtensor_operand(b, like)
    k = _normalize_axes(axes, a.rank, b.rank)
    if a.rank and b.rank and a.dim != b.dim:
        raise ValueError(f"shape-mismatch for sum: dimensions {a.dim} and {b.dim}")
    if k == 0:
        return _outer(symalg.multiply, a, b)
    tdt = _result_dtype(a, b)
    dim = a.dim
    n = a.rank + b.rank - 2 * k
    out_dim = dim if n else 1
    total = comb_total(n, out_dim)
    with torch.cuda.device(a.device):
        buf = torch.empty(total, dtype=tdt, device=a.device)
    tensordot_device(a, b, k, buf, 0, total, tdt)
    res = CudaPermClsSymmetricTensor.from_packed(n, out_dim, buf)
    if issubclass(cls, CudaFlatSymmetricTensor):
        return cls.from_packed(n, out_dim, _flat_buffer(res, tdt))
    return res if cls is CudaPermClsSymmetricTensor else cls.from_packed(n, out_dim, buf)


def tensordot_device(a, b, k: int, out_range_buf: torch.Tensor, begin: int, end: int, tdt: torch.dtype, af=None, bf=None, ws=None,
                     check_flag: bool = True):
    """Raw launch: coordinates [begin, end) of the permcls buffer of ``tensordot(a, b, axes=k)`` into ``out_range_buf`` (which starts
    at coordinate ``begin``) -- the output range is the sharding axis for multi-GPU runs (BASELINE config 3's 168 GB output only
    exists sharded).  fp32 with two free indices on each side runs the tiled tcgen05 kernel (``st_tensordot_is_tiled``), everything
    else the materialised Gram matrix."""
    af = _flat_buffer(a, tdt) if af is None else af
    bf = _flat_buffer(b, tdt) if bf is None else bf
    nbytes = c_i64(0)
    check(lib.st_tensordot_workspace_bytes(a.rank, b.rank, k, c_i64(a.dim), af.element_size(), ctypes.byref(nbytes)))
    tiled = bool(lib.st_tensordot_is_tiled(a.rank, b.rank, k, c_i64(a.dim), af.element_size()))
    if nbytes.value > 150 << 30:
        raise NotImplementedError(f"symtensor_b200.tensordot: the workspace of this shape needs {nbytes.value / 2 ** 30:.0f} GiB "
                                  "(the materialised pair-packed Gram matrix; only fp32 with two free indices per side has the tiled kernel)")
    with torch.cuda.device(af.device):
        if ws is None:
            ws = torch.empty(max(1, (nbytes.value + af.element_size() - 1) // af.element_size()), dtype=tdt, device=af.device)
        check(_fn("st_tensordot", tdt)(a.rank, b.rank, k, c_i64(a.dim), af.data_ptr(), bf.data_ptr(), out_range_buf.data_ptr(), c_i64(begin),
                                       c_i64(end), ws.data_ptr(), _stream_ptr(af.device)))
        if tiled and check_flag and int(ws[:1].view(torch.int32)[0].item()) != 0:
            raise RuntimeError("symtensor_b200.tensordot: the tiled tcgen05 kernel gave up on a barrier (internal error; the result is invalid)")
    return ws


def tensordot_device_ranges(a, b, k: int, outs, ranges, af=None, bf=None, ws=None, check_flag: bool = True):
    """Raw launch (fp32): several DISJOINT coordinate ranges ``[(begin, end), ...]`` of the permcls buffer of ``tensordot(a, b, axes=k)``
    into the buffers ``outs`` in ONE call (``st_tensordot_ranges_f32``): the shard a GPU gets from ``sharding.tensordot22_shards``."""
    tdt = torch.float32
    af = _flat_buffer(a, tdt) if af is None else af
    bf = _flat_buffer(b, tdt) if bf is None else bf
    nbytes = c_i64(0)
    check(lib.st_tensordot_workspace_bytes(a.rank, b.rank, k, c_i64(a.dim), 4, ctypes.byref(nbytes)))
    tiled = bool(lib.st_tensordot_is_tiled(a.rank, b.rank, k, c_i64(a.dim), 4))
    n = len(ranges)
    begins = (ctypes.c_int64 * n)(*[int(r[0]) for r in ranges])
    ends = (ctypes.c_int64 * n)(*[int(r[1]) for r in ranges])
    ptrs = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    with torch.cuda.device(af.device):
        if ws is None:
            ws = torch.empty(max(1, (nbytes.value + 3) // 4), dtype=tdt, device=af.device)
        check(lib.st_tensordot_ranges_f32(a.rank, b.rank, k, c_i64(a.dim), af.data_ptr(), bf.data_ptr(), n, begins, ends, ptrs, ws.data_ptr(),
                                          _stream_ptr(af.device)))
        if tiled and check_flag and int(ws[:1].view(torch.int32)[0].item()) != 0:
            raise RuntimeError("symtensor_b200.tensordot: the tiled tcgen05 kernel gave up on a barrier (internal error; the result is invalid)")
    return ws


def _contract_all_indices_with_matrix(symtensor, W):
    """symtensor/symalg.py:475-496: C[j1..jr] = sum A[i1..ir] W[i1,j1]...W[ir,jr] (W contracted on its first axis)."""
    if not isinstance(symtensor, _SYM):
        return NotImplemented
    cls = type(symtensor)
    Wt = W if isinstance(W, torch.Tensor) else torch.as_tensor(np.asarray(W))
    d = symtensor.dim
    if symtensor.rank > 0 and tuple(Wt.shape) != (d, d):
        raise ValueError(f"W must have shape {(d, d)} to contract a tensor of dimension {d}; received {tuple(Wt.shape)}")
    wd = _TORCH2NP.get(Wt.dtype, np.dtype("float64"))
    tdt = _promote(symtensor.dtype, np.empty(0, dtype=wd if wd.kind == "f" else symtensor.dtype))
    af = _flat_buffer(symtensor, tdt)
    nbytes = c_i64(0)
    check(lib.st_contract_mat_workspace_bytes(symtensor.rank, c_i64(d), af.element_size(), ctypes.byref(nbytes)))
    with torch.cuda.device(af.device):
        Wd = Wt.to(device=af.device, dtype=tdt).contiguous()
        ws = torch.empty(max(1, nbytes.value // af.element_size()), dtype=tdt, device=af.device)
        outf = torch.empty_like(af)
        check(_fn("st_contract_mat", tdt)(symtensor.rank, c_i64(d), af.data_ptr(), Wd.data_ptr(), outf.data_ptr(), ws.data_ptr(),
                                          _stream_ptr(af.device)))
    return _wrap_flat_result(cls, symtensor.rank, d, outf)


def contract_mat_device(symtensor, W, jlo: int, jhi: int, af=None):
    """Raw launch of the matrix contraction for the output components whose first (smallest) mode lies in ``[jlo, jhi)`` --
    the multi-GPU partition by the first output mode (``st_contract_mat_range_*``): returns
    ``(flat_slice, flat_begin, flat_end)``, the contiguous range of the FLAT-ordered output this call computed.  The
    intermediates of the mode chain are sharded the same way, so the workspace shrinks with the range."""
    Wt = W if isinstance(W, torch.Tensor) else torch.as_tensor(np.asarray(W))
    d = symtensor.dim
    if tuple(Wt.shape) != (d, d):
        raise ValueError(f"W must have shape {(d, d)}")
    tdt = symtensor.torch_dtype
    af = _flat_buffer(symtensor, tdt) if af is None else af
    b, e, nbytes = c_i64(0), c_i64(0), c_i64(0)
    check(lib.st_contract_mat_range_bounds(symtensor.rank, c_i64(d), c_i64(jlo), c_i64(jhi), ctypes.byref(b), ctypes.byref(e)))
    check(lib.st_contract_mat_range_workspace_bytes(symtensor.rank, c_i64(d), c_i64(jlo), c_i64(jhi), af.element_size(), ctypes.byref(nbytes)))
    with torch.cuda.device(af.device):
        Wd = Wt.to(device=af.device, dtype=tdt).contiguous()
        ws = torch.empty(max(1, nbytes.value // af.element_size() + 1), dtype=tdt, device=af.device)
        out = torch.empty(max(0, e.value - b.value), dtype=tdt, device=af.device)
        check(_fn("st_contract_mat_range", tdt)(symtensor.rank, c_i64(d), af.data_ptr(), Wd.data_ptr(), out.data_ptr(), c_i64(jlo), c_i64(jhi),
                                                ws.data_ptr(), _stream_ptr(af.device)))
    return out, b.value, e.value


def outer_then_contract_vec(a, b, x):
    """Fused ``contract_all_indices_with_vector(multiply.outer(a, b), x)``: the rank-(ra+rb) tensor is never stored
    (BASELINE config 5's pipeline in one pass).  Returns a rank-0 tensor of the operands' class."""
    like = a
    cls = symalg.result_array(a, b)
    tdt = _promote(np.result_type(a.dtype, b.dtype), x)
    if len(x) != a.dim or a.dim != b.dim:
        raise ValueError("Dimensions of tensors and vector must match")
    af, bf = _flat_buffer(a, tdt), _flat_buffer(b, tdt)
    n = a.rank + b.rank
    total = comb_total(n, a.dim)
    with torch.cuda.device(af.device):
        xd = (x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))).to(device=af.device, dtype=tdt).contiguous()
        out = torch.zeros(32 if cls.layout == 0 else 1, dtype=tdt, device=af.device)
        ws = torch.empty(int(lib.st_outer_vec_workspace_bytes()) // 8, dtype=torch.float64, device=af.device)
        check(_fn("st_outer_vec", tdt)(a.rank, b.rank, c_i64(a.dim), af.data_ptr(), bf.data_ptr(), xd.data_ptr(), out.data_ptr(),
                                       ws.data_ptr(), c_i64(0), c_i64(total), _stream_ptr(af.device)))
    return cls.from_packed(0, 1, out)


for _cls in _SYM:
    _cls.implements(symalg.tensordot)(_tensordot)
    _cls.implements(symalg.contract_all_indices_with_matrix)(_contract_all_indices_with_matrix)
    _cls.implements_ufunc.outer(symalg.add, symalg.subtract, symalg.multiply)(_outer)
