"""symtensor_b200 -- B200-native backend for symtensor's symmetrized-contraction hot path.

Drop-in surface (same names as the reference):

    import symtensor_b200 as st
    A = st.PermClsTorchSymmetricTensor(rank=4, dim=200, data={...})        # data lives in HBM
    s = st.contract_all_indices_with_vector(A, x)                          # one CUDA streaming pass
    st.symalg.multiply.outer(A, B); st.tensordot(A, B, axes=1); st.contract_all_indices_with_matrix(A, W)

The compute path is the C-ABI library ``lib/libsymtensor_b200.so`` (hand-written sm_100a kernels); importing
this package fails loudly if it has not been built.
"""
from . import _cabi, combinatorics, symalg  # noqa: F401
from .base import SymmetricTensor, result_array  # noqa: F401
from .flat import CudaFlatSymmetricTensor, FlatSymmetricTensor  # noqa: F401
from .permcls import (CudaPermClsSymmetricTensor, PermClsSymmetricTensor, PermClsTorchSymmetricTensor,  # noqa: F401
                      TorchPermClsSymmetricTensor)
from . import ops  # noqa: F401  (registers the CUDA implementations)
from .symalg import (add, contract_all_indices_with_matrix, contract_all_indices_with_vector, contract_tensor_list,  # noqa: F401
                     multiply, subtract, tensordot)

__version__ = "0.1.0"
