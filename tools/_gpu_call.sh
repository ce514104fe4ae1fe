timeout 900 python -m pytest tests/test_gpu_vec.py tests/test_plugin.py tests/test_gpu_elementwise.py -x -q -m gpu > gpurun_out/r2g_pytest_vec.log 2>&1; tail -2 gpurun_out/r2g_pytest_vec.log
python - <<'P'
import torch, time, numpy as np
import symtensor_b200 as st
from symtensor_b200 import combinatorics as comb
t=comb.class_table(4,50)
A=st.PermClsTorchSymmetricTensor.from_packed(4,50,torch.rand(t.total,dtype=torch.float64,device="cuda")+0.5)
x=torch.rand(50,dtype=torch.float64,device="cuda")
for _ in range(5): st.contract_all_indices_with_vector(A,x)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(200): r=st.contract_all_indices_with_vector(A,x)
torch.cuda.synchronize(); print("public API, config 1 (rank 4 dim 50 fp64): %.1f us per call" % ((time.perf_counter()-t0)/200*1e6))
P
