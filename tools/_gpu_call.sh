timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py -x -q -m gpu -k "tensordot or config3" > gpurun_out/r2d_pytest_tdot.log 2>&1; tail -4 gpurun_out/r2d_pytest_tdot.log
