python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py -x -q -m gpu -k "outer or config5 or converters" > gpurun_out/r2b_pytest_outer.log 2>&1; tail -2 gpurun_out/r2b_pytest_outer.log
for i in 1; do python tools/one_outer.py 0.5 0.0 1; python tools/one_outer.py 0.5 0.5 1; done > gpurun_out/r2d_outer_plain.log 2>&1
cat gpurun_out/r2d_outer_plain.log
