python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py -x -q -m gpu -k "outer or config5" > gpurun_out/r2b_pytest_outer.log 2>&1; tail -5 gpurun_out/r2b_pytest_outer.log
for rows in 1; do python tools/one_outer.py 0.125 0.5 $rows; python tools/one_outer.py 0.5 0.0 $rows; python tools/one_outer.py 0.5 0.5 $rows; done > gpurun_out/r2b_outer_plain.log 2>&1
cat gpurun_out/r2b_outer_plain.log
ncu --set full --import-source on --clock-control none -k regex:outer_rows -c 1 -s 2 -o gpurun_out/r2b_outer_rows2 python tools/one_outer.py 0.125 0.5 1 > gpurun_out/r2b_outer_ncu.log 2>&1
tail -3 gpurun_out/r2b_outer_ncu.log
