timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_configs.py -x -q -m gpu -k "matrix or config4" > gpurun_out/r2b_pytest_mat.log 2>&1; tail -2 gpurun_out/r2b_pytest_mat.log
timeout 300 python tools/bench_ops.py mat > gpurun_out/r2d_mat_plain.log 2>&1; tail -1 gpurun_out/r2d_mat_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 15 --csv --log-file gpurun_out/r2f_mat_launches.csv python tools/bench_ops.py mat > gpurun_out/r2b_mat_ncu.log 2>&1
