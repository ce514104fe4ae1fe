timeout 600 ncu --set full --import-source on --clock-control none -k regex:mat_pipe -s 2 -c 1 -o gpurun_out/r2b_mat_pipe_step2b python tools/bench_ops.py mat > gpurun_out/r2b_mat_ncu2.log 2>&1
tail -2 gpurun_out/r2b_mat_ncu2.log
