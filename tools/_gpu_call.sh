timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest_all.log 2>&1; tail -5 gpurun_out/r2c_pytest_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
