timeout 600 python tools/shard_times_c3.py 8 > gpurun_out/r2f_c3_shards.log 2>&1; cat gpurun_out/r2f_c3_shards.log
