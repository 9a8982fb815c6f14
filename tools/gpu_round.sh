timeout 600 python tools/segtime.py 10000 50 > gpurun_out/segtime.log 2>&1
