timeout 700 python -m pytest tests/test_gpu_twoview.py tests/test_gpu_mono.py -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_mono.py 2>&1 | head -8
timeout 300 python tools/microbench.py --only D 2>&1 | grep -A12 gpu_ms
