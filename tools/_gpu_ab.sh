timeout 900 python -m pytest tests/test_gpu_imgprep.py tests/test_gpu_stereo.py tests/test_gpu_sizes.py tests/test_gpu_mono.py -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
pr() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['e2e']['value'], d['e2e']['sync_frame_latency_ms']); print({k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if 'gray' in k or 'clahe' in k})" $1; }
timeout 300 python bench.py --steps 300 --warmup 16 --no-cpu --threshold 11032 > gpurun_out/bench_v.json 2>/dev/null; pr gpurun_out/bench_v.json
