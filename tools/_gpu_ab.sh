timeout 700 python -m pytest tests/test_gpu_surf.py tests/test_gpu_stereo.py tests/test_gpu_sizes.py -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
pr() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], d['value'], d['e2e']['value'], d['e2e']['sync_frame_latency_ms']); print({k:round(v['us_per_launch'],1) for k,v in d['kernels'].items() if 'surf' in k})" $1; }
timeout 300 python bench.py --steps 300 --warmup 16 --no-cpu --threshold 11032 > gpurun_out/bench_v.json 2>/dev/null; pr gpurun_out/bench_v.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tv_hyp" -s 2 -c 2 -f -o gpurun_out/r01g python tools/prof_mono.py > gpurun_out/ncu_g.log 2>&1; tail -2 gpurun_out/ncu_g.log | cut -c1-200
