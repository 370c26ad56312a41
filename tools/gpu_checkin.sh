#!/bin/bash
# One GPU call that re-establishes the measured state on a fresh B200 box (run from the repo root under gpurun):
#   gpurun --timeout 900 -- 'bash tools/gpu_checkin.sh'
# 1. the parity suite (-rxX lists the JPEG group, which is non-strict xfail until its first pass on hardware);
# 2. smoke();  3. the bench line;  4. per-kernel numbers of the JPEG ingest and of the 64- / 128-float descriptor rows;
# 5. the ncu launch list of the bench and one --set full capture of the two JPEG kernels.
# Everything lands in gpurun_out/ (scratch); copy what is to be judged into profiles/.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rxX -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --steps 300 --warmup 16 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
python tools/jpeg_probe.py > gpurun_out/jpeg_probe.json 2> gpurun_out/jpeg_probe.err; cat gpurun_out/jpeg_probe.json
python tools/ext_probe.py > gpurun_out/ext_probe.json 2> gpurun_out/ext_probe.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/ncu_launches.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_jpeg_idct|k_jpeg_color" -c 6 -o gpurun_out/jpeg \
    python tools/jpeg_probe.py > gpurun_out/ncu_jpeg.log 2>&1
ls -la gpurun_out | tail -12
