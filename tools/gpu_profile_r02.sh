#!/bin/bash
# Round-2 profiling call (run under gpurun from the repo root): the ncu launch list of the bench command and
# `--set full` captures of the kernels that matter.  The .ncu-rep files stay on the GPU box (they exceed what
# gpurun copies back); their summaries -- tools/ncu_summary.py, tools/ncu_source_lines.py -- land in gpurun_out/ and
# are copied to profiles/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile_r02.sh'
set -u
mkdir -p gpurun_out
R=/tmp/uvo_ncu
mkdir -p $R
# launch list (cold-cache, serialised): kernels are launched directly (--graphs 0) so that every launch is listed
ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 400 --csv --log-file gpurun_out/ncu_launches_r02.csv \
    python bench.py --steps 8 --regions 1 --warmup 3 --no-cpu --graphs 0 --jpeg-threads 0 > gpurun_out/ncu_l_r02.log 2>&1
# full capture of the front-end, matcher and pose kernels of the stereo frame (two frames' worth of launches)
ncu --set full --clock-control none --import-source on \
    -k regex:"k_surf_detect|k_surf_patch|k_surf_vector|k_surf_sort_block|k_knn_tc|k_knn_rerank|k_pnp_chunk|k_pnp_finalize|k_gray_undistort|k_clahe_tile_lut|k_clahe_apply|k_integral_rows|k_integral_cols" \
    -s 1000 -c 40 -o $R/r02_frame python bench.py --steps 4 --regions 1 --warmup 3 --no-cpu --graphs 0 --jpeg-threads 0 \
    > gpurun_out/ncu_f_r02.log 2>&1
python tools/ncu_summary.py $R/r02_frame.ncu-rep gpurun_out/ncu_full_r02_summary.json gpurun_out/ncu_traffic_r02.json
for k in k_surf_detect k_surf_patch k_pnp_chunk k_pnp_finalize k_clahe_tile_lut k_integral_cols; do
  python tools/ncu_source_lines.py $R/r02_frame.ncu-rep $k 1.5 --by-samples > gpurun_out/ncu_lines_r02_$k.txt 2>&1
done
# the JPEG ingest kernels (GPU Huffman decode, IDCT, colour) on a 1280x1024 frame
ncu --set full --clock-control none --import-source on -k regex:"k_jpeg_huff|k_jpeg_idct|k_jpeg_color" -s 9 -c 6 \
    -o $R/r02_jpeg python tools/jh_time.py > gpurun_out/ncu_j_r02.log 2>&1
python tools/ncu_summary.py $R/r02_jpeg.ncu-rep gpurun_out/ncu_full_r02_jpeg_summary.json gpurun_out/ncu_traffic_r02_jpeg.json
python tools/ncu_source_lines.py $R/r02_jpeg.ncu-rep k_jpeg_huff 1.5 --by-samples > gpurun_out/ncu_lines_r02_k_jpeg_huff.txt 2>&1
ls -la gpurun_out | grep r02 | tail -20
