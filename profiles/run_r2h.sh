# round 2h: ncu --set full of the tile-major residual kernel and of a late split_children launch (C4)
set -x
mkdir -p gpurun_out
BATCH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:cv_tile_samples -c 1 -o gpurun_out/r2h_cv_tile -f python profiles/run_full.py c4 > gpurun_out/r2h_ncu1.log 2>&1; tail -2 gpurun_out/r2h_ncu1.log
BATCH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:split_children -s 46 -c 1 -o gpurun_out/r2h_split_children -f python profiles/run_full.py c4 > gpurun_out/r2h_ncu2.log 2>&1; tail -2 gpurun_out/r2h_ncu2.log
BATCH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:walk_accumulate_fast -c 1 -o gpurun_out/r2h_walk_fast -f python profiles/run_full.py c4 > gpurun_out/r2h_ncu3.log 2>&1; tail -2 gpurun_out/r2h_ncu3.log
