# round 2ab: ncu --set full of the shipped tile-major residual kernel (C4), source-level hot spots
set -x
mkdir -p gpurun_out
BATCH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:cv_tile_samples -c 1 -o gpurun_out/r2ab_cv_tile -f python profiles/run_full.py c4 > gpurun_out/r2ab_ncu1.log 2>&1; tail -2 gpurun_out/r2ab_ncu1.log
