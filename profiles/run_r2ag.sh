# round 2ag: ncu --set full of one small-round launch of the CTA-per-split kernel (C4 refinement)
set -x
mkdir -p gpurun_out
BATCH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:split_children -s 20 -c 1 -o gpurun_out/r2ag_split_cta -f python profiles/run_full.py c4 > gpurun_out/r2ag_ncu.log 2>&1; tail -2 gpurun_out/r2ag_ncu.log
