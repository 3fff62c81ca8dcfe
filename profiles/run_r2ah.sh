# round 2ah: launch list of the C3 step (sum of kernel durations against the step time: is the stream ever idle between kernels?)
set -x
mkdir -p gpurun_out
BATCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2ah_launches_c3.csv python profiles/run_full.py c3 > gpurun_out/r2ah_c3_run.log 2>&1; tail -3 gpurun_out/r2ah_c3_run.log
python profiles/summarize_launches.py gpurun_out/r2ah_launches_c3.csv
