# round-2 evidence: launch lists with DRAM bytes (extraction step, AS-norm job) and ncu --set full captures of the two top kernels
set -x
python tools/prof_step.py --passes 2 > gpurun_out/r02_plain_step.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02_launches_step.csv python tools/prof_step.py --passes 2 > gpurun_out/r02_ncu_step.log 2>&1; tail -1 gpurun_out/r02_ncu_step.log
python tools/prof_score.py 2 > gpurun_out/r02_plain_score.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
  --log-file gpurun_out/r02_launches_score.csv python tools/prof_score.py 2 > gpurun_out/r02_ncu_score.log 2>&1; tail -1 gpurun_out/r02_ncu_score.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:asnorm_fused -s 1 -c 1 -o gpurun_out/r02_asnorm_fused -f \
  python tools/prof_score.py 2 > gpurun_out/r02_ncu_full_score.log 2>&1; tail -2 gpurun_out/r02_ncu_full_score.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 130 -c 5 -o gpurun_out/r02_flat_stage3 -f \
  python tools/prof_step.py --passes 2 > gpurun_out/r02_ncu_full_step.log 2>&1; tail -2 gpurun_out/r02_ncu_full_step.log
ls -la gpurun_out/r02_*
