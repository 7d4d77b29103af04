#!/bin/bash
# round-2 closing evidence in one gpurun call: GPU tests, full bench line, per-launch conv times, ncu launch lists (time + DRAM bytes) of
# one extraction step and one AS-norm job, ncu --set full of the fused AS-norm kernel, the pair kernel and the fused chain
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r02b_gputest.log
timeout 900 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 600 gpurun_out/r02b_bench.json; echo
SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so SVX_CONV_TIMES=1 timeout 300 python tools/conv_times.py > gpurun_out/r02b_conv_times.txt 2>&1; tail -2 gpurun_out/r02b_conv_times.txt
python tools/prof_step.py --passes 2 > gpurun_out/r02b_plain_step.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02b_launches_step.csv python tools/prof_step.py --passes 2 > gpurun_out/r02b_ncu_step.log 2>&1; tail -1 gpurun_out/r02b_ncu_step.log
python tools/prof_score.py 2 > gpurun_out/r02b_plain_score.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
  --log-file gpurun_out/r02b_launches_score.csv python tools/prof_score.py 2 > gpurun_out/r02b_ncu_score.log 2>&1; tail -1 gpurun_out/r02b_ncu_score.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:asnorm_fused -s 1 -c 1 -o gpurun_out/r02b_asnorm_fused -f \
  python tools/prof_score.py 2 > gpurun_out/r02b_ncu_full_score.log 2>&1; tail -2 gpurun_out/r02b_ncu_full_score.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_conv|res2_chain|avgpool|fc_kernel|stats_pool" -s 9 -c 9 -o gpurun_out/r02b_tail_chain -f \
  python tools/prof_step.py --passes 2 > gpurun_out/r02b_ncu_full_tail.log 2>&1; tail -2 gpurun_out/r02b_ncu_full_tail.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 50 -c 8 -o gpurun_out/r02b_pair -f \
  python tools/prof_step.py --passes 2 > gpurun_out/r02b_ncu_full_pair.log 2>&1; tail -2 gpurun_out/r02b_ncu_full_pair.log
ls -la gpurun_out/r02b_*
