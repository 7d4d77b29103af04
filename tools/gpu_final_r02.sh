#!/bin/bash
# round-2 final evidence in one gpurun call: GPU tests, full bench line, per-launch times, ncu launch list (time + DRAM bytes),
# ncu --set full of the pair kernel (second pass: stride-2 shortcut, 3x3 stride 2, 1x1 with residual, ...) and of the fused chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r02_gputest_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 400 gpurun_out/r02_bench_final.json; echo
SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so SVX_CONV_TIMES=1 timeout 300 python tools/conv_times.py > gpurun_out/r02_conv_times_final.txt 2>&1; tail -2 gpurun_out/r02_conv_times_final.txt
python tools/prof_step.py --passes 2 > gpurun_out/r02_plain_step.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02_launches_step.csv python tools/prof_step.py --passes 2 > gpurun_out/r02_ncu_step.log 2>&1; tail -1 gpurun_out/r02_ncu_step.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 49 -c 8 -o gpurun_out/r02_pair_final -f \
  python tools/prof_step.py --passes 2 > gpurun_out/r02_ncu_full_pair.log 2>&1; tail -2 gpurun_out/r02_ncu_full_pair.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:res2_chain -s 3 -c 1 -o gpurun_out/r02_chain_final -f \
  python tools/prof_step.py --passes 2 > gpurun_out/r02_ncu_full_chain.log 2>&1; tail -2 gpurun_out/r02_ncu_full_chain.log
ls -la gpurun_out/r02_*final*
