#!/bin/bash
# first runs of the pair kernel's stride-2 tile mode: debug library (2 s watchdog) first, then production + per-launch times
mkdir -p gpurun_out
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
timeout 300 python -m pytest tests/test_gpu_extract.py -m gpu -q --no-header -p no:cacheprovider -x -k "stride2" 2>&1 | tail -30 > gpurun_out/s2.log
SVX_CONV_TIMES=1 timeout 300 python tools/conv_times.py > gpurun_out/conv_times_s2.txt 2>&1
unset SVX_LIB
timeout 600 python -m pytest tests/test_gpu_extract.py -m gpu -q --no-header -p no:cacheprovider -x -k "pair or segments_match" 2>&1 | tail -15 >> gpurun_out/s2.log
cat gpurun_out/s2.log
grep -E " s2 |^step|^conv " gpurun_out/conv_times_s2.txt | cut -c1-150
