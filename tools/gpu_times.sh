#!/bin/bash
# per-launch conv times of the headline step (debug library) + CPU test suite is not run here
mkdir -p gpurun_out
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
SVX_CONV_TIMES=1 timeout 300 python tools/conv_times.py > gpurun_out/conv_times.txt 2>&1
tail -5 gpurun_out/conv_times.txt
