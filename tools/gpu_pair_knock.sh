#!/bin/bash
# knock-out timings of the CTA-pair GEMM (debug library): which part of the tile time is the epilogue's
mkdir -p gpurun_out
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
for k in ${KLIST:-0 1 2 4 6}; do
  SVX_PAIR_KNOCK=$k SVX_CONV_TIMES=1 SVX_PLAN_LOG=1 timeout 300 python tools/conv_times.py > gpurun_out/pair_knock_$k.txt 2>&1
  echo "== knock $k: $(grep -E '^step' gpurun_out/pair_knock_$k.txt) $(grep 'resident CTA pairs' gpurun_out/pair_knock_$k.txt | head -1)"
  grep pair gpurun_out/pair_knock_$k.txt | grep convtime | awk '{k=$9"->"$11" aux "$13; t[k]+=$3; n[k]++} END{for (k in t) printf "   %-24s %.1f us x%d\n", k, t[k]/n[k], n[k]}' | sort
done
