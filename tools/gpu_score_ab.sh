#!/bin/bash
# AS-norm kernel: scoring tests, job time (production library), then knock-outs of the pass-0 / pass-1 epilogues (debug library).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_score.py -x -q -m gpu 2>&1 | tail -3
python tools/score_bench.py 2>&1 | tee gpurun_out/score_bench.txt
for k in ${KLIST:-0 8 16 24 2}; do
  echo "== SVX_ASNORM_KNOCK=$k"
  SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so SVX_ASNORM_KNOCK=$k python tools/prof_score.py 4 2>&1 | tail -2
done 2>&1 | tee gpurun_out/score_knock.txt
