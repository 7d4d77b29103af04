#!/bin/bash
# A/B of several builds of the library on the same box: bash tools/gpu_lib_ab.sh libsvx_v1.so libsvx_v2a.so libsvx.so  (two rounds each)
mkdir -p gpurun_out
for round in 1 2; do
for l in "$@"; do
  SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/$l timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-scoring --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$l', 'emb/s',round(d['value']),'ms',round(d['ms_per_step'],3),'convTF',round(d['roofline']['achieved'],1), d['clocks'])"
done
done
