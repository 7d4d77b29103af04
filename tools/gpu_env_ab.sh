#!/bin/bash
# A/B of debug-library environment knobs on the headline step: bash tools/gpu_env_ab.sh "VAR=val VAR2=val" "..." ; prints step time + 3x3 s1 conv times
mkdir -p gpurun_out
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
i=0
for cfg in "$@"; do
  f=gpurun_out/env_ab_$i.txt; i=$((i+1))
  env $cfg SVX_CONV_TIMES=1 timeout 200 python tools/conv_times.py > $f 2>&1
  echo "== [$cfg] $(grep -E '^step' $f) $(grep -E '^conv ' $f)"
  grep convtime $f | awk '{k=$5" "$6" "$9"->"$11" aux "$13" "$14" b_st "$26; t[k]+=$3; n[k]++} END{for (k in t) printf "   %-44s %.1f us x%d\n", k, t[k]/n[k], n[k]}' | sort | grep -E "${FILTER:-3x3 s1}"
done
