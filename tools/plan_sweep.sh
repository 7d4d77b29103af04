# forced tile shapes for one layer class (debug library): bash tools/plan_sweep.sh "9,96,96" "32,4,1" "32,2,1" "96,2,0" ...
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
L=$1; shift
for cfg in "$@"; do
  f=gpurun_out/plan_${L//,/_}_${cfg//,/_}.txt
  SVX_FORCE_PLAN="$L,$cfg" SVX_CONV_TIMES=1 timeout 120 python tools/conv_times.py > $f 2>&1
  IFS=, read taps cin cout <<< "$L"
  echo "$L -> $cfg: $(grep -E '^step' $f)  layer: $(grep convtime $f | grep "${taps:0:1}x" | awk -v ci=$cin -v co=$cout '$9==ci && $11==co && $7=="s1"{t+=$3;n++} END{if(n) printf "%.1f us x%d", t/n, n}')"
done
