# forced multicast cluster shapes, per-layer times of the deep-stage 1x1 / 3x3 convs (debug library)
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
for cfg in "1 1" "2 1" "4 1" "1 2" "1 4" "2 2" "4 2" "2 4"; do
  set -- $cfg
  f=gpurun_out/mc_$1x$2.txt
  SVX_MC_CN=$1 SVX_MC_CM=$2 SVX_CONV_TIMES=1 SVX_PLAN_LOG=1 timeout 120 python tools/conv_times.py > $f 2>&1
  echo "cn=$1 cm=$2: $(grep -E '^step' $f)"
done
