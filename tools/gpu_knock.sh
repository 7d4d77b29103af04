# Knock-out experiment: role timelines of one layer ($1 = trace-name pattern) with parts of the pipeline switched off.
PAT=${1:-k3x3_cin24_cout24_s1_aux2}
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_knock.so
for K in ${KLIST:-0 1 2 4 9 6 15}; do
  rm -rf /tmp/tr && mkdir -p /tmp/tr
  SVX_FLAT_KNOCK=$K SVX_TRACE_DIR=/tmp/tr python tools/prof_step.py --passes 1 > /dev/null 2>&1
  echo "=== knock $K"
  python tools/trace_report.py /tmp/tr 1 $PAT 2>&1 | head -24
done
