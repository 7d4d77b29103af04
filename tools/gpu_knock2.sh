# Knock-out experiment over several layers: steady ns/unit of the MMA role, full pipeline (0), MMA+loads only (15), issue loop only (47).
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_knock.so
for K in 0 15 47; do
  rm -rf /tmp/tr && mkdir -p /tmp/tr
  SVX_FLAT_KNOCK=$K SVX_TRACE_DIR=/tmp/tr python tools/prof_step.py --passes 1 > /dev/null 2>&1
  echo "=== knock $K"
  for PAT in k1x1_cin32_cout96 k1x1_cin96_cout128 k1x1_cin128_cout96 k3x3_cin24_cout24_s1_aux2 k3x3_cin48_cout48_s1_aux2 k1x1_cin192_cout256 k1x1_cin256_cout192 k3x3_cin96_cout96_s1_aux2 k1x1_cin384_cout512 k1x1_cin512_cout384 k3x3_cin192_cout192_s1_aux2 k1x1_cin768_cout1024 k1x1_cin1024_cout768; do
    python tools/trace_report.py /tmp/tr 1 $PAT 2>&1 | grep -A2 "^==\|role 1" | grep "^==\|role 1\|accfree" | head -3
  done
done
