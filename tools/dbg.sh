T="tests/test_gpu_extract.py::test_segments_match_oracle[umma-res2net50_w24_s4_c32-80-lens2]"
for v in 32 48 1; do
  echo "== mask $v"; SVX_STAGED_MASK=$v timeout 300 python -m pytest "$T" -q --no-header -p no:cacheprovider 2>&1 | grep -E "AssertionError: \(|passed|failed" | head -3
done
