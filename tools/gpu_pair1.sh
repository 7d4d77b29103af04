#!/bin/bash
# first runs of the CTA-pair GEMM: debug library (2 s barrier watchdog), small case first, then the production library + per-launch times
mkdir -p gpurun_out
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
timeout 300 python -m pytest tests/test_gpu_extract.py -m gpu -q --no-header -p no:cacheprovider -x -k "pair_gemm" 2>&1 | tail -30 > gpurun_out/pair1.log
SVX_CONV_TIMES=1 timeout 300 python tools/conv_times.py > gpurun_out/conv_times_pair.txt 2>&1
unset SVX_LIB
timeout 600 python -m pytest tests/test_gpu_extract.py -m gpu -q --no-header -p no:cacheprovider -x -k "pair_gemm or segments_match" 2>&1 | tail -15 >> gpurun_out/pair1.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | tail -1 > gpurun_out/pair1_bench.json
cat gpurun_out/pair1.log
grep -E "pair|^step|^conv " gpurun_out/conv_times_pair.txt | cut -c1-150
python -c "
import json
d=json.loads(open('gpurun_out/pair1_bench.json').read().strip().splitlines()[-1]);print('emb/s',round(d['value']),'ms',round(d['ms_per_step'],2),'convTF',round(d['roofline']['achieved'],1))"
