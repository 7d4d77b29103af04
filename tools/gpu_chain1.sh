#!/bin/bash
# first runs of the fused chain kernel: debug library (2 s barrier watchdog), small case first
mkdir -p gpurun_out
export SVX_LIB=$PWD/voxsrc2020_speaker_verification_b200/libsvx_dbg.so
timeout 300 python -m pytest tests/test_gpu_extract.py -m gpu -q --no-header -p no:cacheprovider -x -k "fused_chain" 2>&1 | tail -30 > gpurun_out/chain1.log
unset SVX_LIB
timeout 600 python -m pytest tests/test_gpu_extract.py -m gpu -q --no-header -p no:cacheprovider -x -k "fused_chain or (segments_match and res2net50_w24_s4_c32 and tensor and not 2d)" 2>&1 | tail -15 >> gpurun_out/chain1.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | tail -1 > gpurun_out/chain1_bench.json
cat gpurun_out/chain1.log
python -c "
import json
d=json.loads(open('gpurun_out/chain1_bench.json').read().strip().splitlines()[-1]);print('emb/s',round(d['value']),'ms',round(d['ms_per_step'],2),'convTF',round(d['roofline']['achieved'],1))"
