#!/bin/bash
# tail kernels (stem, avg pool, statistics pooling, FC): extraction tests, step time, ncu launch list of one step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_dropin.py -m gpu -q -x --no-header -p no:cacheprovider 2>&1 | tail -3
python tools/prof_step.py --passes 3 2>&1 | tail -3
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/tail_launches.csv python tools/prof_step.py --passes 2 > gpurun_out/tail_ncu.log 2>&1; tail -1 gpurun_out/tail_ncu.log
