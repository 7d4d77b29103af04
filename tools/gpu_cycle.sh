# tests + bench + per-launch table (one gpurun call)
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scoring > gpurun_out/bench.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('emb/s',d['value'],'ms',d['ms_per_step'],'conv TF/s',d['roofline']['achieved'],'e2e',d['e2e']['value'])"
python tools/prof_step.py --passes 2 > gpurun_out/prof_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 95 -c 95 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --passes 2 > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
