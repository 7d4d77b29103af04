# tests + bench (one gpurun call)
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -15
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | tee gpurun_out/bench_quick.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('emb/s',round(d['value']),'ms',round(d['ms_per_step'],2),'convTF',round(d['roofline']['achieved'],1),'e2e',round(d['e2e']['value']))"
