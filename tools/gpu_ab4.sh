# like gpu_ab3.sh but every setting is a comma-separated list of VAR=value pairs: ENVS="A=1,B=2 C=3"
for i in 1 2; do for e in $ENVS; do echo -n "$e: "; env $(echo $e | tr ',' ' ') python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2))
except Exception as ex: print('failed', ex)"; done; done
