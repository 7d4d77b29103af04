# bench under several env settings on the same box: ENVS="A=1 B=1 ..." (each run twice, interleaved)
for i in 1 2; do for e in $ENVS; do echo -n "$e: "; env $e python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2))
except Exception as ex: print('failed', ex)"; done; done
