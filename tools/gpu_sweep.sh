# sweep an env knob on one box: VAR, values...
VAR=$1; shift
for v in "$@"; do
  echo -n "$VAR=$v: "
  env $VAR=$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('emb/s',round(d['value']),'ms',round(d['ms_per_step'],2))"
done
