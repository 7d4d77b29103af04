# A/B on the same box: bench with an env switch on and off, twice each
run() { env "$@" python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('emb/s',round(d['value']),'ms',round(d['ms_per_step'],2),'convTF',round(d['roofline']['achieved'],1))"; }
for i in 1 2; do echo "A: $A"; run $A; echo "B: $B"; run $B; done
