timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-scoring 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('emb/s',round(d['value']),'ms',round(d['ms_per_step'],2),'convTF',round(d['roofline']['achieved'],1))"
mkdir -p /tmp/tr && rm -f /tmp/tr/*
SVX_TRACE_DIR=/tmp/tr python tools/prof_step.py --passes 1 > /dev/null 2>&1
python tools/trace_report.py /tmp/tr 7 | awk '/trace0002|trace0005|trace0006/{f=1} f' | grep -E "==|role" -A1 | grep -v "^--"
