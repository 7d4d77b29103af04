timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -8
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"])
a=d["asnorm"]; print("asnorm rows-sharded", a["value"], a["ms_per_job"], "cohort-sharded", a["cohort_rows_sharded"]["value"], a["cohort_rows_sharded"]["ms_per_job"])
for k,v in d["configs"].items(): print(k, v["value"], v.get("frames_per_s"), v["frac_of_sustained_peak"])
PY
