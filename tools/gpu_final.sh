# round-end evidence in one gpurun call: GPU tests, full bench line, launch list with DRAM bytes, plans
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 600 gpurun_out/bench_final.json
bash tools/gpu_prof.sh
