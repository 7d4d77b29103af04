# per-launch table of one headline step (ncu launch list) + the chosen plans
SVX_PLAN_LOG=1 python tools/prof_step.py --passes 2 2>&1 | grep "^plan" | sort | uniq -c > gpurun_out/plans.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 103 -c 103 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --passes 2 > gpurun_out/ncu.log 2>&1; tail -1 gpurun_out/ncu.log
