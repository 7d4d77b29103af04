# launch list of one step under an env setting: $1 = "VAR=value", $2 = output name
timeout 600 env $1 ncu --metrics gpu__time_duration.sum --clock-control none -s 103 -c 103 --csv --log-file gpurun_out/$2 python tools/prof_step.py --passes 2 > /dev/null 2>&1
