mkdir -p /tmp/tr && rm -f /tmp/tr/*
SVX_PAIR_MIN_K=380 SVX_TRACE_DIR=/tmp/tr python tools/prof_step.py --passes 1 > /dev/null 2>&1
G=$(ls /tmp/tr/*k1x1_cin768_cout1024* | head -1)
python tools/trace_report.py /tmp/tr 1 k1x1_cin768_cout1024 > gpurun_out/trace_pair.txt 2>&1
python tools/trace_abs.py $G 40 75 >> gpurun_out/trace_pair.txt 2>&1
