mkdir -p /tmp/tr && rm -f /tmp/tr/*
SVX_TRACE_DIR=/tmp/tr python tools/prof_step.py --passes 1 > /dev/null 2>&1
ls /tmp/tr | head -40 > gpurun_out/trace_files.txt
python tools/trace_report.py /tmp/tr 12 flat > gpurun_out/trace_report.txt 2>&1
python tools/trace_abs.py /tmp/tr/trace0002*.bin 150 165 > gpurun_out/trace_abs_3x3.txt 2>&1
