mkdir -p /tmp/tr && rm -f /tmp/tr/*
SVX_TRACE_DIR=/tmp/tr python tools/prof_step.py --passes 1 > /dev/null 2>&1
ls /tmp/tr > gpurun_out/trace_files.txt
F=$(ls /tmp/tr/*k3x3_cin96_cout96_s1_aux2* | head -2 | tail -1)
python tools/trace_report.py /tmp/tr 3 k3x3_cin96_cout96_s1_aux2 > gpurun_out/trace_s3.txt 2>&1
python tools/trace_abs.py $F 40 75 > gpurun_out/trace_abs_s3.txt 2>&1
G=$(ls /tmp/tr/*k1x1_cin384_cout512* | head -2 | tail -1)
python tools/trace_report.py /tmp/tr 2 k1x1_cin384_cout512 > gpurun_out/trace_s3_conv3.txt 2>&1
python tools/trace_abs.py $G 40 70 > gpurun_out/trace_abs_s3_conv3.txt 2>&1
