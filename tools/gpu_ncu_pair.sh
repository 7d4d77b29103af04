# ncu --set full of three CTA-pair GEMM launches of the second pass (stage 3: 384->512 + residual, 512->384, ...)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s ${SKIP:-20} -c ${COUNT:-3} -o gpurun_out/r02_pair -f python tools/prof_step.py --passes 2 > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log
ls -la gpurun_out/r02_pair.ncu-rep
