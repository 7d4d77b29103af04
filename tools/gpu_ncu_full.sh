# one ncu --set full capture of a few conv_flat launches of the headline step (stage-1 block 2: conv1, 3x3 x3, conv3)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 65 -c 6 -o gpurun_out/r01_flat_full -f python tools/prof_step.py --passes 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
