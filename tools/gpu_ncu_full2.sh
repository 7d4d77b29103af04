# ncu --set full of (a) stage-1 block 1: the three hierarchical 3x3 convs (hybrid, hybrid, direct) and (b) the last two flat launches (stage 4)
# (two small reports: gpurun_out/ may carry at most 64 MiB back)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 79 -c 3 -o gpurun_out/r01_v5_stage1 -f python tools/prof_step.py --passes 2 > gpurun_out/ncu_full_a.log 2>&1; tail -1 gpurun_out/ncu_full_a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_flat -s 142 -c 2 -o gpurun_out/r01_v5_stage4 -f python tools/prof_step.py --passes 2 > gpurun_out/ncu_full_b.log 2>&1; tail -1 gpurun_out/ncu_full_b.log
ls -la gpurun_out/*.ncu-rep
