timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo rc=$? >> gpurun_out/t_gpu.log; tail -3 gpurun_out/t_gpu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:prune4_reg -s 7 -c 7 -f -o gpurun_out/r1_reg_c2 python tools/profile_step.py --config c2 --evals 2 > gpurun_out/ncu_reg.log 2>&1; echo ncu rc=$?
