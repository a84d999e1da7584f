mkdir -p gpurun_out/tb2
ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:traceback_ckpt_tasks_kernelILi8 -s 2 -c 1 -o gpurun_out/tb2/tbckpt2 python tools/stage_times.py 16384 --short > gpurun_out/tb2/log.txt 2>&1
tail -3 gpurun_out/tb2/log.txt
