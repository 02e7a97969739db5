# ncu evidence for profiles/r1d_* (end of round 1): run under gpurun, ONE GPU.
# (1) launch list of one CUDA-graph-replayed C3 step; (2) DRAM traffic of every tap-GEMM / conv1x1 launch of one step;
# (3) full captures: the dominant tap-GEMM (3x3x3 64->64 full resolution), the conv1x1 kernel (128->64 full resolution),
#     the strip-mode 7^3 stem of the super-resolution model (C4, batch 4)
N=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1d_step.csv python tools/step_once.py > gpurun_out/step_once_r1d.log 2>&1
N=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"tapgemm|conv1x1" -c 44 --csv --log-file gpurun_out/tapgemm_traffic_r1d.csv python tools/step_once.py > gpurun_out/step_once_r1d2.log 2>&1
N=1 ncu --set full --clock-control none --import-source on -k regex:tapgemm -s 1 -c 1 -o gpurun_out/prof_tapgemm_c64_r1d python tools/step_once.py > gpurun_out/step_once_r1d3.log 2>&1
N=1 ncu --set full --clock-control none --import-source on -k regex:conv1x1 -s 7 -c 1 -o gpurun_out/prof_conv1x1_r1d python tools/step_once.py > gpurun_out/step_once_r1d4.log 2>&1
N=1 B=4 CONFIG=C4 ncu --set full --clock-control none --import-source on -k regex:tapgemm -s 0 -c 1 -o gpurun_out/prof_tapgemm_stem_strips_r1d python tools/fwd_once.py > gpurun_out/step_once_r1d5.log 2>&1
tail -2 gpurun_out/step_once_r1d5.log
