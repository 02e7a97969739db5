# ncu evidence for profiles/: (1) launch list of one step, (2) DRAM traffic of every tap-GEMM launch, (3) full capture of the dominant launches
N=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1c_step.csv python tools/step_once.py > gpurun_out/step_once.log 2>&1
N=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tapgemm -c 44 --csv --log-file gpurun_out/tapgemm_traffic_r1c.csv python tools/step_once.py > gpurun_out/step_once2.log 2>&1
N=1 ncu --set full --clock-control none --import-source on -k regex:tapgemm -s 1 -c 2 -o gpurun_out/prof_tapgemm_c64_r1c python tools/step_once.py > gpurun_out/step_once3.log 2>&1
tail -2 gpurun_out/step_once3.log
