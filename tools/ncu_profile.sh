python tools/ncu_target.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01b.csv python tools/ncu_target.py > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tower_kernel -s 3 -c 2 -o gpurun_out/prof_tower_r01b python tools/ncu_target.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
