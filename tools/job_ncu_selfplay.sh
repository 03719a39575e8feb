# launch lists of the two self-play paths (host driver on the evaluation service; GPU-resident search)
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log
CMD="python tools/ncu_selfplay_target.py 300"
$CMD > gpurun_out/r02_selfplay_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/r02_launches_selfplay.csv $CMD > gpurun_out/r02_ncu_selfplay.log 2>&1
echo ncu rc=$?
CMD2="python tools/ncu_gpu_selfplay_target.py"
$CMD2 > gpurun_out/r02_gpu_selfplay_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 120 --csv --log-file gpurun_out/r02_launches_gpu_selfplay.csv $CMD2 > gpurun_out/r02_ncu_gpu_selfplay.log 2>&1
echo ncu rc=$?
