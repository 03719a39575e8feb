# launch list of the bench command (short configuration), after the same command has run clean without ncu
CMD="python bench.py --steps 2 --warmup 3 --selfplay-workers 0 --scaled 0 --cpu-seconds 1"
$CMD > gpurun_out/r02_bench_short.json 2> gpurun_out/r02_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo ncu rc=$?
tail -3 gpurun_out/r02_ncu_launches.log
