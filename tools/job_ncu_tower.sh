# full capture of the tower kernel at batch 1024 (device-resident forward), after the same command has run clean
CMD="python tools/ncu_target.py 1024"
$CMD > gpurun_out/r02_ncu_target_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tower_kernel -s 3 -c 2 -f -o gpurun_out/r02_prof_tower_1024 $CMD > gpurun_out/r02_ncu_tower.log 2>&1
echo ncu rc=$?
tail -3 gpurun_out/r02_ncu_tower.log
