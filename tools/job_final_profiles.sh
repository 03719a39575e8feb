# round-2 evidence with the final kernels: parity report, full ncu capture of the tower kernel, launch list of the bench command
python -m pytest tests -m gpu -q -s -k "predict_matches or headline or trained_like or large_batch or golden or trunk" 2>&1 | grep -E "^\.?\[parity|passed|failed" > gpurun_out/r02_parity_report.txt
tail -1 gpurun_out/r02_parity_report.txt
CMD="python tools/ncu_target.py 1024"
$CMD > gpurun_out/r02_ncu_target_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tower_kernel -s 3 -c 2 -f -o gpurun_out/r02_prof_tower_1024 $CMD > gpurun_out/r02_ncu_tower.log 2>&1
echo ncu full rc=$?
