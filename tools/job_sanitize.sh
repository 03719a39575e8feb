set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest0.log 2>&1; tail -3 gpurun_out/r02_pytest0.log
python tools/host_probe.py > gpurun_out/r02_host_probe.log 2>&1; tail -30 gpurun_out/r02_host_probe.log
for tool in memcheck racecheck synccheck; do
  for mode in tower128 tower256 legal selfplay; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $mode > gpurun_out/r02_sanitizer_${tool}_${mode}.log 2>&1
    echo "$tool $mode rc=$?"; tail -4 gpurun_out/r02_sanitizer_${tool}_${mode}.log
  done
done
