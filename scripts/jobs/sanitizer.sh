mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck.log 2>&1
tail -12 gpurun_out/sanitizer_memcheck.log
CFN_RAYS=64 CFN_STEPS=1 $T 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/r2_step512.py > gpurun_out/sanitizer_racecheck.log 2>&1
tail -8 gpurun_out/sanitizer_racecheck.log
