mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
( time timeout -s KILL 600 $CS --tool memcheck --print-limit 20 python tools/sanitize_target.py mc ) > gpurun_out/sanitize_memcheck_mc.log 2>&1; grep -E "ERROR SUMMARY|ok" gpurun_out/sanitize_memcheck_mc.log | tr '\n' ' '; echo
( time timeout -s KILL 600 $CS --tool racecheck --print-limit 20 python tools/sanitize_target.py mc ) > gpurun_out/sanitize_racecheck_mc.log 2>&1; grep -E "RACECHECK SUMMARY|ok" gpurun_out/sanitize_racecheck_mc.log | tr '\n' ' '; echo
( timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'^chain_kernel|classify_kernel|build_records_kernel|emit_kernel|mark_vertices' -c 6 -f -o gpurun_out/r2_prof_mc python tools/ncu_extract.py 256 ) > gpurun_out/r2_ncu_mc_full.log 2>&1; tail -2 gpurun_out/r2_ncu_mc_full.log
