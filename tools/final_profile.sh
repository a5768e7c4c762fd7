#!/bin/bash
# Round-end profiling evidence (run under gpurun, one GPU): launch lists + ncu --set full captures of the dominant kernels +
# compute-sanitizer.  Raw artefacts land in gpurun_out/; tools/summarize_profiles.py turns them into profiles/*.
mkdir -p gpurun_out
TAG=${1:-r2}
# (1) launch list: one shape's extraction at 512^3 (lattice -> marching cubes -> face filter)
( timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_extract512_launches.csv python tools/ncu_extract.py 512 ) > gpurun_out/${TAG}_ncu_extract.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_extract.log
# (2) launch list: the persistent sampler (8 DDPM steps, batch 8)
( timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_sampler_launches.csv python tools/ncu_persist.py 8 8 ) > gpurun_out/${TAG}_ncu_sampler_l.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_sampler_l.log
# (3) --set full: persistent sampler (one 8-step launch)
( timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:unet_persistent -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_persist python tools/ncu_persist.py 8 8 ) > gpurun_out/${TAG}_ncu_persist.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_persist.log
# (4) --set full: decoder layer chain, MC chain, classify, records, emit (256^3 extraction)
( timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'tc_chain_kernel|chain_kernel|classify_kernel|build_records_kernel|emit_kernel' -c 12 -f -o gpurun_out/${TAG}_prof_extract python tools/ncu_extract.py 256 ) > gpurun_out/${TAG}_ncu_extract_full.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_extract_full.log
ls -la gpurun_out/${TAG}_*.ncu-rep gpurun_out/${TAG}_*.csv
# (5) compute-sanitizer
bash tools/sanitize.sh
