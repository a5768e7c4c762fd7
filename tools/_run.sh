mkdir -p gpurun_out
SURFD_UNET_DEBUG=64 timeout -s KILL 120 python tools/sampler_profile.py 2>&1 | grep -v Warning | tail -6 > gpurun_out/sampler_attn_defer.txt 2>&1
cat gpurun_out/sampler_attn_defer.txt
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_extract.csv python tools/ncu_extract.py 256 > gpurun_out/ncu_extract.log 2>&1
tail -2 gpurun_out/ncu_extract.log | cut -c1-300; wc -l gpurun_out/launches_extract.csv
