mkdir -p gpurun_out
timeout -s KILL 300 python tools/pipeline_probe.py 256 2>&1 | grep -v Warning > gpurun_out/pipeline_probe.txt
grep -E "iter2|iter1" -A 9 gpurun_out/pipeline_probe.txt | tail -34
tail -8 gpurun_out/pipeline_probe.txt
