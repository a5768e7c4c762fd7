mkdir -p gpurun_out
( timeout -s KILL 200 python -m pytest tests/test_gpu_unet.py -x -q -k "large_batches" ) > gpurun_out/pytest_large.log 2>&1; tail -15 gpurun_out/pytest_large.log
