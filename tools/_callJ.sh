set -x
timeout 200 python -m pytest tests/test_gpu_scaled.py -m gpu -x -q -k "sharded" > gpurun_out/pytest_gpu_J.log 2>&1
tail -12 gpurun_out/pytest_gpu_J.log
timeout 400 python bench.py > gpurun_out/bench_J.json 2> gpurun_out/bench_J.err
cut -c1-200 gpurun_out/bench_J.json; tail -2 gpurun_out/bench_J.err
