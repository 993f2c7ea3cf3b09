set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_H.log 2>&1
tail -30 gpurun_out/pytest_gpu_H.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_H1.json 2> gpurun_out/bench_H1.err
cut -c1-330 gpurun_out/bench_H1.json; tail -3 gpurun_out/bench_H1.err
VM_NO_SIMPLE=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_H2.json 2> gpurun_out/bench_H2.err
cut -c1-330 gpurun_out/bench_H2.json
