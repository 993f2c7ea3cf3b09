set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_I.log 2>&1
tail -4 gpurun_out/pytest_gpu_I.log
timeout 600 python bench.py > gpurun_out/bench_I.json 2> gpurun_out/bench_I.err
cut -c1-200 gpurun_out/bench_I.json; tail -2 gpurun_out/bench_I.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_I_ref.json 2> gpurun_out/bench_I_ref.err
cut -c1-200 gpurun_out/bench_I_ref.json
timeout 300 python bench.py --config c5s --no-e2e --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_I_c5s.json 2> gpurun_out/bench_I_c5s.err
cut -c1-200 gpurun_out/bench_I_c5s.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_I.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-graphs > gpurun_out/ncu_I1.log 2>&1
tail -1 gpurun_out/ncu_I1.log | cut -c1-200
timeout 400 ncu --set full --import-source on --clock-control none -k "regex:k_dense_fast|k_special" -s 8 -c 2 -o gpurun_out/prof_r1_simple -f python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --no-graphs > gpurun_out/ncu_I2.log 2>&1
tail -3 gpurun_out/ncu_I2.log | cut -c1-200
ls -la gpurun_out/prof_r1_simple.ncu-rep
