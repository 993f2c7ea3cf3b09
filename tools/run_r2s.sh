timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
F="--no-graphs --no-cpu-baseline --no-e2e --no-parity --no-c5"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/r2s_launches_c4.csv python bench.py --config c4 --steps 6 --warmup 2 $F > gpurun_out/r2s_ncu_c4.log 2>&1; echo rc=$?
