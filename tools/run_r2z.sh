timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-c5 > gpurun_out/r2z_c4.json 2> gpurun_out/r2z_c4.err; echo "c4 rc=$?"
python tools/show_bench.py gpurun_out/r2z_c4.json
F="--no-graphs --no-cpu-baseline --no-e2e --no-parity --no-c5"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/r2z_launches_c4.csv python bench.py --config c4 --steps 6 --warmup 2 $F > gpurun_out/r2z_ncu_c4.log 2>&1; echo rc=$?
