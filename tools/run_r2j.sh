set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 50 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 20 --warmup 3 --no-graphs --no-cpu-baseline --no-e2e --no-parity --no-c5 > gpurun_out/r2j_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_shortcut|k_dense_fast" -s 6 -c 4 -o gpurun_out/r2j_full -f python bench.py --steps 6 --warmup 3 --no-graphs --no-cpu-baseline --no-e2e --no-parity --no-c5 > gpurun_out/r2j_ncufull.log 2>&1
ls -la gpurun_out/
