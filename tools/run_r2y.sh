timeout 600 python -m pytest tests -m gpu -x -q -k "all_reporter or dense_reporting or config4 or random or bit_identical or underflowed" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2y_pytest.log
python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-c5 > gpurun_out/r2y_c4.json 2> gpurun_out/r2y_c4.err; echo "c4 rc=$?"
python tools/show_bench.py gpurun_out/r2y_c4.json
