timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3e_pytest.log
python bench.py --steps 50 --warmup 5 > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/r3e_bench.json
python bench.py --config c4 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-c5 > gpurun_out/r3e_c4.json 2> gpurun_out/r3e_c4.err; echo "c4 rc=$?"
python tools/show_bench.py gpurun_out/r3e_c4.json
