timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 50 --warmup 5 > gpurun_out/r3p_bench.json 2> gpurun_out/r3p_bench.err; echo rc=$?
python tools/show_bench.py gpurun_out/r3p_bench.json
