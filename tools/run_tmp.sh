timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err; echo rc=$?
python tools/show_bench.py gpurun_out/r3k_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
