export VIMURE_B200_LIB=$PWD/vimure_b200/_lib/x/libx.so
timeout 600 python -m pytest tests -m gpu -x -q -k "all_reporter or dense_reporting or config4" 2>&1 | tail -3
python bench.py --config c4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-c5 > gpurun_out/r3d_c4.json 2> gpurun_out/r3d_c4.err; echo "c4 rc=$?"
python tools/show_bench.py gpurun_out/r3d_c4.json
