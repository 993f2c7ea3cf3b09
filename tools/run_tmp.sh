export VIMURE_B200_LIB=$PWD/vimure_b200/_lib/x/libx.so
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r3m_bench.json 2> gpurun_out/r3m_bench.err; echo rc=$?
python tools/show_bench.py gpurun_out/r3m_bench.json
python -c "
import json; d=json.loads(open('gpurun_out/r3m_bench.json').read().strip().splitlines()[-1]); print('parity', d['parity']['pass'])"
