export VIMURE_B200_LIB=$PWD/vimure_b200/_lib/x/libx.so
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-c5"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2u_$name.json 2> gpurun_out/r2u_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2u_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step=%.4f' % d['ms_per_step'])
    print('   e2e', d['e2e']['wall_s_all_runs'], d['e2e']['timings_s'])
except Exception as e:
    print('$name', 'ERR', e); print(open('gpurun_out/r2u_$name.err').read()[-1500:])
PY
}
run graphs VM_X_SPLIT=1
B="$B --no-graphs"
run nographs VM_X_SPLIT=1
