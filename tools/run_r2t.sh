export VIMURE_B200_LIB=$PWD/vimure_b200/_lib/x/libx.so
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-c5"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2t_$name.json 2> gpurun_out/r2t_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2t_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step=%.4f dense_ms=%.4f elbo=%r' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['elbo_final']))
    print('   e2e', d['e2e']['wall_s_all_runs'], d['e2e']['timings_s'])
except Exception as e:
    print('$name', 'ERR', e); print(open('gpurun_out/r2t_$name.err').read()[-1500:])
PY
}
run sums_aux X=0
run split VM_X_SPLIT=1
run sums_aux2 X=0
