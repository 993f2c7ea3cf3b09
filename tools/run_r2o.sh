export VIMURE_B200_LIB=$PWD/vimure_b200/_lib/x/libx.so
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-c5"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2o_$name.json 2> gpurun_out/r2o_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2o_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step=%.4f dense_ms=%.4f elbo=%r' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['elbo_final']))
except Exception as e:
    print('$name', 'ERR', e)
PY
}
run pd2 X=0
run prologue_only VM_X_TMA_MODE=2
run loads_no_sts VM_X_TMA_MODE=5
