export VIMURE_B200_LIB=$PWD/vimure_b200/_lib/x/libx.so
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-c5"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2m_$name.json').read().strip().splitlines()[-1])
    print('$name', 'ms_per_step=%.4f dense_ms=%.4f elbo=%r' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['elbo_final']))
except Exception as e:
    print('$name', 'ERR', e)
PY
}
run base X=0
run nopatch VM_X_NOPATCH=1
run nopatch_ov148 VM_X_NOPATCH=1 VM_X_OVERLAP=1 VM_X_PERSIST=148
run nopatch_ov296 VM_X_NOPATCH=1 VM_X_OVERLAP=1 VM_X_PERSIST=296
run nopatch_ov444 VM_X_NOPATCH=1 VM_X_OVERLAP=1 VM_X_PERSIST=444
run nopatch_ov2_296 VM_X_NOPATCH=1 VM_X_OVERLAP=2 VM_X_PERSIST=296
run patch_ov296 VM_X_OVERLAP=1 VM_X_PERSIST=296
run persist592_serial VM_X_PERSIST=592
