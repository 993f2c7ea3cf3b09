B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-parity"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2p_$name.json 2> gpurun_out/r2p_$name.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2p_$name.json').read().strip().splitlines()[-1])
    c5=(d.get('configs') or {}).get('c5') or {}
    print('$name', 'ms_per_step=%.4f dense_ms=%.4f elbo=%r c5_ms=%s c5_dense=%s' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['elbo_final'], c5.get('ms_per_step'), (c5.get('roofline') or {}).get('kernel_ms')))
except Exception as e:
    print('$name', 'ERR', e); print(open('gpurun_out/r2p_$name.err').read()[-1500:])
PY
}
run tma X=0
run notma VM_X_NOTMA=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2p_pytest.log
