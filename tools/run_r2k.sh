set -x
./tools/store_patterns > gpurun_out/r2k_store_patterns.txt 2>&1
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-c5"
for m in 0 1 2; do VM_X_OVERLAP=$m $B > gpurun_out/r2k_overlap$m.json 2> gpurun_out/r2k_overlap$m.err; done
VM_X_OVERLAP=1 $B --no-graphs > gpurun_out/r2k_overlap1_nographs.json 2>&1
VM_PACK_TRACE=1 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-parity --no-c5 > gpurun_out/r2k_packtrace.log 2>&1
