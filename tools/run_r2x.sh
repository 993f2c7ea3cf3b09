F="--no-graphs --no-cpu-baseline --no-e2e --no-parity --no-c5"
ncu --set full --clock-control none --import-source on -k regex:"^k_special|^k_gamma_partial" -s 4 -c 3 -o gpurun_out/r2x_c4_full -f python bench.py --config c4 --steps 4 --warmup 2 $F > gpurun_out/r2x_ncu_c4.log 2>&1; echo rc=$?
