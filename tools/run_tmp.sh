timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -12
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline --no-c5 --no-parity > gpurun_out/r3o_bench2.json 2> gpurun_out/r3o_bench2.err; echo rc=$?
python tools/show_bench.py gpurun_out/r3o_bench2.json; python -c "
import json; d=json.loads(open('gpurun_out/r3o_bench2.json').read().strip().splitlines()[-1]); print(d['e2e']['timings_s'], d['e2e']['h2d_bytes_per_step'])"
