# 4-GPU box, final build, default (unicast peer-memory) exchange: the 2/4-GPU pytest cases, then the iteration bench at 4 and 2 GPUs
python -m pytest tests/test_gpu_dp.py -m gpu -x -q > gpurun_out/pytest_gpu_dp.log 2>&1; echo "pytest dp rc=$?"; tail -3 gpurun_out/pytest_gpu_dp.log
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 4 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu_end.json 2> gpurun_out/bench_4gpu_end.err; echo "bench 4 rc=$?"
timeout 300 $T --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_end.json 2> gpurun_out/bench_2gpu_end.err; echo "bench 2 rc=$?"
python - <<'P'
import json
for f in ['bench_4gpu_end','bench_2gpu_end']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f,d['value'],d['ms_per_step'],d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
P
