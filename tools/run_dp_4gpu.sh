# 4-GPU box, final build: iteration bench with the default (unicast peer-memory) exchange, with the in-switch exchange
# (DFGAN_PEER_MULTIMEM=1) and with the sharded optimiser step (DFGAN_DP_SHARDED=1); trainer check with the in-switch exchange
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu_end.json 2> gpurun_out/bench_4gpu_end.err; echo "bench 4 default rc=$?"
DFGAN_PEER_MULTIMEM=1 timeout 300 $T --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu_mc.json 2> gpurun_out/bench_4gpu_mc.err; echo "bench 4 multimem rc=$?"
DFGAN_DP_SHARDED=1 timeout 300 $T --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu_sharded.json 2> gpurun_out/bench_4gpu_sharded.err; echo "bench 4 sharded rc=$?"
DFGAN_PEER_MULTIMEM=1 timeout 200 $T --master-port 29534 tools/dp_check.py > gpurun_out/dp_check_4gpu_mc.log 2>&1; echo "dp_check multimem rc=$?"; tail -3 gpurun_out/dp_check_4gpu_mc.log
python - <<'P'
import json
for f in ['bench_4gpu_end','bench_4gpu_mc','bench_4gpu_sharded']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f,d['value'],d['ms_per_step'],d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
P
grep -n "dfgan peer\|dfgan\]" gpurun_out/bench_4gpu_*.json gpurun_out/bench_4gpu_*.err | head -5
