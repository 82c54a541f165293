# one 8-GPU box: the in-switch (multimem) exchange at N = 8 -- kernel check + bandwidth sweep (also times the unicast
# kernel on the same buffers), trainer check, iteration bench with the in-switch and with the unicast exchange
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 8 --master-port 29511 tools/peer_check.py > gpurun_out/peer_check_8gpu_mc.log 2>&1; echo "peer_check rc=$?"
timeout 300 $T --nproc-per-node 8 --master-port 29512 tools/dp_check.py > gpurun_out/dp_check_8gpu_mc.log 2>&1; echo "dp_check rc=$?"
timeout 300 $T --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_mc.json 2> gpurun_out/bench_8gpu_mc.err; echo "bench 8 multimem rc=$?"
DFGAN_PEER_MULTIMEM=0 timeout 300 $T --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_uc.json 2> gpurun_out/bench_8gpu_uc.err; echo "bench 8 unicast rc=$?"
timeout 300 $T --nproc-per-node 4 --master-port 29515 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu_mc.json 2> gpurun_out/bench_4gpu_mc.err; echo "bench 4 multimem rc=$?"
grep "MB\|range\|replays\|exchange kernel\|unicast\|==\|dfgan\]" gpurun_out/peer_check_8gpu_mc.log; tail -3 gpurun_out/dp_check_8gpu_mc.log
python - <<'P'
import json
for f in ['bench_8gpu_mc','bench_8gpu_uc','bench_4gpu_mc']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f,d['value'],d['ms_per_step'],d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
P
