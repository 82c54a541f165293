T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29511 tools/peer_check.py > gpurun_out/peer_check_2gpu.log 2>&1; echo "peer_check rc=$?"
timeout 300 $T --master-port 29512 tools/dp_check.py > gpurun_out/dp_check_peer_2gpu.log 2>&1; echo "dp_check rc=$?"
timeout 300 $T --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_peer.json 2> gpurun_out/bench_2gpu_peer.err; echo "bench peer rc=$?"
grep -v "^\s*$" gpurun_out/peer_check_2gpu.log | grep -v "^W10\|^E10\|File \|^\s\s\|OMP_NUM\|^\*\*\*" | tail -22
tail -4 gpurun_out/dp_check_peer_2gpu.log
python - <<'P'
import json
for f in ['bench_2gpu_peer']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f,d['value'],d['ms_per_step'],d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
P
