# one 8-GPU box: peer-memory exchange at N=8 (kernel check, trainer check, iteration bench), then the MA-GP step's
# strong-scaling points (BASELINE configs[3]: batch 64 in total over 8 / 4 / 2 / 1 GPUs)
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $T --nproc-per-node 8 --master-port 29511 tools/peer_check.py > gpurun_out/peer_check_8gpu.log 2>&1; echo "peer_check rc=$?"
timeout 300 $T --nproc-per-node 8 --master-port 29512 tools/dp_check.py > gpurun_out/dp_check_8gpu.log 2>&1; echo "dp_check rc=$?"
timeout 300 $T --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_peer.json 2> gpurun_out/bench_8gpu_peer.err; echo "bench 8 rc=$?"
for n in 8 4 2; do
  timeout 200 $T --nproc-per-node $n --master-port 2952$n bench.py --workload gp --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_gp_b64_${n}gpu.json 2> gpurun_out/bench_gp_b64_${n}gpu.err; echo "gp $n rc=$?"
done
timeout 200 python bench.py --workload gp --steps 20 --warmup 5 > gpurun_out/bench_gp_b64_1gpu.json 2> gpurun_out/bench_gp_b64_1gpu.err; echo "gp 1 rc=$?"
grep "MB\|range\|replays\|bit for bit" gpurun_out/peer_check_8gpu.log; tail -4 gpurun_out/dp_check_8gpu.log
python - <<'P'
import json
for f in ['bench_8gpu_peer','bench_gp_b64_8gpu','bench_gp_b64_4gpu','bench_gp_b64_2gpu','bench_gp_b64_1gpu']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f,d['value'],d['ms_per_step'])
    except Exception as e: print(f,'ERR',e)
P
