T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $T --master-port 29540 tools/dp_check.py > gpurun_out/dp_check_default_2gpu.log 2>&1; echo "dp_check default rc=$?"
DFGAN_DP_SHARDED=1 timeout 300 $T --master-port 29541 tools/dp_check.py > gpurun_out/dp_check_sharded_2gpu.log 2>&1; echo "dp_check sharded rc=$?"
grep -v "^\s*$\|OMP_NUM\|^\*\*\*" gpurun_out/dp_check_sharded_2gpu.log | tail -9
DFGAN_DP_SHARDED=1 timeout 300 $T --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_sharded.json 2> gpurun_out/bench_2gpu_sharded.err; echo "bench sharded rc=$?"
timeout 300 $T --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_unsharded.json 2> gpurun_out/bench_2gpu_unsharded.err; echo "bench unsharded rc=$?"
DFGAN_DP_SHARDED=1 timeout 300 $T --master-port 29544 tools/prof_timeline.py > gpurun_out/timeline_2gpu_sharded.csv 2> gpurun_out/timeline_2gpu_sharded.err; echo "timeline rc=$?"
python - <<'P'
import json
for f in ['bench_2gpu_sharded','bench_2gpu_unsharded']:
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f,d['value'],d['ms_per_step'],d['e2e']['value'],d['config']['losses'])
    except Exception as e: print(f,'ERR',e)
P
grep -n "dfgan peer" gpurun_out/*2gpu*shard*.json gpurun_out/*2gpu*shard*.err | head -3
