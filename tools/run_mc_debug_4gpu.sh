T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
CUDA_LAUNCH_BLOCKING=1 timeout 200 $T --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/mcdbg_blocking.json 2> gpurun_out/mcdbg_blocking.err; echo "blocking rc=$?"
grep -n "Error\|line [0-9]*, in\|dfgan peer" gpurun_out/mcdbg_blocking.err | grep -v "site-packages" | head -30
DFGAN_DP_NO_OVERLAP=1 timeout 200 $T --master-port 29522 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/mcdbg_noovl.json 2> gpurun_out/mcdbg_noovl.err; echo "no-overlap rc=$?"
grep -n "Error\|dfgan peer" gpurun_out/mcdbg_noovl.err | grep -v "site-packages" | head -5
