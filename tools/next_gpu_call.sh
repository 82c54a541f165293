#!/usr/bin/env bash
# First GPU call of the next round: everything that was written after round 1's GPU budget was spent, in one gpurun.
#   gpurun --timeout 900 -- 'bash tools/next_gpu_call.sh'
# Outputs land in gpurun_out/ (copy what is to be judged into profiles/).
set -u
mkdir -p gpurun_out
python tools/bringup_configs.py                    > gpurun_out/configs.log    2>&1; echo "configs rc=$?"
python tools/prof_shell.py --overlap 0             > gpurun_out/prof_shell0.json 2> gpurun_out/prof_shell0.err; echo "prof_shell overlap=0 rc=$?"
python tools/prof_shell.py --overlap 1             > gpurun_out/prof_shell1.json 2> gpurun_out/prof_shell1.err; echo "prof_shell overlap=1 rc=$?"
python -m pytest tests -m gpu -x -q                > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"
python bench.py --steps 20 --warmup 5              > gpurun_out/bench.json     2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/configs.log gpurun_out/pytest_gpu.log
cat gpurun_out/prof_shell0.json gpurun_out/prof_shell1.json
cut -c1-300 gpurun_out/bench.json
