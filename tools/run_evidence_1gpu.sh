# round-end evidence on one B200 (no profiler in this call): GPU test suite, the bench line with both baselines,
# per-launch table, kernel timeline of graph replays
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python tools/prof_perlaunch.py > gpurun_out/perlaunch.txt 2> gpurun_out/perlaunch.err; echo "perlaunch rc=$?"
python tools/prof_timeline.py > gpurun_out/timeline.csv 2> gpurun_out/timeline.err; echo "timeline rc=$?"
python tools/timeline_report.py gpurun_out/timeline.csv > gpurun_out/timeline_report.txt
python tools/timeline_tail.py gpurun_out/timeline.csv 40 > gpurun_out/timeline_tail.txt
python bench.py --workload sample --steps 20 --warmup 5 > gpurun_out/bench_sample_b256.json 2> gpurun_out/bench_sample.err; echo "sample rc=$?"
cut -c1-400 gpurun_out/bench.json
