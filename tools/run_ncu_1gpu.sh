# ncu call of the round (one profiler per gpurun call): launch list with DRAM bytes of ~2 eager iterations, then one
# --set full capture of the two largest GEMM kernels.  Each ncu run follows a plain run of the same command (&&).
set -u
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_q2.json 2> gpurun_out/bench_q2.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_q2.json
python tools/prof_timeline.py > gpurun_out/timeline4.csv 2> gpurun_out/timeline4.err; python tools/timeline_tail.py gpurun_out/timeline4.csv 0 "linear|adam" | tail -12
export DFGAN_NO_GRAPH=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1800 -c 950 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_halo2_kernel|wgrad_umma_kernel" -s 60 -c 4 -o gpurun_out/prof $CMD > gpurun_out/ncu2.log 2>&1; echo "set full rc=$?"
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>/dev/null; ls -la gpurun_out/prof* gpurun_out/launches.csv; tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
