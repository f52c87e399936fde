# final 1-GPU pass of the round: tests, smoke, both bench arms, the probes whose numbers changed, the stress
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_1gpu.log; tail -n 3 gpurun_out/r02_gputest_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$?"; tail -c 600 gpurun_out/r02_bench_reference.json
python tools/misaligned_probe.py 28 > gpurun_out/r02_misaligned.md 2>&1
S="python tools/ncu_sweep.py run 28 --cases gpurun_out/r02_small_cases.json --dtypes bool,int8,uint8,int16,uint16"
$S > gpurun_out/r02_small_run.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small_sweep.csv $S > gpurun_out/r02_ncu_small.log 2>&1
echo "small sweep rc=$?"
python tools/ncu_sweep.py time 28 --cases gpurun_out/r02_small_events.json --dtypes bool,int8,uint8,int16,uint16 > gpurun_out/r02_small_time.log 2>&1
python tools/reduce_breakdown.py ADD > gpurun_out/r02_reduce_breakdown.txt 2>&1
timeout 200 python tools/stress.py > gpurun_out/r02_stress.txt 2>&1; echo "stress rc=$?"; tail -n 2 gpurun_out/r02_stress.txt
