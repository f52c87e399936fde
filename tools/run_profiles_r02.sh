# one gpurun call (1 GPU): every ncu / probe artefact that goes into profiles/r02_*  (see B200_PROFILING.md: each command
# runs plain first, then under ncu, with `&&` in between)
set -x
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-detail"
$B > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ew2_vec_kernel<pncu::AddF<double>" -s 3 -c 2 -o gpurun_out/r02_add_f64 $B > gpurun_out/r02_ncu_full.log 2>&1
echo "set full rc=$?"
S="python tools/ncu_sweep.py run 28 --cases gpurun_out/r02_small_cases.json --dtypes bool,int8,uint8,int16,uint16"
$S > gpurun_out/r02_small_run.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_small_sweep.csv $S > gpurun_out/r02_ncu_small.log 2>&1
echo "small sweep rc=$?"
python tools/ncu_sweep.py time 28 --cases gpurun_out/r02_small_events.json --dtypes bool,int8,uint8,int16,uint16 > gpurun_out/r02_small_time.log 2>&1
A="python tools/time_unary.py ASINH float32 -80 80"
$A > gpurun_out/r02_asinh_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"ew1_vec_kernel<.*AsinhF32" -s 3 -c 1 -o gpurun_out/r02_asinh_f32 $A > gpurun_out/r02_ncu_asinh.log 2>&1
echo "asinh rc=$?"
python tools/misaligned_probe.py 28 > gpurun_out/r02_misaligned.md 2>&1
python tools/pn_benchmark.py > gpurun_out/r02_pn_benchmark.txt 2>&1
python tools/trig_table.py > gpurun_out/r02_trig_table.md 2>&1
python tools/reduce_breakdown.py ADD > gpurun_out/r02_reduce_breakdown.txt 2>&1
ls -la gpurun_out | head -40
