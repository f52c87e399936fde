set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_2gpu.log; tail -4 gpurun_out/r02_gputest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"
python tools/resident_probe.py 28 30 > gpurun_out/r02_resident_2gpu.json 2>&1; echo "probe rc=$?"
PNUMPY_CUDA_DEVICES=1 python tools/resident_probe.py 28 > gpurun_out/r02_resident_1gpu.json 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/exchange_stress.py 200 > gpurun_out/r02_exchange_stress_n2.json 2>&1; echo "stress rc=$?"
timeout 900 compute-sanitizer --tool memcheck --target-processes all --log-file gpurun_out/r02_memcheck.log python -m pytest tests -m gpu -x -q -k "exchange or one_launch or fused or shard_over or two_pass or nan_reductions" > gpurun_out/r02_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --target-processes all --log-file gpurun_out/r02_racecheck.log python -m pytest tests -m gpu -x -q -k "exchange or one_launch or fused_exchange or two_pass or nan_reductions" > gpurun_out/r02_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/r02_memcheck_pytest.log gpurun_out/r02_racecheck_pytest.log; tail -5 gpurun_out/r02_memcheck.log gpurun_out/r02_racecheck.log
