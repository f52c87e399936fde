# one gpurun --gpus 8 call: the in-process dispatcher over 8 GPUs, the bench line at N=8, the exchange stress
set -x
python tools/resident_probe.py 28 30 --trace > gpurun_out/r02_resident_8gpu.txt 2>&1; echo "probe rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench8 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tools/exchange_stress.py 200 > gpurun_out/r02_exchange_stress_n8.json 2>&1; echo "stress rc=$?"
timeout 600 python -m pytest tests -m gpu -x -q -k "shard or lanes or exchange or managed or resident or one_launch" > gpurun_out/r02_gputest_8gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02_gputest_8gpu.log
PNUMPY_CUDA_DEVICES=4 python tools/resident_probe.py 28 30 > gpurun_out/r02_resident_4gpu.txt 2>&1; echo "probe4 rc=$?"
head -22 gpurun_out/r02_resident_8gpu.txt
