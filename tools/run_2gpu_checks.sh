# one gpurun --gpus 2 call: the whole GPU suite on two GPUs, the bench line at N=2, the in-process dispatcher probe,
# the exchange stress
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest_2gpu.log; tail -n 3 gpurun_out/r02_gputest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench2 rc=$?"
python tools/resident_probe.py 28 30 > gpurun_out/r02_resident_2gpu.txt 2>&1; echo "probe rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/exchange_stress.py 200 > gpurun_out/r02_exchange_stress_n2.json 2>&1; echo "stress rc=$?"; tail -n 1 gpurun_out/r02_exchange_stress_n2.json
head -14 gpurun_out/r02_resident_2gpu.txt
