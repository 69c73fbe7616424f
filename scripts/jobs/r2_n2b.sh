# 2-GPU job: parameter identity in the three execution plans, throughput
set -x
T=r2_n2b
timeout 900 python -m pytest tests/test_ddp_gpu.py -x -q > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 50 --warmup 3 --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; cut -c1-330 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
