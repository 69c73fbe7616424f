# the driver's own 8-GPU command at the final commit
set -x
T=r2_n8b
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 50 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; cut -c1-330 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
