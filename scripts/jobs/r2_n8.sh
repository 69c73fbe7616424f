# 8-GPU job: throughput with the NCCL all-reduces captured in the one graph vs the cut plan
set -x
T=r2_n8
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29611 bench.py --gpus $N --steps 40 --warmup 3 --graph-collectives --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench_graph_$N.json 2> gpurun_out/${T}_bench_graph_$N.err; echo rc=$?; cut -c1-330 gpurun_out/${T}_bench_graph_$N.json; tail -3 gpurun_out/${T}_bench_graph_$N.err
timeout 240 $TR --master-port 29612 bench.py --gpus $N --steps 40 --warmup 3 --no-graph-collectives --no-cpu-baseline --no-parity-line > gpurun_out/${T}_bench_cut_$N.json 2> gpurun_out/${T}_bench_cut_$N.err; echo rc=$?; cut -c1-330 gpurun_out/${T}_bench_cut_$N.json; tail -3 gpurun_out/${T}_bench_cut_$N.err
