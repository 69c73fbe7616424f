# 2-GPU job: parameter identity + throughput, NCCL captured in the one graph (a communicator per stream) vs the cut plan
set -x
T=r2_n2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29601 scripts/ddp_check.py --graph-collectives > gpurun_out/${T}_ddp_graph.log 2>&1; echo rc=$?; tail -6 gpurun_out/${T}_ddp_graph.log
timeout 240 $TR --master-port 29602 scripts/ddp_check.py --no-graph-collectives > gpurun_out/${T}_ddp_cut.log 2>&1; echo rc=$?; tail -6 gpurun_out/${T}_ddp_cut.log
timeout 300 $TR --master-port 29603 bench.py --gpus 2 --steps 50 --warmup 3 --graph-collectives > gpurun_out/${T}_bench_graph.json 2> gpurun_out/${T}_bench_graph.err; echo rc=$?; cut -c1-330 gpurun_out/${T}_bench_graph.json; tail -3 gpurun_out/${T}_bench_graph.err
timeout 300 $TR --master-port 29604 bench.py --gpus 2 --steps 50 --warmup 3 --no-graph-collectives > gpurun_out/${T}_bench_cut.json 2> gpurun_out/${T}_bench_cut.err; echo rc=$?; cut -c1-330 gpurun_out/${T}_bench_cut.json; tail -3 gpurun_out/${T}_bench_cut.err
SCDA_ONE_COMM=1 timeout 300 $TR --master-port 29605 bench.py --gpus 2 --steps 50 --warmup 3 --no-graph-collectives > gpurun_out/${T}_bench_cut1comm.json 2> gpurun_out/${T}_bench_cut1comm.err; echo rc=$?; cut -c1-330 gpurun_out/${T}_bench_cut1comm.json
