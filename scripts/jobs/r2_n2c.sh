# N GPUs: CTA cap of the head-bucket communicator, A/B
set -x
T=r2_n2c
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29620
for cap in 0 4 16 0 8; do
  port=$((port+1))
  SCDA_NCCL_OPT_CTAS=$cap timeout 240 $TR --master-port $port bench.py --gpus $N --steps 40 --warmup 3 --no-cpu-baseline --no-parity-line 2> gpurun_out/${T}_${N}_cap$cap.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('cap $cap', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
