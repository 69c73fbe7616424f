# third detector bucket at N GPUs: parameter identity + A/B
set -x
T=r2_ae
N=${1:-2}
timeout 600 python -m pytest tests/test_iteration_parity_gpu.py tests/test_ddp_gpu.py -x -q -k "bf16x3 or graph-collectives or identical" > gpurun_out/${T}_test.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_test.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29640
for v in 0 1 0 1; do
  port=$((port+1))
  SCDA_MID_BUCKET=$v timeout 240 $TR --master-port $port bench.py --gpus $N --steps 40 --warmup 3 --no-cpu-baseline --no-parity-line 2> gpurun_out/${T}_${N}_mid$v.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('mid $v', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
