# A/B on one box: new library vs the previous one (scda_b200/csrc/libold_b200.so), bench interleaved
set -x
T=r2_ac
cd scda_b200/csrc; cp libscda_b200.so libnew_b200.so; cd ../..
for rep in 1 2; do
  for v in new old; do
    cp scda_b200/csrc/lib${v}_b200.so scda_b200/csrc/libscda_b200.so
    timeout 300 python bench.py --steps 60 --no-cpu-baseline --no-parity-line 2> gpurun_out/${T}_${v}${rep}.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$v', d['ms_per_step'], d['value'])"
  done
done
cp scda_b200/csrc/libnew_b200.so scda_b200/csrc/libscda_b200.so
