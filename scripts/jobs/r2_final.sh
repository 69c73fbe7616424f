# round-2 final evidence job (1 GPU): bench line, configs 2/3/5, operator / convolution benches, launch list of one
# replayed iteration, ncu --set full of the convolution (fwd / dgrad / wgrad) and RoI / NMS kernels summarised ON THE BOX
# (the .ncu-rep files are too large to bring back), GPU suite
set -x
T=r2_final
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$?; tail -3 gpurun_out/${T}_gputest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo rc=$?; tail -c 300 gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench.json
for c in 2 3 5; do timeout 300 python bench.py --config $c > gpurun_out/${T}_config$c.json 2> gpurun_out/${T}_config$c.err; echo rc=$?; done
timeout 300 python scripts/opbench.py > gpurun_out/${T}_opbench.jsonl 2> gpurun_out/${T}_opbench.err; echo rc=$?
timeout 300 python scripts/convbench.py > gpurun_out/${T}_convbench.jsonl 2> gpurun_out/${T}_convbench.err; echo rc=$?
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python scripts/step_profile.py > gpurun_out/${T}_launches.log 2>&1; tail -2 gpurun_out/${T}_launches.log
L="conv1_1p conv1_2 conv2_1 conv2_2 conv3_1 conv3_2 conv4_1 conv4_2 conv5_x"
for p in fwd dgrad; do
  timeout 400 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${T}_conv_$p python scripts/prof_conv.py --pass=$p > gpurun_out/${T}_conv_$p.log 2>&1; tail -1 gpurun_out/${T}_conv_$p.log
  python scripts/ncu_summary.py /tmp/${T}_conv_$p.ncu-rep $L > gpurun_out/${T}_ncu_conv_$p.txt 2>&1
done
timeout 400 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${T}_conv_wgrad python scripts/prof_conv.py --pass=wgrad > gpurun_out/${T}_conv_wgrad.log 2>&1; tail -1 gpurun_out/${T}_conv_wgrad.log
python scripts/ncu_summary.py /tmp/${T}_conv_wgrad.ncu-rep > gpurun_out/${T}_ncu_conv_wgrad.txt 2>&1
timeout 400 ncu --set full --clock-control none -k regex:'roi_|nms_' -c 14 -f -o /tmp/${T}_ops python scripts/prof_ops.py all > gpurun_out/${T}_ops.log 2>&1; tail -1 gpurun_out/${T}_ops.log
python scripts/ncu_summary.py /tmp/${T}_ops.ncu-rep > gpurun_out/${T}_ncu_ops.txt 2>&1
SCDA_TIMESTAMPS=1 timeout 300 python scripts/phase_times.py > gpurun_out/${T}_phases.txt 2>&1; tail -16 gpurun_out/${T}_phases.txt
du -sh gpurun_out
