#!/bin/bash
# Round-2 follow-up evidence run (under gpurun, one GPU): launch list and ncu --set full of the HBM-bound kernels with the
# ReLU-specialised pooling kernels in place; compute-sanitizer memcheck / racecheck on the GhostVLAD and metric-loss kernels.
O=gpurun_out/r02q
mkdir -p $O
NCU=$(which ncu || echo /usr/local/cuda/bin/ncu)
CS=$(which compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 104 -c 110 --csv --log-file $O/launches.csv \
    python tools/profile_step.py 4 > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv $O/launches_step.md > /dev/null 2>&1 || echo "summarize_launches failed"
timeout 900 $NCU --set full --clock-control none -k regex:'bn_act|stats_pool|opt_step|pool_bn|pack_input|head_prep|head_finish' -s 36 -c 18 \
    -o $O/hbm_full -f python tools/profile_step.py 3 > $O/hbm_full.log 2>&1
$NCU -i $O/hbm_full.ncu-rep --page raw --csv > $O/hbm_full_raw.csv 2>/dev/null
python tools/summarize_ncu_full.py $O/hbm_full_raw.csv $O/hbm_kernels_ncu.md > /dev/null 2>&1 || echo "summarize hbm failed"
rm -f $O/hbm_full.ncu-rep
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python -m pytest -q -x tests/test_vlad_gpu.py -k "forward_backward or golden" \
      > $O/sanitizer_${tool}_vlad.log 2>&1
  echo "$tool vlad: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitizer_${tool}_vlad.log | tail -1) | $(grep -E 'passed|failed' $O/sanitizer_${tool}_vlad.log | tail -1)"
  timeout 900 $CS --tool $tool --print-limit 20 python -m pytest -q -x tests/test_metric_gpu.py -k "forward_backward or known_answers" \
      > $O/sanitizer_${tool}_metric.log 2>&1
  echo "$tool metric: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitizer_${tool}_metric.log | tail -1) | $(grep -E 'passed|failed' $O/sanitizer_${tool}_metric.log | tail -1)"
done
ls -la $O | head -20
