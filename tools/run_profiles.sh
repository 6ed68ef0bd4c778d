#!/bin/bash
# One-GPU evidence run (under gpurun): ncu launch list of an eager step, ncu --set full of the step's GEMM launches and of
# the HBM-bound kernels, compute-sanitizer memcheck / racecheck / synccheck of a small full step and of GEMM self-test cases.
# Outputs under gpurun_out/r02p/ (summaries are copied to profiles/ by hand).
O=gpurun_out/r02p
mkdir -p $O
NCU=$(which ncu || echo /usr/local/cuda/bin/ncu)
# ---- launch list (cold-cache, serialised: compare SHARES)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 104 -c 110 --csv --log-file $O/launches.csv \
    python tools/profile_step.py 4 > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv $O/launches_step.md > /dev/null 2>&1 || echo "summarize_launches failed"
# ---- full metric set on the 24 GEMM launches of the third eager step
timeout 900 $NCU --set full --clock-control none -k regex:gemm_kernel -s 48 -c 24 -o $O/gemm_full -f \
    python tools/profile_step.py 3 > $O/gemm_full.log 2>&1
$NCU -i $O/gemm_full.ncu-rep --page raw --csv > $O/gemm_full_raw.csv 2>/dev/null
python tools/summarize_ncu_full.py $O/gemm_full_raw.csv $O/gemm_ncu_full.md --traffic-json=$O/gemm_ncu_traffic.json > /dev/null 2>&1 || echo "summarize gemm failed"
# ---- source-level capture of ONE short-K forward launch (tdnn4 fwd = 4th bf16-epilogue pair launch of a step)
timeout 600 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_kernel<\(int\)0, \(int\)2>' -s 21 -c 1 -o $O/gemm_tdnn4_src -f \
    python tools/profile_step.py 3 > $O/gemm_src.log 2>&1
$NCU -i $O/gemm_tdnn4_src.ncu-rep --page source --csv > $O/gemm_tdnn4_source.csv 2>/dev/null
# ---- HBM-bound kernels of one step
timeout 900 $NCU --set full --clock-control none -k regex:'bn_act|stats_pool|opt_step|pool_bn|pack_input|head_prep|head_finish' -s 36 -c 18 \
    -o $O/hbm_full -f python tools/profile_step.py 3 > $O/hbm_full.log 2>&1
$NCU -i $O/hbm_full.ncu-rep --page raw --csv > $O/hbm_full_raw.csv 2>/dev/null
python tools/summarize_ncu_full.py $O/hbm_full_raw.csv $O/hbm_kernels_ncu.md > /dev/null 2>&1 || echo "summarize hbm failed"
rm -f $O/gemm_full.ncu-rep $O/hbm_full.ncu-rep          # keep the merge-back under the size limit; the raw CSVs stay
# ---- compute-sanitizer on a small full training step (forward, backward, optimizer) and on GEMM self-test cases
CS=$(which compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python tools/parity_debug.py 16 60 200 > $O/sanitizer_${tool}_step.log 2>&1
  echo "$tool step: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/sanitizer_${tool}_step.log | tail -1)"
done
for c in conv_fwd_stats_pairs dgrad_bnbwd_relu wgrad mn_mn_split edges_bf16_mn; do
  timeout 600 $CS --tool memcheck --print-limit 20 python tools/gemm_selftest.py --case $c > $O/sanitizer_memcheck_gemm_$c.log 2>&1
  echo "memcheck $c: $(grep -E 'ERROR SUMMARY' $O/sanitizer_memcheck_gemm_$c.log | tail -1)"
done
timeout 600 $CS --tool racecheck --print-limit 20 python tools/gemm_selftest.py --case conv_fwd_stats_pairs > $O/sanitizer_racecheck_gemm_stats.log 2>&1
echo "racecheck gemm stats: $(grep -E 'RACECHECK SUMMARY' $O/sanitizer_racecheck_gemm_stats.log | tail -1)"
ls -la $O | head -40
