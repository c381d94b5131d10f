#!/bin/bash
# evidence with the final defaults: ncu launch lists (omni, mgpstr), full capture of the TMA-store GEMM epilogue kernel
mkdir -p gpurun_out
for wl in omni mgpstr; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02f_launches_$wl.csv python tools/one_step.py $wl > gpurun_out/ncu_$wl.log 2>&1; echo "ncu $wl rc=$?"
  python tools/ncu_launches.py gpurun_out/r02f_launches_$wl.csv > gpurun_out/r02f_launches_${wl}_summary.txt 2>&1; gzip -f gpurun_out/r02f_launches_$wl.csv; head -16 gpurun_out/r02f_launches_${wl}_summary.txt
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 6 --launch-count 1 \
    -o gpurun_out/r02f_prof_gemm_vit_fc1_tma python tools/one_step.py mgpstr > gpurun_out/ncu_gemm_tma.log 2>&1; tail -2 gpurun_out/ncu_gemm_tma.log
ls -la gpurun_out/r02f_* | head
