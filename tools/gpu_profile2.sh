#!/bin/bash
# ncu evidence for profiles/: launch list of one full step + full captures of the fused attention kernels
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 11000 --csv --log-file gpurun_out/r01b_launches_step.csv \
    python tools/one_step.py > gpurun_out/ncu_step.log 2>&1
python tools/ncu_launches.py gpurun_out/r01b_launches_step.csv > gpurun_out/r01b_launches_step_summary.txt 2>&1
gzip -f gpurun_out/r01b_launches_step.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cross_attn_mq --launch-skip 100 --launch-count 1 \
    -o gpurun_out/r01b_prof_xattn_q16 python tools/one_step.py > gpurun_out/ncu_xq16.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cross_attn_mq --launch-skip 600 --launch-count 1 \
    -o gpurun_out/r01b_prof_xattn_mq python tools/one_step.py > gpurun_out/ncu_xmq.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:window_attention_split --launch-skip 6 --launch-count 1 \
    -o gpurun_out/r01b_prof_wattn python tools/one_step.py > gpurun_out/ncu_wattn.log 2>&1
tail -2 gpurun_out/ncu_step.log gpurun_out/ncu_xq16.log gpurun_out/ncu_xmq.log gpurun_out/ncu_wattn.log; head -32 gpurun_out/r01b_launches_step_summary.txt
ls -la gpurun_out/*.ncu-rep
