#!/bin/bash
# round-2 captures of the roofline launches with the TMA-store epilogue: Swin stage-2 fc1 (three passes) and ViT fc1 (bf16)
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 1 -o gpurun_out/r02f_prof_gemm_fc1 \
    python tools/bench_one_gemm.py 65536 2048 512 1 1 1 > gpurun_out/ncu_fc1b.log 2>&1; tail -2 gpurun_out/ncu_fc1b.log
ALM_OCR_OPTIONS="nsplit=1" timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 4 -c 1 -o gpurun_out/r02f_prof_gemm_vit_fc1 \
    python tools/bench_one_gemm.py 131584 3072 768 1 1 1 > gpurun_out/ncu_fc1v.log 2>&1; tail -2 gpurun_out/ncu_fc1v.log
python tools/bench_one_gemm.py 65536 2048 512 1 1 1
ALM_OCR_OPTIONS="nsplit=1" python tools/bench_one_gemm.py 131584 3072 768 1 1 1
ALM_OCR_OPTIONS="nsplit=1,gemm_plain_epilogue=1" python tools/bench_one_gemm.py 131584 3072 768 1 1 1
ALM_OCR_OPTIONS="gemm_plain_epilogue=1" python tools/bench_one_gemm.py 65536 2048 512 1 1 1
ls -la gpurun_out/r02f_prof_gemm*.ncu-rep
