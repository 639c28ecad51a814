#!/bin/bash
# Round-2 (third session) evidence capture on the GPU box, final build. Every ncu command runs only after the same command exited 0 without ncu.
set -x
mkdir -p gpurun_out
# 1. headline command: launch list + DRAM traffic (cooperative launches cannot be replayed by ncu: LQT_FK_NOCOOP=1, same grid)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check --no-c3 --c4-utterances 0"
LQT_FK_NOCOOP=1 timeout 300 $CMD > gpurun_out/r2s3_cap_plain.json 2> gpurun_out/r2s3_cap_plain.err || exit 1
LQT_FK_NOCOOP=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2s3_final_launches.csv $CMD > gpurun_out/r2s3_cap_ncu.log 2>&1
tail -2 gpurun_out/r2s3_cap_ncu.log
# 2. frame kernel, ncu --set full (one launch: 9 prefill rows + 8 frames)
python tools/fk_ncu_run.py --frames 8 || exit 1
LQT_FK_NOCOOP=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:frame_kernel -c 1 -f -o gpurun_out/r2s3_fk python tools/fk_ncu_run.py --frames 8 > gpurun_out/r2s3_fk_ncu.log 2>&1
tail -2 gpurun_out/r2s3_fk_ncu.log
# 3. phase timeline of CTA 0 (profiling build with marks)
python tools/fk_timeline.py --frames 12 > gpurun_out/r2s3_timeline_cta0.txt 2>&1
tail -1 gpurun_out/r2s3_timeline_cta0.txt
