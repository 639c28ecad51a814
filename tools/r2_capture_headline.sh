set -x
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check --no-c3 --c4-utterances 0"
LQT_FK_NOCOOP=1 timeout 300 $CMD > gpurun_out/r2_cap_plain.json 2> gpurun_out/r2_cap_plain.err || exit 1
LQT_FK_NOCOOP=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_final_launches.csv $CMD > gpurun_out/r2_cap_ncu.log 2>&1
tail -3 gpurun_out/r2_cap_ncu.log
wc -l gpurun_out/r2_final_launches.csv
