#!/bin/bash
# instruction-cache hit rate of the frame kernel for each library given (ncu, one pass, 8 frames)
for lib in "$@"; do
  echo "== $lib"
  LQT_B200_LIB=$(realpath $lib) LQT_FK_NOCOOP=1 timeout 300 ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gpu__time_duration.sum --clock-control none -k regex:frame_kernel -c 1 python tools/fk_ncu_run.py --frames 8 2>&1 | grep -E "icc|duration"
done
