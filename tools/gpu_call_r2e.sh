#!/bin/bash
# Round 2, fifth call (1 GPU): A/B inside ONE box -- L2 access policy of the packed kernel's scratch on/off (time and DRAM
# bytes), wavefront column block 32,2 vs 24,3 on the full configs[3]; Kimura tests.
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_identity.py -m gpu -q --timeout 200 ; echo "exit $?" ) > gpurun_out/r2e_tests.log 2>&1
tail -n 4 gpurun_out/r2e_tests.log
for w in c2 c3 c5s; do
  for pol in "" "1"; do
    echo "== $w TSQ_NO_L2_POLICY=$pol"
    ( TSQ_NO_L2_POLICY=$pol timeout 200 python tools/prof_run.py $w 4 2>&1 | tail -n 2 )
  done
done 2>&1 | tee gpurun_out/r2e_l2_ab.log
for pol in "" "1"; do
  ( TSQ_NO_L2_POLICY=$pol timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gotoh16 -c 2 --csv --log-file gpurun_out/r2e_dram_c2_nopolicy${pol}.csv python tools/prof_run.py c2 2 > /dev/null 2>&1 )
  grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/r2e_dram_c2_nopolicy${pol}.csv | cut -d, -f13-15 | tail -n 4
done
for v in "32,2" "24,3" "32,2" "24,3"; do
  echo "== c4 full TSQ_FORCE_KW16=$v"
  ( TSQ_FORCE_KW16=$v timeout 300 python tools/prof_run.py c4 2 2>&1 | tail -n 1 )
done 2>&1 | tee gpurun_out/r2e_kw_ab.log
