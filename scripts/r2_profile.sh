#!/bin/bash
# round-2 ncu evidence: launch list of the bench command + one full capture per section of the hot path.  One GPU.
# gpurun brings back at most 64 MiB: the reports are summarised on the box and only the two small ones travel.
O=gpurun_out
mkdir -p $O
if [ "$1" != "--skip-launch-list" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 2 --no-workloads > $O/r2_launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $O/r2_launches_bench.csv)"
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2_launches_box3.csv \
    python scripts/profile_once.py box3 > $O/r2_launches_box3.log 2>&1
echo "box3 launch list rc=$? lines=$(wc -l < $O/r2_launches_box3.csv)"
for sec in lsm box5 roi mean distill; do
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/r2_$sec -f \
      python scripts/profile_once.py $sec > $O/r2_ncu_$sec.log 2>&1
  echo "$sec rc=$? $(tail -1 $O/r2_ncu_$sec.log)"
done
python scripts/ncu_summary.py $O/r2_lsm.ncu-rep $O/r2_box5.ncu-rep $O/r2_roi.ncu-rep $O/r2_mean.ncu-rep $O/r2_distill.ncu-rep > $O/r2_ncu_full_summary.txt 2>&1
wc -l $O/r2_ncu_full_summary.txt
rm -f $O/r2_box5.ncu-rep $O/r2_mean.ncu-rep $O/r2_distill.ncu-rep
du -sh $O
