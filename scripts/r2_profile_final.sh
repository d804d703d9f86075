O=gpurun_out
mkdir -p $O
for sec in lsm box5; do
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $O/r2b_$sec -f python scripts/profile_once.py $sec > $O/r2b_ncu_$sec.log 2>&1
  echo "$sec rc=$? $(tail -1 $O/r2b_ncu_$sec.log)"
done
python scripts/ncu_summary.py $O/r2b_lsm.ncu-rep $O/r2b_box5.ncu-rep > $O/r2b_ncu_full_summary.txt 2>&1
rm -f $O/r2b_box5.ncu-rep
cat $O/r2b_ncu_full_summary.txt | cut -c1-250
