cd /root/repo
(python tools/ab_s2.py s1 3; python tools/ab_s2.py s2d 3) > gpurun_out/ab3.log 2>&1
for s in s1 s3 s4 s5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_list_$s.csv python tools/prof_s2.py 8 $s > gpurun_out/r2_list_$s.log 2>&1
done
grep "^s" gpurun_out/ab3.log
