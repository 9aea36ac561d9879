# usage: bash tools/run_prof.sh <scene> <tag>   -> gpurun_out/<tag>.ncu-rep (k_stack_frames only, full set)
cd /root/repo
python tools/prof_s2.py 8 $1 > gpurun_out/$2_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_stack_frames -c 2 -o gpurun_out/$2 -f python tools/prof_s2.py 8 $1 > gpurun_out/$2_ncu.log 2>&1
tail -2 gpurun_out/$2_plain.log
