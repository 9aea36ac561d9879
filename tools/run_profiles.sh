# Round-2 profile artefacts (one B200): every command first runs plain, then under ncu.
cd /root/repo
O=gpurun_out
python tools/prof_s2.py 8 s2 > $O/r02_prof_s2_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_stack_frames -s 1 -c 1 -o $O/r02_ncu_full_k_stack_frames -f python tools/prof_s2.py 8 s2 > $O/r02_ncu_full.log 2>&1
for s in s2 s1 s3 s4 s5; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_launches_$s.csv python tools/prof_s2.py 8 $s > $O/r02_launches_$s.log 2>&1
done
python bench.py --quick --steps 2 --warmup 3 > $O/r02_bench_quick_plain.json 2> $O/r02_bench_quick_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --quick --steps 2 --warmup 3 > $O/r02_launches_bench.log 2>&1
python tools/bench_kernels.py > $O/r02_kernels.txt 2>&1
python tools/prof_chroma.py && ncu --set full --clock-control none -k regex:k_chroma -c 1 -o $O/r02_ncu_full_k_chroma_key -f python tools/prof_chroma.py > $O/r02_ncu_chroma.log 2>&1
tail -2 $O/r02_prof_s2_plain.log; cat $O/r02_kernels.txt
