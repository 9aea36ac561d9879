# usage: bash tools/ab_all.sh <tag> [lib]   -- device us/frame of every scene with the given library build
cd /root/repo
[ -n "$2" ] && export MOVIS_B200_LIB=$2
for s in s2 s1 s4 s2d s3 s5; do python tools/ab_s2.py $s 3 2>&1 | grep "^s" ; done | tee gpurun_out/ab_$1.log
