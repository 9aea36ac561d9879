# usage: bash tools/ab_pair.sh <tag> <scenes...>   -- current library vs libmovis_b200_base.so, interleaved, same box
cd /root/repo
tag=$1; shift
for rep in 1 2; do
for s in "$@"; do
  MOVIS_B200_LIB=movis_b200/csrc/libmovis_b200_base.so python tools/ab_s2.py $s 3 2>&1 | grep "^s" | sed 's/^/base /'
  python tools/ab_s2.py $s 3 2>&1 | grep "^s" | sed 's/^/new  /'
done; done | cut -c1-60 | tee gpurun_out/abpair_$tag.log
