# usage: bash tools/runv.sh NSIDE LMAX "variant list" "transform list" ; cur = the in-tree build
ns=${1:-2048}; lm=${2:-4095}
for v in ${3:-cur}; do
  if [ "$v" = cur ]; then unset HPSHT_B200_LIB; else export HPSHT_B200_LIB=$PWD/healpixmpi.jl_b200/variants/libhpsht_b200_$v.so; fi
  echo "== $v"; for t in ${4:-2_adj 1_adj 2_a2m 1_a2m}; do python tools/run_one.py $ns $lm ${t%_*} ${t#*_} 3 | tail -1; done
done
