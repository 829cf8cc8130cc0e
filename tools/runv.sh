# usage: bash tools/runv.sh NSIDE LMAX "variant list" ; "" = the in-tree build
ns=${1:-2048}; lm=${2:-4095}
for v in ${3:-cur}; do
  if [ "$v" = cur ]; then unset HPSHT_B200_LIB; else export HPSHT_B200_LIB=$PWD/healpixmpi.jl_b200/variants/libhpsht_b200_$v.so; fi
  echo "== $v"; for t in "2 adj" "1 adj" "2 a2m" "1 a2m"; do python tools/run_one.py $ns $lm $t 3 | tail -1; done
done
