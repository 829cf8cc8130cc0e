#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags...]: builds healpixmpi.jl_b200/variants/libhpsht_b200_NAME.so from the
# current sources with extra preprocessor flags (kernel experiments; select with HPSHT_B200_LIB=<path>).
set -e
here="$(cd "$(dirname "$0")/.." && pwd)"
src="$here/healpixmpi.jl_b200/csrc"
out="$here/healpixmpi.jl_b200/variants"
name="$1"; shift
mkdir -p "$out/obj_$name"
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp,-Wno-unknown-pragmas -ccbin /usr/bin/g++"
for f in legendre ringfft api; do
  # ringfft / api objects are reused from the main build unless the flags name them
  if [ "$f" = "${VARIANT_FILE:-legendre}" ] || [ -n "$VARIANT_ALL" ]; then
    $NV "$@" -c "$src/$f.cu" -o "$out/obj_$name/$f.o" &
  else
    cp "$src/build/$f.o" "$out/obj_$name/$f.o"
  fi
done
wait
/usr/local/cuda/bin/nvcc -shared -o "$out/libhpsht_b200_$name.so" "$out/obj_$name"/*.o -Xcompiler -fopenmp -lgomp -ldl 2>/dev/null
rm -rf "$out/obj_$name"
echo "$out/libhpsht_b200_$name.so"
