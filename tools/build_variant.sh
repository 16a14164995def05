#!/bin/bash
# Experimental kernel variants for A/B runs on the GPU box: rebuilds the small-path translation units with extra -D flags
# and links them with the stock objects into vec-ode_b200/variants/libvecode_b200_<name>.so (select it at run time with
# VECODE_B200_SO=<path>). Usage: tools/build_variant.sh <name> "<nvcc -D flags>" [tu ...]   (default TU: rk_small_vdp)
set -e
NAME=$1; FLAGS=$2; shift 2
TUS=${@:-rk_small_vdp}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/vec-ode_b200/csrc; OBJ=$ROOT/build/csrc; VOBJ=$ROOT/build/variants/$NAME
mkdir -p $VOBJ $ROOT/vec-ode_b200/variants
make -C $SRC -j8 > /dev/null
OBJS=""
for o in $OBJ/*.o; do
  b=$(basename $o .o)
  if [[ " $TUS " == *" $b "* ]]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I$OBJ -Xcompiler -fPIC -Xptxas -v $FLAGS -c $SRC/$b.cu -o $VOBJ/$b.o 2> $VOBJ/$b.ptxas.log
    OBJS="$OBJS $VOBJ/$b.o"
  else
    OBJS="$OBJS $o"
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $ROOT/vec-ode_b200/variants/libvecode_b200_$NAME.so $OBJS -lcudart -ldl
echo "built variant $NAME ($FLAGS)"
