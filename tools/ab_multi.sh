# same-box comparison of several builds of the library (open-groundingdino_b200/build/lib_*.so), two rounds
set -e
CASES=${CASES:-cfg2:I:0,cfg2:U:0}
for i in 1 2; do
  for lib in open-groundingdino_b200/build/lib_*.so; do
    echo "== $(basename $lib)"; MSDA_B200_LIB=$PWD/$lib python tools/sweep.py --cases $CASES --iters 20
  done
done
