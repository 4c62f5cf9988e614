# same-box A/B of two builds of the library: MSDA_B200_LIB points the loader at the other one
set -e
CASES=${CASES:-cfg2:I:0,cfg2:U:0,cfg1:I:0,cfg4enc:I:0}
if [ -n "$AB_TESTS" ]; then
  timeout 900 python -m pytest $AB_TESTS -x -q > gpurun_out/ab_pytest.log 2>&1 || { tail -30 gpurun_out/ab_pytest.log; exit 1; }
  tail -2 gpurun_out/ab_pytest.log
fi
for i in 1 2; do
  if [ -z "$AB_NO_BASE" ]; then echo "== base"; MSDA_B200_LIB=$PWD/open-groundingdino_b200/build/lib_base.so python tools/sweep.py --cases $CASES --iters 20; fi
  echo "== new"; python tools/sweep.py --cases $CASES --iters 20
done
