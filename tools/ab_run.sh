set -e
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.log 2>&1 || { tail -20 gpurun_out/ab_pytest.log; exit 1; }
tail -2 gpurun_out/ab_pytest.log
CASES=cfg2:I:0,cfg2:U:0,cfg1:I:0,cfg3:I:0,cfg4enc:I:0
for i in 1 2; do
  echo "== base"; MSDA_B200_LIB=$PWD/open-groundingdino_b200/build/lib_base.so python tools/sweep.py --cases $CASES --iters 20
  echo "== new"; python tools/sweep.py --cases $CASES --iters 20
done
