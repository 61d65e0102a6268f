set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2_run8.log; : > $O
for mode in 1 2 3; do
  echo "== pw mode $mode (bit0: no dot products, bit1: no weight traffic)" >> $O
  QWEN_MEGA_MODE=$mode timeout 200 python scripts/phase_profile.py 4b 4096 2>&1 | grep -v "^\[" >> $O
done
python - >> $O 2>&1 <<'PY'
import ctypes as C, sys
sys.path.insert(0, '.')
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
ql.lib.qwen_cuda_int8_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
t = C.c_float(0)
for iters in (2000, 20000):
    rc = ql.lib.qwen_cuda_int8_peak(iters, 5, C.byref(t))
    print("int8 peak iters", iters, "rc", rc, "TOPS", t.value, ql.err() if rc else "")
PY
