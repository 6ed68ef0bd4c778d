import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_train_step_gpu import _run_case, CASES
which = sys.argv[1] if len(sys.argv) > 1 else "c1_softmax_sgd"
for name, lt, extra, gs, lr in CASES:
    if name != which: continue
    r = _run_case(lt, extra, gs, lr)
    print({k: v for k, v in r.items() if not isinstance(v, dict)})
    for n, e in r["grad_err"].items():
        print("  %-40s vs-emu %.4g   vs-fp64 %.4g  cos64 %.5f   param %.3g" % (n, e, r["grad_err64"].get(n, 0), r["grad_cos64"].get(n, 1), r["param_err"].get(n, 0)))
