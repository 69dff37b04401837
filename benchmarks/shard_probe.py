import sys, faulthandler
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from ggmlsharp_b200 import native as N
import test_gpu_rowsplit_inproc as T
print("gpus", T._ngpu(), flush=True)
one, enc, X, dims = T._run(1, (N.Q4_0, N.Q4_0, N.Q4_0), split=0, cache=False)
print("single ok", flush=True)
two, _, _, _ = T._run(1, (N.Q4_0, N.Q4_0, N.Q4_0), split=2, cache=False)
print("split ok", all(np.array_equal(one[k], two[k]) for k in one), flush=True)
