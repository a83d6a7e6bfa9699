"""Find the first libgnm / torch call that invalidates a CUDA-graph capture of the Trainer step.
usage: capture_debug.py [fixture] [fold_tails 0|1]"""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Golden, SEED0
from graph_neural_mapping_b200 import engine, lib, ops
from graph_neural_mapping_b200.driver import Trainer
from graph_neural_mapping_b200.models import GIN_InfoMaxReg

name = sys.argv[1] if len(sys.argv) > 1 else "schaefer400_b16_noeps"
engine.FOLD_BN_TAILS = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True
g = Golden(name); c = g.cfg
dev = torch.device("cuda")
m = GIN_InfoMaxReg(c["num_layers"], c["num_mlp_layers"], c["input_dim"], c["hidden_dim"], 2, 0.0, c["learn_eps"],
                   c["graph_pooling_type"], c["neighbor_pooling_type"], dev)
m.load_state_dict(g.state_dict()); m = m.to(dev); m.train()


def status():
    v = ctypes.c_int(0)
    lib.load().gnm_stream_capture_status(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), ctypes.byref(v))
    return v.value


seen = {"bad": None}
for fname in dir(ops):
    fn = getattr(ops, fname)
    if callable(fn) and not fname.startswith("_") and fn.__module__ == ops.__name__ and not isinstance(fn, type):
        def make(fname, fn):
            def wrapped(*a, **k):
                before = status()
                out = fn(*a, **k)
                after = status()
                if before == 1 and after == 2 and seen["bad"] is None:
                    seen["bad"] = fname
                    print("CAPTURE INVALIDATED inside ops.%s" % fname, flush=True)
                elif before == 2 and seen["bad"] is None:
                    seen["bad"] = "before " + fname
                    print("capture was already invalid BEFORE ops.%s (a torch op in between)" % fname, flush=True)
                return out
            return wrapped
        setattr(ops, fname, make(fname, fn))
tr = Trainer(m, lr=0.005, beta=c["beta"])
graphs = g.graphs()
for i in range(4):
    np.random.seed(1)
    try:
        print("step", i, float(tr.step(graphs)), flush=True)
    except Exception as e:
        print("step", i, "FAILED:", str(e).splitlines()[0], flush=True)
        break
print("fold_tails", engine.FOLD_BN_TAILS, "first offender:", seen["bad"])
