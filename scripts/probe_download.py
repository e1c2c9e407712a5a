import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from singlet_b200 import api, synth
A = synth.synth_scipy(2000, 300000, 0.01, seed=5)
w0 = synth.w_init(20, 2000, seed=6)
h = api.Handle(0)
os.environ["SGL_UPLOAD_THREADS"] = "1"   # one worker: the old single-copy download path (n_workers < 2)
a = api.c_nmf(A, None, 0.0, 3, False, 0.01, 0.01, 0, 0, 0, w0, h)
h.close()
del os.environ["SGL_UPLOAD_THREADS"]
h2 = api.Handle(0)
b = api.c_nmf(A, None, 0.0, 3, False, 0.01, 0.01, 0, 0, 0, w0, h2)
print("h equal:", np.array_equal(a["h"], b["h"]), "w equal:", np.array_equal(a["w"], b["w"]), a["h"].shape)
