"""Experiment (debug build): where do the roles of assign_tc_kernel wait?  TPQ_B200_LIB=.../libtpq_b200_dbg.so"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchpq_b200 as T
lib = T._lib.lib
f = lib.tpq_debug_set_tc_wait; f.restype = None; f.argtypes = [ctypes.c_void_p]
l, d, n, k = 64, 64, 1_000_000, 256
data = torch.randn(l, d, n, device="cuda"); cent = data[:, :, :k].contiguous()
for ev in (False, True):
    T.fn.max_sim(data, cent, exact=False, exact_values=ev); torch.cuda.synchronize()
    w = torch.zeros(8, dtype=torch.int64, device="cuda"); f(ctypes.c_void_p(w.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); T.fn.max_sim(data, cent, exact=False, exact_values=ev); e1.record(); torch.cuda.synchronize()
    f(ctypes.c_void_p(0))
    w = w.cpu().tolist(); ctas = 148; tot = w[7] / ctas
    names = ["producer: empty_a", "mma: tmem_empty", "mma: full_a", "argmax(x8 warps): tmem_full", "argmax(x8): cand_empty",
             "finisher(x4 per tile): full_a", "finisher(x4): cand_full"]
    div = [1, 1, 1, 8, 8, 4, 4]
    print(f"exact_values={ev} ms={e0.elapsed_time(e1):.3f} cycles/CTA={tot:.0f}")
    for nm, v, dv in zip(names, w[:7], div):
        print(f"   {nm:36s} {v / ctas / dv / tot * 100:6.1f} % of the CTA's life (per warp)")
