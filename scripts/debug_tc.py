import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torchpq_b200 as T
torch.manual_seed(0)
l, d, n, k = 1, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 256, int(sys.argv[2]) if len(sys.argv) > 2 else 256
data, cent = torch.randn(l, d, n).cuda(), torch.randn(l, d, k).cuda()
dbg = torch.zeros(128, 256, device="cuda")
lib = ctypes.CDLL(T._lib.LIB_PATH)
lib.tpq_debug_set_tc_dump.argtypes = [ctypes.c_void_p]
lib.tpq_debug_set_tc_dump(dbg.data_ptr())
sim, lab = T.fn.max_sim(data, cent, exact=False)
torch.cuda.synchronize()
ref = (data[0, :, :128].T @ cent[0]).cpu().numpy()          # [128, k]
got = dbg.cpu().numpy()[:, :k]
err = np.abs(got - ref)
print("max err", err.max(), "mean", err.mean(), "ref scale", np.abs(ref).mean())
ok = err < 0.05
print("row match rate per 32-row quad:", [float(ok[r:r+32].mean()) for r in range(0, 128, 32)])
print("col match rate per 32-col group:", [round(float(ok[:, c:c+32].mean()), 2) for c in range(0, k, 32)])
# try to identify structure: does got equal partial sums over some k-blocks?
for kb in range(d // 8):
    part = (data[0, kb*8:(kb+1)*8, :128].T @ cent[0, kb*8:(kb+1)*8]).cpu().numpy()
    print("kb", kb, "corr with got", float(np.corrcoef(part.ravel(), got.ravel())[0, 1]))
print("corr full", float(np.corrcoef(ref.ravel(), got.ravel())[0, 1]))
print(got[:4, :8]); print(ref[:4, :8])
