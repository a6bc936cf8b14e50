import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdirt_b200 import _engine as E
from sdirt_b200.prescription import load_lens_json
recs, descs, head = load_lens_json(os.path.join(os.path.dirname(E.__file__), "lenses", "rf50mm.json"))
h = E.LensHandle(recs, 62.25)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
m = 200000
th = torch.rand(m, generator=g) * 2 * np.pi
rr = torch.sqrt(torch.rand(m, generator=g) * 6.019352912902832 ** 2)
pup = torch.stack([rr * torch.cos(th), rr * torch.sin(th)], -1).to(dev).contiguous()
lib = E.lib()
lib.sdirt_debug_strict_pair.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
for pt in ([-86.98888, 2320.7637, -12153.938], [-5909.853, -2731.2825, -17124.674], [-17.78, 849.34, -6574.4], [3.0, 2.0, -500.0]):
    p = torch.tensor(pt, device=dev)
    mm = torch.zeros(8, dtype=torch.int32, device=dev); ex = torch.zeros(16, device=dev)
    rc = lib.sdirt_debug_strict_pair(h.ptr if hasattr(h, "ptr") else h._h, 0.589, C.c_void_p(p.data_ptr()), C.c_void_p(pup.data_ptr()), m, 22.51324462890625,
                                     C.c_void_p(mm.data_ptr()), C.c_void_p(ex.data_ptr()), None)
    torch.cuda.synchronize()
    print(pt, "rc", rc, "mismatches [ox oy oz dx dy dz alive rays]:", mm.tolist())
    e = ex.cpu().numpy()
    print("   scalar", e[:6], "\n   packed", e[6:12], " sample", e[12:14])
