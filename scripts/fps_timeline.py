"""Per-round timeline of the async-cluster FPS kernel (thread 0 of CTA 0, rounds 100..103), in SM cycles."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camliflow_b200 import native
from camliflow_b200.csrc import furthest_point_sampling
dev = torch.device("cuda:0")
pc = ((torch.rand(2, 8192, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 10).to(dev)
buf = torch.zeros(32, dtype=torch.int64, device=dev)
furthest_point_sampling(pc, 4096); torch.cuda.synchronize()
native.lib().camli_fps_set_timeline(ctypes.c_void_p(buf.data_ptr()))
furthest_point_sampling(pc, 4096); torch.cuda.synchronize()
native.lib().camli_fps_set_timeline(ctypes.c_void_p(0))
t = buf.cpu().tolist()
for r in range(4):
    a = t[r * 5:(r + 1) * 5]
    nxt = t[(r + 1) * 5] if r < 3 else None
    print("round %d: compute %d | redux+send %d | wait %d | pick %d | -> next round top %s" %
          (100 + r, a[1] - a[0], a[2] - a[1], a[3] - a[2], a[4] - a[3], (nxt - a[4]) if nxt else "-"))
