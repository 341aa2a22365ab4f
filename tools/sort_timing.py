"""Debug: per-phase globaltimer stamps of the cluster sort (HB_CS_TIMING=1), bench-shaped input."""
import os, sys, ctypes as C
os.environ['HB_CS_TIMING'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hybridbackend_b200 as hb
import bench
dev = 'cuda:0'
B, D = 65536, 32
sizes = [min(n, 2000000) for n in bench.CRITEO_SIZES]
sizes[0] = 39884406
rng = np.random.RandomState(1)
tabs = [torch.zeros(n, D, device=dev) for n in sizes]
gl = hb.embedding.GroupLookup(tabs, ['mean'] * 26, overlap_backward_sort=False)
ids = [torch.from_numpy(bench.gen_ids_numpy(rng, B, n, 'zipf', 1.05)).to(dev) for n in sizes]
grad = torch.randn(B, 26 * D, device=dev)
opt = hb.training.Adagrad(0.01)
for _ in range(3):
  gl.forward(ids); gl.backward_update(grad, opt)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 64)()
L = hb._lib.lib()
L.hbDebugSortTiming.argtypes = [C.c_void_p]
print('rc', L.hbDebugSortTiming(buf))
raw = [int(x) for x in buf]
t = [x for x in raw[:24] if x]
print('cluster end times (us after start of cluster 0):', [round((x - raw[0]) / 1e3, 1) for x in raw[24:50]])
print('passes per feature:', [1 if n + 2 <= 512 else (2 if n + 2 <= 262144 else 3) for n in sizes])
print([round((b - a) / 1e3, 2) for a, b in zip(t, t[1:])], 'total us', (t[-1] - t[0]) / 1e3)
