import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ganecdotes_b200 import _lib as L
from ganecdotes_b200.stylegan2.model import Generator
torch.manual_seed(0)
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
if len(sys.argv) > 2 and sys.argv[2] == "pidray":       # BagGAN channel map 512 ... 16 (BASELINE config 4)
    from ganecdotes_b200.baggan import baggan_channels
    g = Generator(256, 512, 8, channels=baggan_channels()).cuda()
else:
    g = Generator(256, 512, 8).cuda()
g.tag_layers = True
lat = torch.randn(b, g.n_latent, 512, device="cuda")
for _ in range(2): g.synthesize(lat, None, False)
torch.cuda.synchronize()
L.event_log = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): g.synthesize(lat, None, False)
e1.record(); torch.cuda.synchronize()
print(f"synthesis B={b}: {e0.elapsed_time(e1)/3:.3f} ms per forward ({90.24*b/(e0.elapsed_time(e1)/3)/1e0:.1f} GFLOP/ms alg)")
agg = {}
for name, a, c, work, _ in L.event_log:
    t = agg.setdefault(name, [0.0, 0.0, 0]); t[0] += a.elapsed_time(c); t[1] += work; t[2] += 1
for name, (ms, work, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    unit = "TF" if "modconv" in name else "GB/s"
    rate = work / ms / (1e9 if unit == "TF" else 1e6)
    print(f"{name:28s} n={n:3d} {ms/3:8.3f} ms/fwd  {rate:8.1f} {unit}")
