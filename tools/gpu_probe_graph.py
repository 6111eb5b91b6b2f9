"""Probe: the ffhq-256 training step (8 latents) as stream launches vs one CUDA-graph replay (engine.TrainGraph)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ganecdotes_b200 import _lib as L  # noqa: E402
from ganecdotes_b200.hfc_with_swav import engine as E  # noqa: E402
from ganecdotes_b200.stylegan2.model import Generator  # noqa: E402

cfg = bench.FFHQ
dev = torch.device("cuda")
torch.manual_seed(42)
gen = Generator(256, 512, 8).to(dev)
mean_latent = gen.style(torch.randn(4096, 512).to(dev)).mean(0, keepdim=True)
proj = torch.nn.Linear(5376, 512, bias=False).to(dev)
proto = torch.nn.Linear(512, 5000).to(dev)
head = E.SwavHead(proj.weight.data, proto.weight.data, proto.bias.data, 0.01, 0.9, 0.01, 3, 1)
scfg = E.StepConfig(hlen=5376, patch_size=20000, num_patches=5, niters=10, eps=0.005, temperature=0.01, truncation=0.7,
                    perturb_std=[1.0] * 6)
ws = L.SinkhornWorkspace(5000, dev)
b = 8


def draw(seed):
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    view = lambda: E.ViewDraws(layer_no=[int(rs.randint(6)) for _ in range(b)], pert_z=torch.randn(b, 12, 512, generator=g),
                               angle=[float(rs.uniform(-10, 10)) for _ in range(b)], flip=[bool(rs.rand() < 0.5) for _ in range(b)])
    return E.StepDraws(z=torch.randn(b, 512, generator=g), view_s=view(), view_t=view(),
                       perms=[[torch.randperm(65536, generator=g) for _ in range(b)] for _ in range(5)])


inputs = [E.prepare_step_inputs(gen, draw(i), scfg, dev) for i in range(14)]
for i in range(3):
    E.swav_train_step_device(gen, head, mean_latent, inputs[i], scfg, None, ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3, 8):
    l_eager = E.swav_train_step_device(gen, head, mean_latent, inputs[i], scfg, None, ws)
e1.record()
torch.cuda.synchronize()
print("eager  ms/step", e0.elapsed_time(e1) / 5, float(l_eager))
g = E.TrainGraph(gen, head, mean_latent, inputs[8], scfg, None, ws)
l = g(inputs[8])
torch.cuda.synchronize()
e0.record()
for i in range(9, 14):
    l_graph = g(inputs[i])
e1.record()
torch.cuda.synchronize()
print("graph  ms/step", e0.elapsed_time(e1) / 5, float(l_graph), "launches/replay", g.launches_per_replay)
