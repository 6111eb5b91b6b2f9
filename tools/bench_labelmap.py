"""Label-map throughput of the evaluate.py inference path on one GPU:
generator forward -> per-pixel codes (per-resolution projection) -> OneShotSegmentor head -> arg-max.
Prints one JSON line; the CPU leg times the oracle port on the host cores for one image."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganecdotes_b200 import _lib as L
from ganecdotes_b200.hfc_with_swav import OneShotSegmentor, engine as E
from ganecdotes_b200.stylegan2.model import Generator


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    size_key = sys.argv[2] if len(sys.argv) > 2 else "XXS"
    model = sys.argv[3] if len(sys.argv) > 3 else "ffhq"          # ffhq | pidray (BagGAN channel map, hlen 2528)
    torch.manual_seed(42)
    if model == "pidray":
        from ganecdotes_b200.baggan import baggan_channels
        gen = Generator(256, 512, 8, channels=baggan_channels()).cuda()
        hlen, trunc = 2528, 0.9
    else:
        gen = Generator(256, 512, 8).cuda()
        hlen, trunc = 5376, 0.7
    head = OneShotSegmentor(512, 12, size=size_key).cuda().eval()
    wp = (torch.randn(512, hlen) / hlen ** 0.5).cuda()
    with torch.no_grad():
        mean_latent = gen.style(torch.randn(1024, 512).cuda()).mean(0, keepdim=True)
        w = gen.style(torch.randn(b, 512).cuda())

    def step():
        preds, _, planes = E.predict_codes(gen, wp, w, mean_latent, trunc, hlen, want_planes=True)
        return head.predict_labels(preds, planes)

    for _ in range(3):
        labels = step()
    torch.cuda.synchronize()
    L.event_log = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 10
    for _ in range(steps):
        labels = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    agg = {}
    for name, a, c, work in L.event_log:
        t = agg.setdefault(name, [0.0, 0])
        t[0] += a.elapsed_time(c) / steps
        t[1] += 1
    L.event_log = None
    # end to end: host latents in, host label maps out
    zs = torch.randn(b, 512).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.no_grad():
            w2 = gen.style(zs.cuda(non_blocking=True))
        preds, _, planes = E.predict_codes(gen, wp, w2, mean_latent, trunc, hlen, want_planes=True)
        host = head.predict_labels(preds, planes).cpu()
    dt = (time.perf_counter() - t0) / 5
    # CPU port, one image
    from oracle import ganecdotes_oracle as O
    sd = {k: v.detach().cpu() for k, v in gen.state_dict().items()}
    seg = {k: v.detach().cpu() for k, v in head.state_dict().items()}
    t0 = time.perf_counter()
    p1, _ = O.predict_codes(sd, w[:1].cpu(), mean_latent.cpu(), trunc, wp.cpu(), hlen)
    y = O.one_shot_segmentor(seg, p1, 12, size_key)
    ref_labels = y.max(1)[1]
    cpu_s = time.perf_counter() - t0
    agree = (ref_labels == labels[:1].cpu()).float().mean().item()
    print(json.dumps({
        "metric": "label-map pixels/sec (" + model + "-256 predict_swav_codes + OneShotSegmentor " + size_key + " + argmax)",
        "value": b * 65536 / (ms * 1e-3), "unit": "pixels/s", "images_per_call": b, "ms_per_call": ms,
        "e2e": {"value": b * 65536 / dt, "unit": "pixels/s", "h2d_bytes_per_call": b * 512 * 4,
                "d2h_bytes_per_call": b * 65536 * 8},
        "cpu_baseline": {"value": 65536 / cpu_s, "unit": "pixels/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "1 image"},
        "label_agreement_with_cpu_port_image0": agree,
        "stages_ms": {k: round(v[0], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])},
    }))


if __name__ == "__main__":
    main()
