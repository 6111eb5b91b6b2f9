#!/usr/bin/env python
"""Benchmark of the ffhq-256 hfc_with_swav pretrain step (BASELINE.json metric:
per-pixel feature vectors / s).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...    # CPU oracle port on the host cores

One "step" = one optimiser step of SwAVClustering.pretrain over B latents per GPU
(weak scaling): 2 latent-perturbed views x 5 patches x 20000 random pixels per latent,
D = 5376, C = 512, K = 5000 prototypes, 10 Sinkhorn iterations, LARC+SGD - i.e.
B * 200000 per-pixel feature vectors per GPU per step, synthetic latents, random-init
StyleGAN2-256 + head (seed 42).  `value` is timed on the device with inputs resident in
HBM; `e2e` goes through the public step API from pinned host buffers (latents, random
draws, sampled-pixel indices) and reads the loss back every step.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FFHQ = dict(size=256, style_dim=512, n_mlp=8, hlen=5376, nclasses=512, nprototypes=5000, patch=20000, npatch=5,
            niters=10, eps=0.005, temperature=0.01, truncation=0.7, n_layers=6, perturb_std=[1.0] * 6,
            lr=0.01, momentum=0.9, trust=0.01)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """`nvidia-smi -lms 100` running beside the timed region: SM clocks and throttle reasons."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.t0, self.t1 = index, None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu=timestamp,{self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def summary(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                out = ""
            import datetime
            for line in out.strip().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) < 7:
                    continue
                try:
                    ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except Exception:
                    continue
                if self.t0 is not None and (ts < self.t0 - 0.05 or ts > self.t1 + 0.05):
                    continue
                rows.append(c[1:])
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if len(r) > 2 + i and r[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows)}


# ----------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ----------------------------------------------------------------------------------------

def cpu_reference_step(state, cfg, seed):
    """A bounded sample of the workload on the CPU: 1 latent, 2 views, 1 of the 5 patches
    (20000 px) through the full step (synthesis, upsample+concat, rotate/flip, sampling,
    projection, prototypes, 2x Sinkhorn, loss, backward, LARC+SGD).  Returns vectors done."""
    from oracle import ganecdotes_oracle as O
    sd, mean_latent, wp, wk, bk = state
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    z = torch.randn(1, cfg["style_dim"], generator=g)
    w = O.style_mlp(sd, z)
    rows = {}
    hw = cfg["size"] ** 2
    perm = torch.randperm(hw, generator=g)
    for v in "st":
        hf, _ = O.view_features(sd, w, mean_latent, cfg["truncation"], int(rs.randint(cfg["n_layers"])),
                                torch.randn(2 * cfg["n_layers"], cfg["style_dim"], generator=g), cfg["n_layers"],
                                cfg["perturb_std"], cfg["hlen"])
        hf = O.rotate_flip(hf, float(rs.uniform(-10, 10)), bool(rs.rand() < 0.5))
        rows[v] = [O.sample_rows(hf, perm, cfg["patch"])]
        del hf
    out = O.swav_step(rows["s"], rows["t"], wp, wk, bk, cfg["niters"], cfg["eps"], cfg["temperature"], None,
                      cfg["lr"], cfg["momentum"], cfg["trust"])
    state[2], state[3], state[4] = out["params"]
    return 2 * cfg["patch"]


def make_cpu_state(cfg):
    from oracle import ganecdotes_oracle as O
    torch.manual_seed(42)
    sd = O.init_generator_state(cfg["size"], cfg["style_dim"], cfg["n_mlp"], 42, randomize_small=False)
    mean_latent = O.style_mlp(sd, torch.randn(4096, cfg["style_dim"])).mean(0, keepdim=True)
    proj = torch.nn.Linear(cfg["hlen"], cfg["nclasses"], bias=False)
    proto = torch.nn.Linear(cfg["nclasses"], cfg["nprototypes"])
    return [sd, mean_latent, proj.weight.data.clone(), proto.weight.data.clone(), proto.bias.data.clone()]


def time_cpu(cfg, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    state = make_cpu_state(cfg)
    with torch.no_grad():
        pass
    for i in range(warmup):
        cpu_reference_step(state, cfg, 1000 + i)
    t0 = time.perf_counter()
    vecs = 0
    for i in range(steps):
        vecs += cpu_reference_step(state, cfg, 2000 + i)
    dt = time.perf_counter() - t0
    return vecs / dt, dt / max(steps, 1), torch.get_num_threads()


def workload_name(b):
    return (f"ffhq-256 hfc_with_swav pretrain step: {b} latents/GPU x 2 views x 5 patches x 20000 px, "
            f"D=5376 C=512 K=5000, Sinkhorn 10 it, fwd+bwd+LARC/SGD")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = FFHQ
    value, sec_per_step, cores = time_cpu(cfg, args.steps, max(args.warmup, 1))
    sample = "per step: 1 latent x 2 views x 1 of 5 patches (20000 px) through the full step on the host CPU"
    line = {
        "impl": "reference", "metric": "per-pixel feature vectors/sec (ffhq-256 SwAV step)", "value": value,
        "unit": "vectors/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.latents_per_gpu), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "vectors/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------

def run_ours(args):
    import torch.distributed as dist
    from ganecdotes_b200 import _lib as L
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.stylegan2.model import Generator

    cfg = FFHQ
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = E.DistGroup(dist.group.WORLD, rank, world)
    L.load()
    b = args.latents_per_gpu

    # ---- model + head, identical on every rank (seed 42 = the reference's seed, lib/util/util.py:21-24)
    torch.manual_seed(42)
    np.random.seed(42)
    gen = Generator(cfg["size"], cfg["style_dim"], cfg["n_mlp"]).to(dev)
    gen.passes = args.passes_fwd
    with torch.no_grad():
        mean_latent = gen.style(torch.randn(4096, cfg["style_dim"]).to(dev)).mean(0, keepdim=True)
    proj = torch.nn.Linear(cfg["hlen"], cfg["nclasses"], bias=False).to(dev)
    proto = torch.nn.Linear(cfg["nclasses"], cfg["nprototypes"]).to(dev)
    head = E.SwavHead(proj.weight.data, proto.weight.data, proto.bias.data, cfg["lr"], cfg["momentum"], cfg["trust"],
                      args.passes_fwd, args.passes_bwd, proto_f16=args.proto_f16)
    scfg = E.StepConfig(hlen=cfg["hlen"], patch_size=cfg["patch"], num_patches=cfg["npatch"], niters=cfg["niters"],
                        eps=cfg["eps"], temperature=cfg["temperature"], truncation=cfg["truncation"],
                        perturb_std=cfg["perturb_std"])
    ws = L.SinkhornWorkspace(cfg["nprototypes"], dev)

    def draw(seed):
        g = torch.Generator().manual_seed(seed)
        rs = np.random.RandomState(seed)
        hw = cfg["size"] ** 2

        def view():
            return E.ViewDraws(layer_no=[int(rs.randint(cfg["n_layers"])) for _ in range(b)],
                               pert_z=torch.randn(b, 2 * cfg["n_layers"], cfg["style_dim"], generator=g).pin_memory(),
                               angle=[float(rs.uniform(-10, 10)) for _ in range(b)],
                               flip=[bool(rs.rand() < 0.5) for _ in range(b)])
        return E.StepDraws(z=torch.randn(b, cfg["style_dim"], generator=g).pin_memory(), view_s=view(), view_t=view(),
                           perms=[[torch.randperm(hw, generator=g) for _ in range(b)] for _ in range(cfg["npatch"])])

    nsteps = args.warmup + args.steps
    draws = [draw(10_000 * (rank + 1) + i) for i in range(nsteps)]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-timed region
    inputs = [E.prepare_step_inputs(gen, d, scfg, dev) for d in draws]       # resident in HBM
    for i in range(args.warmup):
        E.swav_train_step_device(gen, head, mean_latent, inputs[i], scfg, group, ws)
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sync()
    sampler.mark_begin()
    L.launch_count = 0
    L.event_log = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for i in range(args.steps):
        loss = E.swav_train_step_device(gen, head, mean_latent, inputs[args.warmup + i], scfg, group, ws)
    e1.record()
    sync()
    sampler.mark_end()
    clocks = sampler.summary() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = L.launch_count
    log = L.event_log
    L.event_log = None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    vec_per_step = world * b * 2 * cfg["npatch"] * cfg["patch"]
    value = vec_per_step * args.steps / (ms * 1e-3)
    final_loss = float(loss)

    # per-kernel totals (CUDA events recorded around every launch of the timed region)
    stages = {}
    for name, a, c, work in log:
        s = stages.setdefault(name, [0.0, 0.0, 0])
        s[0] += a.elapsed_time(c)
        s[1] += work
        s[2] += 1
    pk = peaks()
    tensor_bound = {"gemm", "modconv"}
    # tensor-core passes issued per algorithmic FLOP (split-bf16 = 3 MMAs per product)
    issued = {"gemm_prototype_fwd": 1 if args.proto_f16 else 3, "gemm_projection_fwd": args.passes_fwd,
              "modconv": args.passes_fwd, "modconv_up": args.passes_fwd, "gemm_dzn_bwd": args.passes_bwd,
              "gemm_gproto_bwd": args.passes_bwd, "gemm_gproj_bwd": args.passes_bwd}
    stage_rows = []
    for name, (tms, work, cnt) in sorted(stages.items(), key=lambda kv: -kv[1][0]):
        is_tensor = name.split("_")[0] in tensor_bound
        peak = pk["tf_sustained"] if is_tensor else pk["hbm"]
        achieved = (work / (tms * 1e-3)) / (1e12 if is_tensor else 1e9) if tms > 0 else 0.0
        stage_rows.append({"kernel": name, "bound": "tensor" if is_tensor else "hbm", "launches": cnt,
                           "ms_per_step": tms / args.steps, "share": tms / ms, "achieved": achieved, "peak": peak,
                           "unit": "TFLOP/s" if is_tensor else "GB/s", "frac": achieved / peak})
        if is_tensor:
            stage_rows[-1]["mma_passes"] = issued.get(name, 1)
            stage_rows[-1]["frac_issued"] = issued.get(name, 1) * achieved / peak
    top = stage_rows[0] if stage_rows else None
    roofline = None
    if top:
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        traffic = None
        tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(top["kernel"], {}).get("dram_bytes_per_launch")
        tms, work, cnt = stages[top["kernel"]]
        roofline = {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
                    "unit": top["unit"], "frac": top["frac"], "traffic": traffic,
                    "algorithmic_per_launch": work / cnt, "launches": cnt, "ms_per_launch": tms / cnt,
                    "peak_source": pk["src"],
                    "note": "algorithmic FLOPs/bytes per launch / mean CUDA-event launch time; a 3-pass "
                            "split-bf16 GEMM issues 3x its algorithmic FLOPs on the tensor pipe"}

    # ---------------------------------------------------------------- the other score-GEMM operand mode
    # (same steps, device-timed, reported next to the headline; see DESIGN.md §4.2 for the tolerances)
    alt = None
    if not args.no_alt:
        head.proto_f16 = not args.proto_f16
        alt_steps = max(1, min(args.steps, 5))
        for i in range(2):
            E.swav_train_step_device(gen, head, mean_latent, inputs[i % len(inputs)], scfg, group, ws)
        sync()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for i in range(alt_steps):
            E.swav_train_step_device(gen, head, mean_latent, inputs[args.warmup + i], scfg, group, ws)
        a1.record()
        sync()
        alt_ms = a0.elapsed_time(a1)
        if world > 1:
            t = torch.tensor([alt_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            alt_ms = t.item()
        alt = {"score_gemm": "fp16x1" if head.proto_f16 else "bf16x3", "steps": alt_steps,
               "ms_per_step": alt_ms / alt_steps, "value": vec_per_step * alt_steps / (alt_ms * 1e-3),
               "unit": "vectors/s"}
        head.proto_f16 = args.proto_f16

    # ---------------------------------------------------------------- end-to-end region
    # public API from pinned host buffers: host bookkeeping + H2D of step i+1 are issued while
    # the GPU runs step i; the loss of every step is read back to the host.
    # Same schedule as SwAVClustering.pretrain: inputs of step i+1 go through pinned memory on a side
    # stream while step i computes; the loss of step i is read back once step i+1 has been launched.
    # The pipeline first runs `warmup` untimed steps (allocator pools of the side stream, pinned staging
    # buffers), then K timed steps in steady state: every timed step contains one staging (host bookkeeping
    # + pinned H2D of the NEXT step's inputs), one optimiser step and one device->host loss read.
    e2e_steps = max(1, args.steps)
    e2e_warm = max(1, args.warmup)
    side = torch.cuda.Stream(device=dev)
    sync()
    inp = E.prepare_step_inputs(gen, draws[args.warmup], scfg, dev, stream=side)
    h2d = inp.h2d_bytes
    pending = None
    host_ms = {"prepare": 0.0, "launch": 0.0, "wait": 0.0}
    t0 = None
    for i in range(e2e_warm + e2e_steps):
        if i == e2e_warm:
            if pending is not None:
                _ = float(pending)
                pending = None
            sync()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
        ta = time.perf_counter()
        l = E.swav_train_step_device(gen, head, mean_latent, inp, scfg, group, ws)
        tb = time.perf_counter()
        inp = E.prepare_step_inputs(gen, draws[args.warmup + (i + 1) % args.steps], scfg, dev, stream=side)
        tc = time.perf_counter()
        if pending is not None:
            _ = float(pending)                                         # device -> host read of the loss
        pending = l
        td = time.perf_counter()
        if i >= e2e_warm:
            host_ms["launch"] += (tb - ta) * 1e3
            host_ms["prepare"] += (tc - tb) * 1e3
            host_ms["wait"] += (td - tc) * 1e3
    _ = float(pending)
    sync()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    e2e_value = vec_per_step * e2e_steps / dt

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = time_cpu(cfg, 1, 1)
        cpu_base = {"value": v, "unit": "vectors/s", "cores": cores, "kind": "port",
                    "sample": "1 latent x 2 views x 1 of 5 patches (20000 px), full step, after 1 warm-up step",
                    "sec_per_sample_step": sec}
    if rank == 0:
        line = {
            "metric": "per-pixel feature vectors/sec (ffhq-256 SwAV step)", "value": value, "unit": "vectors/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": f"bf16x{args.passes_fwd}-split fwd" + (" (score GEMM fp16x1 on unit-norm operands)" if args.proto_f16 else "") +
                     f" / bf16x{args.passes_bwd} bwd operands, fp32 accumulate + fp32 everywhere else",
            "data": "synthetic",
            "config": {"workload": workload_name(b), "latents_per_gpu": b, "global_latents": b * world,
                       "vectors_per_step": vec_per_step, "l2": "inputs_larger_than_l2 (3.2 GB score matrices)",
                       "generator": "StyleGAN2-256 random init seed 42", "sinkhorn": "joint-batch (distributed)"},
            "roofline": roofline, "roofline_stages": stage_rows, "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": "vectors/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "ms_per_step": dt * 1e3 / e2e_steps,
                    "host_ms_per_step": {k: v / e2e_steps for k, v in host_ms.items()}},
            "gpu_launches": launches, "clocks": clocks, "final_loss": final_loss, "alt_score_gemm_mode": alt,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--latents-per-gpu", type=int, default=8)
    ap.add_argument("--passes-fwd", type=int, default=3, choices=[1, 3])
    ap.add_argument("--passes-bwd", type=int, default=1, choices=[1, 3])
    ap.add_argument("--no-alt", action="store_true", help="skip the extra timing of the other score-GEMM mode")
    ap.add_argument("--proto-f16", action="store_true",
                    help="pixel x prototype score GEMM on single fp16 planes of the unit-norm operands (|dS| 1.3e-5 "
                         "rms, codes within 3e-3 rms) instead of the default 3-plane bf16 split (2e-7 / 5e-5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
