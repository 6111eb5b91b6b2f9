#!/usr/bin/env python
"""Benchmarks of the per-pixel hidden-feature clustering path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W                 # headline: ffhq-256 hfc_with_swav step
    python bench.py --workload car-512 | pidray-256-labelmap | kmeans-assign      # BASELINE configs 3 / 4 / 5
    python bench.py --impl reference [--workload ...]             # the same workload on the host CPU (oracle port)

Headline (BASELINE metric, per-pixel feature vectors / s): one "step" = one optimiser step of
SwAVClustering.pretrain over B latents per GPU (weak scaling): 2 latent-perturbed views x 5 patches x 20000
random pixels per latent, D = 5376, C = 512, K = 5000 prototypes, 10 Sinkhorn iterations, LARC+SGD - i.e.
B * 200000 per-pixel feature vectors per GPU per step, synthetic latents, random-init StyleGAN2-256 + head
(seed 42).  `value` is timed on the device with the step's raw inputs resident in HBM (the index bookkeeping
kernels are part of the step); `e2e` goes through the public step API from pinned host buffers (latents, random
draws, sampled-pixel indices) and reads the loss back every step.

Every line carries `roofline` (dominant kernel, live CUDA-event timing), `cpu_baseline` (the oracle port on the
host cores, bounded sample) and `e2e`.  The headline line also carries `eager_gpu` (the same step in eager
PyTorch on the same B200 and the reference's own upfirdn2d / fused_bias_act CUDA kernels built for sm_100a,
beside the gx_ kernels), `alt_score_gemm_mode`, `alt_bwd_passes3`, and - on several GPUs - `dist_parity`.
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

FFHQ = dict(name="ffhq-256", size=256, style_dim=512, n_mlp=8, hlen=5376, nclasses=512, nprototypes=5000, patch=20000,
            npatch=5, niters=10, eps=0.005, temperature=0.01, truncation=0.7, n_layers=6, perturb_std=[1.0] * 6,
            lr=0.01, momentum=0.9, trust=0.01)
# configs/segmentors/hfc_with_swav_car_config.py:51-65 on BASELINE.json's 512^2 generator
CAR = dict(FFHQ, name="car-512", size=512, nprototypes=4000, eps=0.01, temperature=0.01)
METRIC = {"ffhq-256": "per-pixel feature vectors/sec (ffhq-256 SwAV step)",
          "car-512": "per-pixel feature vectors/sec (car-512 SwAV step)",
          "pidray-256-labelmap": "label-map pixels/sec (pidray-256 predict_swav_codes + argmax)",
          "kmeans-assign": "per-pixel k-means assignments/sec (ffhq-256 hfc_kmeans, 5 layers)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """SM clocks and throttle reasons sampled beside the timed region: NVML queries from a thread of this process
    every 20 ms (pynvml), else `nvidia-smi -lms 100` as a child process.  (The child process perturbs short timed
    regions - its polls stall CUDA calls for milliseconds - so NVML is preferred.)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.t0, self.t1 = index, None, None, None
        self.thread, self.stop, self.samples = None, False, []

    def _nvml_loop(self, nv, h):
        bits = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
                 else nv.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0))),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0))),
                ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap",
                                         getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0)))]
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.samples.append((time.time(), float(sm), float(mx), [n for n, b in bits if b and (r & b)]))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is numeric
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if len(ids) > self.index else self.index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
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
        if self.thread is not None:
            self.stop = True
            self.thread.join(timeout=1.0)
            sel = [s_ for s_ in self.samples if self.t0 is None or self.t0 - 0.05 <= s_[0] <= self.t1 + 0.05]
            sm = [s_[1] for s_ in sel]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(s_[2] for s_ in sel) if sel else None,
                    "reasons": sorted({n for s_ in sel for n in s_[3]}), "samples": len(sel), "source": "nvml"}
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


def dist_env():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def stage_table(log, steps, total_ms, issued=None):
    """per-kernel totals from the CUDA events recorded around every launch of the timed region.  A tensor-core stage
    is reported against the roofline that bounds it: its algorithmic FLOPs x MMA passes at the sustained bf16 peak,
    or - when that takes less time than moving its operands and output once at the HBM peak (short-K GEMMs: the
    per-resolution projections, the k-means scores) - its algorithmic bytes."""
    issued = issued or {}
    stages = {}
    for ev in log:
        name, a, c, work = ev[:4]
        nbytes = ev[4] if len(ev) > 4 else 0.0
        s = stages.setdefault(name, [0.0, 0.0, 0, 0.0])
        s[0] += a.elapsed_time(c)
        s[1] += work
        s[2] += 1
        s[3] += nbytes
    pk = peaks()
    rows = []
    for name, (tms, work, cnt, nbytes) in sorted(stages.items(), key=lambda kv: -kv[1][0]):
        is_tensor = name.split("_")[0].split("@")[0] in {"gemm", "modconv", "segmentor"}
        p = issued.get(name.split("@")[0], 1) if is_tensor else 1
        if is_tensor and nbytes > 0 and nbytes / (pk["hbm"] * 1e9) > p * work / (pk["tf_sustained"] * 1e12):
            # HBM-bound contraction
            achieved = nbytes / (tms * 1e-3) / 1e9 if tms > 0 else 0.0
            rows.append({"kernel": name, "bound": "hbm", "launches": cnt, "ms_per_step": tms / steps, "share": tms / total_ms,
                         "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
                         "note": "tensor-core contraction bound by its operand + output bytes (short K)",
                         "tflops": work / (tms * 1e-3) / 1e12 if tms > 0 else 0.0})
            stages[name] = [tms, nbytes, cnt, nbytes]
            continue
        peak = pk["tf_sustained"] if is_tensor else pk["hbm"]
        achieved = (work / (tms * 1e-3)) / (1e12 if is_tensor else 1e9) if tms > 0 else 0.0
        row = {"kernel": name, "bound": "tensor" if is_tensor else "hbm", "launches": cnt,
               "ms_per_step": tms / steps, "share": tms / total_ms, "achieved": achieved, "peak": peak,
               "unit": "TFLOP/s" if is_tensor else "GB/s", "frac": achieved / peak}
        if is_tensor:
            row["mma_passes"] = p
            row["frac_issued"] = p * achieved / peak
            if p > 1:   # precision-adjusted ceiling of the algorithmic fraction
                row["ceiling"] = f"{p}-pass split-bf16 (fp32-grade products) => algorithmic cap {1.0 / p:.0%} of the bf16 peak"
        rows.append(row)
    return rows, stages, pk


def roofline_of(rows, stages, pk):
    if not rows:
        return None
    top = rows[0]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(top["kernel"], {}).get("dram_bytes_per_launch")
    tms, work, cnt = stages[top["kernel"]][:3]
    return {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
            "unit": top["unit"], "frac": top["frac"], "traffic": traffic, "algorithmic_per_launch": work / cnt,
            "launches": cnt, "ms_per_launch": tms / cnt, "peak_source": pk["src"],
            "note": "algorithmic FLOPs/bytes per launch / mean CUDA-event launch time; a 3-pass split-bf16 GEMM "
                    "issues 3x its algorithmic FLOPs on the tensor pipe"}


# ----------------------------------------------------------------------------------------
# CPU legs (reference arm / cpu_baseline): the oracle port on the host cores
# ----------------------------------------------------------------------------------------

def make_cpu_state(cfg, device="cpu"):
    from oracle import ganecdotes_oracle as O
    torch.manual_seed(42)
    sd = O.init_generator_state(cfg["size"], cfg["style_dim"], cfg["n_mlp"], 42, randomize_small=False)
    mean_latent = O.style_mlp(sd, torch.randn(4096, cfg["style_dim"])).mean(0, keepdim=True)
    proj = torch.nn.Linear(cfg["hlen"], cfg["nclasses"], bias=False)
    proto = torch.nn.Linear(cfg["nclasses"], cfg["nprototypes"])
    st = [sd, mean_latent, proj.weight.data.clone(), proto.weight.data.clone(), proto.bias.data.clone()]
    if device != "cpu":
        st = [{k: v.to(device) for k, v in sd.items()}] + [t.to(device) for t in st[1:]]
    return st


def oracle_swav_step(state, cfg, seed, npatch=None, device="cpu"):
    """The reference's step (ref swav_clustering.py:320-460) as restated by the oracle: 1 latent, 2 views, `npatch`
    patches of 20000 px through synthesis, upsample+concat, rotate/flip, sampling, projection, prototypes,
    2 x Sinkhorn per patch, loss, backward, LARC+SGD.  Synthesis and rotation are done once per view and shared
    by the patches, as in the reference.  Returns the number of per-pixel feature vectors processed."""
    from oracle import ganecdotes_oracle as O
    sd, mean_latent, wp, wk, bk = state
    npatch = npatch or cfg["npatch"]
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    z = torch.randn(1, cfg["style_dim"], generator=g).to(device)
    w = O.style_mlp(sd, z)
    hw = cfg["size"] ** 2
    perms = [torch.randperm(hw, generator=g).to(device) for _ in range(npatch)]
    rows = {}
    for v in "st":
        hf, _ = O.view_features(sd, w, mean_latent, cfg["truncation"], int(rs.randint(cfg["n_layers"])),
                                torch.randn(2 * cfg["n_layers"], cfg["style_dim"], generator=g).to(device),
                                cfg["n_layers"], cfg["perturb_std"], cfg["hlen"])
        hf = O.rotate_flip(hf, float(rs.uniform(-10, 10)), bool(rs.rand() < 0.5))
        rows[v] = [O.sample_rows(hf, perms[p], cfg["patch"]) for p in range(npatch)]
        del hf
    out = O.swav_step(rows["s"], rows["t"], wp, wk, bk, cfg["niters"], cfg["eps"], cfg["temperature"], None,
                      cfg["lr"], cfg["momentum"], cfg["trust"])
    state[2], state[3], state[4] = out["params"]
    return 2 * npatch * cfg["patch"]


def time_oracle_swav(cfg, steps, warmup, device="cpu"):
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    state = make_cpu_state(cfg, device)
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    for i in range(warmup):
        oracle_swav_step(state, cfg, 1000 + i, device=device)
    sync()
    t0 = time.perf_counter()
    vecs = 0
    for i in range(steps):
        vecs += oracle_swav_step(state, cfg, 2000 + i, device=device)
    sync()
    dt = time.perf_counter() - t0
    return vecs / dt, dt / max(steps, 1), torch.get_num_threads()


def swav_workload_name(cfg, b):
    return (f"{cfg['name']} hfc_with_swav pretrain step: {b} latents/GPU x 2 views x {cfg['npatch']} patches x "
            f"{cfg['patch']} px, D={cfg['hlen']} C={cfg['nclasses']} K={cfg['nprototypes']}, eps={cfg['eps']} "
            f"T={cfg['temperature']}, Sinkhorn {cfg['niters']} it, fwd+bwd+LARC/SGD")


def swav_config(cfg, b, world, exchange=None):
    """the `config` block of a SwAV-step line - the same for this repo's arm and for the reference arm"""
    return {"workload": swav_workload_name(cfg, b), "latents_per_gpu": b, "global_latents": b * world,
            "vectors_per_step": world * b * 2 * cfg["npatch"] * cfg["patch"],
            "l2": f"inputs_larger_than_l2 ({4e-9 * b * cfg['patch'] * cfg['nprototypes']:.1f} GB score matrices)",
            "generator": f"StyleGAN2-{cfg['size']} random init seed 42", "sinkhorn": "joint-batch (distributed)",
            "exchange": exchange}


CPU_SWAV_SAMPLE = ("per step: ONE latent x 2 views x all 5 patches x 20000 px (200000 vectors) through the full step "
                   "of the reference as restated by the oracle, on the host CPU; the GPU arm runs the same step on "
                   "{b} latents per GPU")


def run_reference(args):
    """`--impl reference`: the reference's CPU path (oracle port; the reference is a script repository, not
    installable) with all host threads, rank 0 only."""
    world, rank, _ = dist_env()
    if rank != 0:
        return
    wl = args.workload
    if wl in ("ffhq-256", "car-512"):
        cfg = FFHQ if wl == "ffhq-256" else CAR
        b = args.latents_per_gpu or (8 if wl == "ffhq-256" else 2)
        warm = 1 if args.warmup > 0 else 0
        value, sec, cores = time_oracle_swav(cfg, args.steps, warm)
        sample = CPU_SWAV_SAMPLE.format(b=b)
        line = {"impl": "reference", "metric": METRIC[wl], "value": value, "unit": "vectors/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(swav_config(cfg, b, max(1, args.gpus)), sample=sample),
                "cpu_baseline": {"value": value, "unit": "vectors/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "vectors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
    elif wl == "pidray-256-labelmap":
        v, sec, cores, sample = cpu_labelmap(args.steps, 1 if args.warmup > 0 else 0)
        line = {"impl": "reference", "metric": METRIC[wl], "value": v, "unit": "pixels/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": labelmap_workload_name(args.images_per_gpu), "sample": sample},
                "cpu_baseline": {"value": v, "unit": "pixels/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
    else:
        v, sec, cores, sample = cpu_kmeans(args.steps, 1 if args.warmup > 0 else 0)
        line = {"impl": "reference", "metric": METRIC[wl], "value": v, "unit": "pixels/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": kmeans_workload_name(args.images_per_gpu), "sample": sample},
                "cpu_baseline": {"value": v, "unit": "pixels/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------
# GPU-side baselines on the same B200 (headline workload only)
# ----------------------------------------------------------------------------------------

def _cuda_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def eager_gpu_block(cfg, ours_vectors_per_s_per_gpu):
    """(1) the oracle's step - eager PyTorch: cuDNN convs, torch.matmul (fp32, TF32 off), ATen Sinkhorn / loss /
    autograd - on the same B200, one latent per step like the reference; (2) the reference's own CUDA kernels
    (lib/gan/optim/upfirdn2d_kernel.cu, fused_bias_act_kernel.cu, built unmodified for sm_100a into oracle/_ref/)
    beside gx_upfirdn2d / gx_fused_bias_act on a StyleGAN2 256^2 layer shape.  Baselines, not the product."""
    from ganecdotes_b200 import _lib as L
    out = {}
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        v, sec, _ = time_oracle_swav(cfg, 2, 1, device="cuda")
        out["eager_step"] = {"value": v, "unit": "vectors/s", "sec_per_step": sec,
                             "what": "oracle step (eager PyTorch fp32: cuDNN / cuBLAS / ATen) on cuda:0, 1 latent x 2 "
                                     "views x 5 patches x 20000 px per step",
                             "ours_over_eager": ours_vectors_per_s_per_gpu / v}
    except Exception as e:  # baseline only: never fails the bench
        out["eager_step"] = {"error": repr(e)[:200]}
    torch.cuda.empty_cache()
    try:
        from oracle import build_ref
        up, fb = build_ref.load("ref_upfirdn2d"), build_ref.load("ref_fused_bias_act")
        if up is None or fb is None:
            out["reference_cuda_ops"] = {"unavailable": "oracle/_ref not built (python oracle/build_ref.py)"}
            return out
        # the blur after the 128 -> 256 up-conv of StyleGAN2-256 (ref model.py:166-182): [B*C, 257, 257, 1], pad (1,1)
        b, c, h = 16, 128, 257
        x = torch.randn(b * c, h, h, 1, device="cuda")
        k1 = torch.tensor([1., 3., 3., 1.], device="cuda")
        k = (k1[None, :] * k1[:, None]) / k1.sum() ** 2 * 4
        y_ref = up.upfirdn2d(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
        y_gx = L.upfirdn2d_raw(x, k, 1, 1, 1, 1, 1, 1, 1, 1)
        err_up = (y_ref - y_gx).abs().max().item()
        nbytes = (x.numel() + y_ref.numel()) * 4
        t_ref = _cuda_ms(lambda: up.upfirdn2d(x, k, 1, 1, 1, 1, 1, 1, 1, 1))
        t_gx = _cuda_ms(lambda: L.upfirdn2d_raw(x, k, 1, 1, 1, 1, 1, 1, 1, 1))
        xa = torch.randn(b, c, 256, 256, device="cuda")
        bias = torch.randn(c, device="cuda")
        empty = torch.empty(0, device="cuda")
        z_ref = fb.fused_bias_act(xa, bias, empty, 3, 0, 0.2, 2 ** 0.5)
        z_gx = L.fused_bias_act_raw(xa, bias, None, 3, 0, 0.2, 2 ** 0.5)
        err_fb = (z_ref - z_gx).abs().max().item()
        t_ref2 = _cuda_ms(lambda: fb.fused_bias_act(xa, bias, empty, 3, 0, 0.2, 2 ** 0.5))
        t_gx2 = _cuda_ms(lambda: L.fused_bias_act_raw(xa, bias, None, 3, 0, 0.2, 2 ** 0.5))
        nb2 = 2 * xa.numel() * 4
        hbm = peaks()["hbm"]
        out["reference_cuda_ops"] = {
            "upfirdn2d_blur_256": {"shape": [b * c, h, h, 1], "reference_ms": t_ref, "gx_ms": t_gx,
                                   "reference_gbs": nbytes / t_ref / 1e6, "gx_gbs": nbytes / t_gx / 1e6,
                                   "gx_frac_hbm": nbytes / t_gx / 1e6 / hbm, "max_abs_diff": err_up},
            "fused_bias_act_256": {"shape": [b, c, 256, 256], "reference_ms": t_ref2, "gx_ms": t_gx2,
                                   "reference_gbs": nb2 / t_ref2 / 1e6, "gx_gbs": nb2 / t_gx2 / 1e6,
                                   "gx_frac_hbm": nb2 / t_gx2 / 1e6 / hbm, "max_abs_diff": err_fb},
            "what": "the reference's lib/gan/optim kernels compiled unmodified for sm_100a (oracle/_ref) vs the gx_ ops"}
    except Exception as e:
        out["reference_cuda_ops"] = {"error": repr(e)[:200]}
    return out


# ----------------------------------------------------------------------------------------
# SwAV training step (ffhq-256 headline, car-512)
# ----------------------------------------------------------------------------------------

def run_swav(args, cfg):
    import torch.distributed as dist
    from ganecdotes_b200 import _lib as L
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.hfc_with_swav.swav_clustering import SwAVClustering
    from ganecdotes_b200.stylegan2.model import Generator

    world, rank, local_rank = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = E.DistGroup(dist.group.WORLD, rank, world)
    L.load()
    headline = cfg["name"] == "ffhq-256"
    b = args.latents_per_gpu or (8 if headline else 2)

    # ---- model + head, identical on every rank (seed 42 = the reference's seed, lib/util/util.py:21-24)
    torch.manual_seed(42)
    np.random.seed(42)
    gen = Generator(cfg["size"], cfg["style_dim"], cfg["n_mlp"]).to(dev)
    gen.passes = args.passes_fwd
    with torch.no_grad():
        mean_latent = gen.style(torch.randn(4096, cfg["style_dim"]).to(dev)).mean(0, keepdim=True)
    proj = torch.nn.Linear(cfg["hlen"], cfg["nclasses"], bias=False).to(dev)
    proto = torch.nn.Linear(cfg["nclasses"], cfg["nprototypes"]).to(dev)
    w_init = [proj.weight.data.clone(), proto.weight.data.clone(), proto.bias.data.clone()]

    def new_head(passes_bwd=None):
        return E.SwavHead(w_init[0].clone(), w_init[1].clone(), w_init[2].clone(), cfg["lr"], cfg["momentum"],
                          cfg["trust"], args.passes_fwd, passes_bwd or args.passes_bwd, proto_f16=args.proto_f16)

    head = new_head()
    scfg = E.StepConfig(hlen=cfg["hlen"], patch_size=cfg["patch"], num_patches=cfg["npatch"], niters=cfg["niters"],
                        eps=cfg["eps"], temperature=cfg["temperature"], truncation=cfg["truncation"],
                        perturb_std=cfg["perturb_std"])
    ws = L.SinkhornWorkspace(cfg["nprototypes"], dev)
    if group is not None and os.environ.get("GX_SINKHORN_EXCHANGE", "ll") != "nccl":
        group.ensure_ll(cfg["nprototypes"], dev)

    def draw(seed, nb=b):
        g = torch.Generator().manual_seed(seed)
        rs = np.random.RandomState(seed)
        hw = cfg["size"] ** 2

        def view():
            return E.ViewDraws(layer_no=[int(rs.randint(cfg["n_layers"])) for _ in range(nb)],
                               pert_z=torch.randn(nb, 2 * cfg["n_layers"], cfg["style_dim"], generator=g).pin_memory(),
                               angle=[float(rs.uniform(-10, 10)) for _ in range(nb)],
                               flip=[bool(rs.rand() < 0.5) for _ in range(nb)])
        return E.StepDraws(z=torch.randn(nb, cfg["style_dim"], generator=g).pin_memory(), view_s=view(), view_t=view(),
                           perms=[[torch.randperm(hw, generator=g) for _ in range(nb)] for _ in range(cfg["npatch"])])

    nsteps = args.warmup + args.steps
    draws = [draw(10_000 * (rank + 1) + i) for i in range(nsteps)]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return x

    def timed_steps(hd, first, count):
        """`count` steps on fresh inputs (index bookkeeping included), device-timed; max over ranks"""
        inp = [E.prepare_step_inputs(gen, draws[first + i], scfg, dev) for i in range(count)]
        sync()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        loss = None
        for i in range(count):
            loss = E.swav_train_step_device(gen, hd, mean_latent, inp[i], scfg, group, ws)
        a1.record()
        sync()
        return max_over_ranks(a0.elapsed_time(a1)), loss

    # ---------------------------------------------------------------- device-timed region
    # raw inputs (latents, draws, sampled-pixel source indices) resident in HBM; everything else is in the step
    inputs = [E.prepare_step_inputs(gen, d, scfg, dev) for d in draws]
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
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = L.launch_count
    log = L.event_log
    L.event_log = None
    del inputs
    vec_per_step = world * b * 2 * cfg["npatch"] * cfg["patch"]
    value = vec_per_step * args.steps / (ms * 1e-3)
    final_loss = float(loss)
    if group is not None and group.ll is not None:
        group.ll.check()

    issued = {"gemm_prototype_fwd": 1 if args.proto_f16 else 3, "gemm_projection_fwd": args.passes_fwd,
              "modconv": args.passes_fwd, "modconv_up": args.passes_fwd, "gemm_dzn_bwd": args.passes_bwd,
              "gemm_gproto_bwd": args.passes_bwd, "gemm_gproj_bwd": args.passes_bwd}
    stage_rows, stages, pk = stage_table(log, args.steps, ms, issued)
    roofline = roofline_of(stage_rows, stages, pk)

    # ---------------------------------------------------------------- alternative precisions, device-timed
    alt = alt_bwd = alt_sk = None
    if not args.no_alt:
        alt_steps = max(1, min(args.steps, 5))
        head.proto_f16 = not args.proto_f16          # the other score-GEMM operand mode
        timed_steps(head, 0, min(2, nsteps))
        alt_ms, _ = timed_steps(head, args.warmup, alt_steps)
        alt = {"score_gemm": "fp16x1" if head.proto_f16 else "bf16x3", "steps": alt_steps,
               "ms_per_step": alt_ms / alt_steps, "value": vec_per_step * alt_steps / (alt_ms * 1e-3),
               "unit": "vectors/s"}
        head.proto_f16 = args.proto_f16
        if args.passes_bwd != 3:                      # what fp32-grade gradients cost (3-pass split-bf16 backward)
            h3 = new_head(passes_bwd=3)
            timed_steps(h3, 0, min(2, nsteps))
            b3_ms, _ = timed_steps(h3, args.warmup, alt_steps)
            alt_bwd = {"passes_bwd": 3, "steps": alt_steps, "ms_per_step": b3_ms / alt_steps,
                       "value": vec_per_step * alt_steps / (b3_ms * 1e-3), "unit": "vectors/s",
                       "note": "backward GEMMs on 3-plane split-bf16 operands: gradients within 3e-3 of fp32 "
                               "instead of 2e-2 (default bf16x1)"}
            del h3
        # the Sinkhorn passes 3.. on the fp32 scores instead of the 16-bit cache (the round-1 path)
        cache_was = scfg.sinkhorn_cache16
        scfg.sinkhorn_cache16 = False
        timed_steps(head, 0, min(2, nsteps))
        sk_ms, _ = timed_steps(head, args.warmup, alt_steps)
        scfg.sinkhorn_cache16 = cache_was
        alt_sk = {"sinkhorn_cache16": False, "steps": alt_steps, "ms_per_step": sk_ms / alt_steps,
                  "value": vec_per_step * alt_steps / (sk_ms * 1e-3), "unit": "vectors/s",
                  "note": "every Sinkhorn pass streams the fp32 scores (column scalings to ~1e-5 instead of ~5e-4; "
                          "the codes are computed from the fp32 scores in both modes)"}

    # ---------------------------------------------------------------- end-to-end region
    # Same schedule as SwAVClustering.pretrain: host bookkeeping + pinned H2D of step i+1 go on a side stream while
    # step i computes; the loss of step i is read back once step i+1 has been launched.  `warmup` untimed steps
    # (allocator pools of the side stream, pinned staging buffers), then K timed steps in steady state: every timed
    # step contains one staging, one optimiser step and one device->host loss read.
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
    dt = max_over_ranks(time.perf_counter() - t0)
    e2e_value = vec_per_step * e2e_steps / dt
    del inp

    # ---------------------------------------------------------------- multi-GPU parity, driver-visible
    dist_parity = None
    if world > 1:
        # one latent per rank, sharded over the ranks (distributed Sinkhorn over the NVLink exchange + gradient
        # all-reduce) against the SAME global batch in one process on rank 0 (no group): loss and weight updates
        gdraws = draw(777, nb=world)                   # identical on every rank (same seed)
        h_sh = new_head(passes_bwd=3)
        l_sh = E.swav_train_step(gen, h_sh, mean_latent, SwAVClustering.shard_draws(gdraws, rank, world), scfg, group,
                                 ws)
        torch.cuda.synchronize()
        if rank == 0:
            h_one = new_head(passes_bwd=3)
            l_one = E.swav_train_step(gen, h_one, mean_latent, gdraws, scfg, None, ws)
            torch.cuda.synchronize()
            upd = []
            for a, r_, w0 in zip((h_sh.w_proj, h_sh.w_proto, h_sh.b_proto), (h_one.w_proj, h_one.w_proto, h_one.b_proto),
                                 w_init):
                w0n = torch.nn.functional.normalize(w0, dim=1) if w0.dim() == 2 and w0.shape[0] == cfg["nprototypes"] else w0
                upd.append(((a - r_).norm() / (r_ - w0n).norm().clamp_min(1e-30)).item())
            rel_loss = abs(l_sh.item() - l_one.item()) / abs(l_one.item())
            dist_parity = {"what": f"{world} latents sharded 1 per rank vs the same batch in one process on rank 0",
                           "loss_sharded": l_sh.item(), "loss_single": l_one.item(), "rel_loss_diff": rel_loss,
                           "rel_update_diff": {"w_proj": upd[0], "w_proto": upd[1], "b_proto": upd[2]},
                           "exchange": "ll-nvlink" if group.ll is not None else "nccl",
                           "ok": bool(rel_loss < 1e-4 and max(upd) < 1e-2)}
            del h_one
        del h_sh
        sync()

    cpu_base = eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores = time_oracle_swav(cfg, 1, 0)
        cpu_base = {"value": v, "unit": "vectors/s", "cores": cores, "kind": "port",
                    "sample": CPU_SWAV_SAMPLE.format(b=b) + " - 1 step, no warm-up", "sec_per_sample_step": sec}
    if rank == 0 and world == 1 and headline and not args.no_eager:
        del head
        torch.cuda.empty_cache()
        eager = eager_gpu_block(cfg, value)
    if rank == 0:
        line = {
            "metric": METRIC[cfg["name"]], "value": value, "unit": "vectors/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": f"bf16x{args.passes_fwd}-split fwd" + (" (score GEMM fp16x1 on unit-norm operands)" if args.proto_f16 else "") +
                     f" / bf16x{args.passes_bwd} bwd operands, fp32 accumulate + fp32 everywhere else" +
                     ("" if scfg.sinkhorn_cache16 is False or os.environ.get("GX_SINKHORN_CACHE16", "1") == "0" else
                      " (Sinkhorn iterations 3.. stream an fp16 cache of the row-normalised kernel matrix)"),
            "data": "synthetic",
            "config": swav_config(cfg, b, world, None if world == 1 else ("ll-nvlink" if group.ll is not None else "nccl")),
            "roofline": roofline, "roofline_stages": stage_rows, "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": "vectors/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "steps": e2e_steps, "ms_per_step": dt * 1e3 / e2e_steps,
                    "host_ms_per_step": {k: v / e2e_steps for k, v in host_ms.items()}},
            "gpu_launches": launches, "clocks": clocks, "final_loss": final_loss, "alt_score_gemm_mode": alt,
            "alt_bwd_passes3": alt_bwd, "alt_sinkhorn_fp32_passes": alt_sk, "dist_parity": dist_parity,
            "eager_gpu": eager,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        if group.ll is not None:
            dist.barrier()
            group.ll.close()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------
# BASELINE config 4: pidray-256 label maps (evaluate.py inference: predict_swav_codes + argmax)
# ----------------------------------------------------------------------------------------

def labelmap_workload_name(b):
    return (f"pidray-256 predict_swav_codes: {b} images/GPU per step, BagGAN generator (sum C = 2528) -> per-pixel "
            f"codes (projection 2528 -> 512) -> arg-max int64 label maps 256x256, truncation 0.9")


def cpu_labelmap(steps, warmup):
    from oracle import ganecdotes_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    sd = O.init_generator_state(256, 512, 8, 42, channels=O.baggan_channels(), randomize_small=False)
    mean_latent = O.style_mlp(sd, torch.randn(1024, 512)).mean(0, keepdim=True)
    wp = torch.randn(512, 2528) / 2528 ** 0.5
    n = 0
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        w = O.style_mlp(sd, torch.randn(1, 512))
        O.predict_codes(sd, w, mean_latent, 0.9, wp, 2528)
        n += 65536 if i >= warmup else 0
    dt = time.perf_counter() - t0
    return n / dt, dt / max(steps, 1), torch.get_num_threads(), "per step: 1 image (65536 label pixels) on the host CPU"


def run_labelmap(args):
    import torch.distributed as dist
    from ganecdotes_b200 import _lib as L
    from ganecdotes_b200.baggan import baggan_channels
    from ganecdotes_b200.hfc_with_swav import engine as E
    from ganecdotes_b200.stylegan2.model import Generator
    world, rank, local_rank = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)      # replicas: only the timing barrier / max uses it
    L.load()
    b = args.images_per_gpu
    torch.manual_seed(42)
    gen = Generator(256, 512, 8, channels=baggan_channels()).to(dev)
    hlen, trunc = 2528, 0.9
    wp = (torch.randn(512, hlen) / hlen ** 0.5).to(dev)
    with torch.no_grad():
        mean_latent = gen.style(torch.randn(1024, 512).to(dev)).mean(0, keepdim=True)
    zs = [torch.randn(b, 512, generator=torch.Generator().manual_seed(100 * rank + i)).pin_memory()
          for i in range(args.warmup + args.steps)]
    ws_dev = [gen.style(z.to(dev)) for z in zs]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step(w):
        return E.predict_codes(gen, wp, w, mean_latent, trunc, hlen)[1]

    # ---- eager pass: CUDA events around every launch -> the per-stage table
    for i in range(args.warmup):
        step(ws_dev[i])
    sync()
    L.event_log = []
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for i in range(args.steps):
        step(ws_dev[args.warmup + i])
    a1.record()
    sync()
    eager_ms = a0.elapsed_time(a1)
    log = L.event_log
    L.event_log = None
    rows, stages, pk = stage_table(log, args.steps, eager_ms, {"gemm_projection_fwd": 3, "modconv": 3, "modconv_up": 3})
    # ---- timed region: the same batch as one CUDA-graph replay per step (engine.PredictGraph; bit-identical results)
    graph = None if args.no_graph else E.PredictGraph(gen, wp, mean_latent, trunc, hlen, b)
    run = (lambda w: graph(w)[1]) if graph is not None else step
    for i in range(args.warmup):
        run(ws_dev[i])
    sync()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sampler.mark_begin()
    L.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        labels = run(ws_dev[args.warmup + i])
    e1.record()
    sync()
    sampler.mark_end()
    clocks = sampler.summary() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = L.launch_count
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    px = world * b * 65536
    value = px * args.steps / (ms * 1e-3)
    # e2e: pinned host z in, int64 label maps back on the host, every step
    sync()
    host = [torch.empty((b, 256, 256), dtype=torch.int64).pin_memory() for _ in range(2)]   # pinned, double-buffered
    t0 = time.perf_counter()
    for i in range(args.steps):
        with torch.no_grad():
            w = gen.style(zs[args.warmup + i].to(dev, non_blocking=True))
        host[i % 2].copy_(run(w), non_blocking=True)
    sync()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = t.item()
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores, sample = cpu_labelmap(3, 1)
        cpu_base = {"value": v, "unit": "pixels/s", "cores": cores, "kind": "port", "sample": sample + ", 3 steps"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC["pidray-256-labelmap"], "value": value, "unit": "pixels/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3-split (fp32-grade) convs + projection, fp32 elsewhere",
            "data": "synthetic",
            "config": {"workload": labelmap_workload_name(b), "images_per_gpu": b,
                       "l2": "inputs_larger_than_l2 (codes 512 x 65536 x 4 B = 134 MB per image)",
                       "parallelism": "replicas (no collective)",
                       "launch": "one CUDA-graph replay per step (engine.PredictGraph)" if graph is not None else "stream launches"},
            "eager_ms_per_step": eager_ms / args.steps,
            "roofline": roofline_of(rows, stages, pk), "roofline_stages": rows, "cpu_baseline": cpu_base,
            "e2e": {"value": px * args.steps / dt, "unit": "pixels/s", "h2d_bytes_per_step": b * 512 * 4,
                    "d2h_bytes_per_step": b * 65536 * 8, "ms_per_step": dt * 1e3 / args.steps},
            "gpu_launches": launches, "clocks": clocks, "label_checksum": int(labels.sum().item())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------
# BASELINE config 5: per-pixel k-means assignment on GAN features (hfc_kmeans predict)
# ----------------------------------------------------------------------------------------
KM_CLUSTERS = [4, 8, 16, 32, 64]       # clusters_per_layer of the five layers (8^2 ... 128^2)


def kmeans_workload_name(b):
    return (f"ffhq-256 hfc_kmeans predict: {b} images/GPU per step, layers 8^2..128^2 (C = 1024,1024,1024,1024,512; "
            f"K = {KM_CLUSTERS}), nearest-centre assignment + one-hot NEAREST maps at 256^2; features resident")


def cpu_kmeans(steps, warmup):
    from oracle import ganecdotes_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(42)
    sd = O.init_generator_state(256, 512, 8, 42, randomize_small=False)
    _, feats, _ = O.generator_forward(sd, torch.randn(1, 512), 1.0, None, False)
    layers = O.regroup_features(feats)[1:1 + len(KM_CLUSTERS)]        # maps (2n+1, 2n+2) concatenated per layer
    cen = [torch.randn(k, f.shape[1]) for f, k in zip(layers, KM_CLUSTERS)]
    px = sum(f.shape[2] ** 2 for f in layers)
    t0 = None
    for i in range(warmup + steps):
        if i == warmup:
            t0 = time.perf_counter()
        O.kmeans_layer_maps(layers, cen, 256)
    dt = time.perf_counter() - t0
    return px * steps / dt, dt / max(steps, 1), torch.get_num_threads(), "per step: 1 image, 5 layers, on the host CPU"


def run_kmeans(args):
    import torch.distributed as dist
    from ganecdotes_b200 import _lib as L
    from ganecdotes_b200.hfc_kmeans.hfc_kmeans_clustering import FlatKMeansAssign
    from ganecdotes_b200.stylegan2.model import Generator
    world, rank, local_rank = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()
    b = args.images_per_gpu
    torch.manual_seed(42)
    gen = Generator(256, 512, 8).to(dev)
    lat = torch.randn(b, gen.n_latent, 512, generator=torch.Generator().manual_seed(rank)).to(dev)
    _, feats = gen.synthesize(lat, None, need_image=False)
    feats_nchw = [f.permute(0, 3, 1, 2) for f in feats]
    cen = [torch.randn(k, feats[2 * n + 1].shape[3] * 2, generator=torch.Generator().manual_seed(n))
           for n, k in enumerate(KM_CLUSTERS)]
    km = FlatKMeansAssign(cen, 256, dev)
    px_img = sum(feats[2 * n + 1].shape[1] ** 2 for n in range(len(KM_CLUSTERS)))
    alg_bytes = sum(feats[2 * n + 1].numel() * 8 for n in range(len(KM_CLUSTERS)))      # both maps read once, fp32

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def assign_only():
        for n in range(len(KM_CLUSTERS)):
            f1, f2 = feats[2 * n + 1], feats[2 * n + 2]
            with L.timed("kmeans_assign_layer", (f1.numel() + f2.numel()) * 4.0):    # features read once, fp32
                L.kmeans_assign(f1.reshape(-1, f1.shape[3]), km.centers[n], f2.reshape(-1, f2.shape[3]))

    for _ in range(args.warmup):
        km.predict(feats_nchw)
    sync()
    # the step (ten launches, ~0.7 ms) as one CUDA-graph replay, like the label-map workload: as stream launches from
    # Python it is bound by the host (0.8 ... 3 ms per step from run to run on the same code)
    graph, per_replay = None, 0
    if not args.no_graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                km.predict(feats_nchw)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            c0 = L.launch_count
            with torch.cuda.graph(graph):
                g_out = km.predict(feats_nchw)
            per_replay = L.launch_count - c0
            for _ in range(2):
                graph.replay()
            sync()
        except Exception as exc:      # capture refused: stream launches
            print(f"kmeans-assign: CUDA-graph capture failed ({exc}); timing stream launches", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sampler.mark_begin()
    L.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
            L._count(per_replay)
            maps, labs = g_out
        else:
            maps, labs = km.predict(feats_nchw)
    e1.record()
    sync()
    sampler.mark_end()
    clocks = sampler.summary() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = L.launch_count
    L.event_log = []
    for _ in range(args.steps):
        assign_only()
    torch.cuda.synchronize()
    # the layer-level entry (algorithmic bytes: both feature maps read once) is the roofline line; its inner kernels
    # (split_planes -> gemm_kmeans_scores -> kmeans_argmin on the tensor-core route) are listed after it
    log = L.event_log
    L.event_log = None
    rows, stages, pk = stage_table([e for e in log if e[0] == "kmeans_assign_layer"], args.steps, ms, {})
    rows += stage_table([e for e in log if e[0] != "kmeans_assign_layer"], args.steps, ms, {"gemm_kmeans_scores": 3})[0]
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    px = world * b * px_img
    # e2e: the reference-facing call is predict_hfc_vectors(input_latent) (ref baseline/hfc_kmeans/segmentor.py:168-226):
    # W+ latents from pinned host memory -> generator -> per-layer assignment -> one-hot maps; label maps back to the host
    host_lat = lat.cpu().pin_memory()
    e2e_steps = max(1, min(args.steps, 10))
    host_labs = None
    step_ms = []

    def e2e_step():
        nonlocal host_labs
        _, f_e = gen.synthesize(host_lat.to(dev, non_blocking=True), None, need_image=False)
        _, labs_ = km.predict([f.permute(0, 3, 1, 2) for f in f_e])
        if host_labs is None:
            host_labs = [torch.empty(l.shape, dtype=l.dtype).pin_memory() for l in labs_]
        for h_, l_ in zip(host_labs, labs_):
            h_.copy_(l_, non_blocking=True)
        return labs_

    for _ in range(2):                       # warm-up of this path (pinned buffers, allocator)
        e2e_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        labs = e2e_step()
        torch.cuda.synchronize()             # the label maps of this step are on the host
        step_ms.append((time.perf_counter() - ts) * 1e3)
    sync()
    dt = time.perf_counter() - t0
    # the same with the FEATURES coming from pinned host memory (clusterer.predict on host arrays, ref
    # hfc_kmeans_clustering.py:169-208): bound by the 0.9 GB host-to-device copy per step
    host_feats = [feats[i].cpu().pin_memory() for i in range(1, 2 * len(KM_CLUSTERS) + 1)]
    sync()
    t1 = time.perf_counter()
    for _ in range(e2e_steps):
        dfe = [feats_nchw[0]] + [h.to(dev, non_blocking=True).permute(0, 3, 1, 2) for h in host_feats] + feats_nchw[11:]
        _, labs2 = km.predict(dfe)
        host = [l.cpu() for l in labs2]
    sync()
    dt_feat = time.perf_counter() - t1
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, cores, sample = cpu_kmeans(3, 1)
        cpu_base = {"value": v, "unit": "pixels/s", "cores": cores, "kind": "port", "sample": sample + ", 3 steps"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC["kmeans-assign"], "value": px * args.steps / (ms * 1e-3), "unit": "pixels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": kmeans_workload_name(b), "images_per_gpu": b,
                       "l2": f"inputs_larger_than_l2 ({alg_bytes / 1e6:.0f} MB of features per step)" if alg_bytes > 126e6
                       else f"flush: none; {alg_bytes / 1e6:.0f} MB of features per step (< L2: raise --images-per-gpu)",
                       "parallelism": "replicas (no collective)",
                       "launch": "one CUDA-graph replay per step" if graph is not None else "stream launches"},
            "roofline": roofline_of(rows, stages, pk), "roofline_stages": rows, "cpu_baseline": cpu_base,
            "e2e": {"value": px * e2e_steps / dt, "unit": "pixels/s", "h2d_bytes_per_step": host_lat.numel() * 4,
                    "d2h_bytes_per_step": b * px_img * 4, "ms_per_step": dt * 1e3 / e2e_steps,
                    "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(max(step_ms)),
                    "path": "host W+ latents -> generator -> assignment + maps -> host label maps (predict_hfc_vectors);"
                            " ~300 stream launches issued from Python per step: bound by the host"},
            "e2e_features_from_host": {"value": px * e2e_steps / dt_feat, "unit": "pixels/s",
                                       "h2d_bytes_per_step": sum(h.numel() * 4 for h in host_feats),
                                       "d2h_bytes_per_step": b * px_img * 4, "ms_per_step": dt_feat * 1e3 / e2e_steps},
            "gpu_launches": launches, "clocks": clocks, "label_checksum": int(sum(int(l.sum()) for l in labs))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ffhq-256", choices=list(METRIC))
    ap.add_argument("--latents-per-gpu", type=int, default=0, help="SwAV workloads (default 8 for ffhq-256, 2 for car-512)")
    ap.add_argument("--images-per-gpu", type=int, default=16, help="label-map / k-means workloads")
    ap.add_argument("--passes-fwd", type=int, default=3, choices=[1, 3])
    ap.add_argument("--passes-bwd", type=int, default=1, choices=[1, 3])
    ap.add_argument("--no-alt", action="store_true", help="skip the extra timings of the other precision modes")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch / reference-kernel GPU baselines")
    ap.add_argument("--proto-f16", action="store_true",
                    help="pixel x prototype score GEMM on single fp16 planes of the unit-norm operands (|dS| 1.3e-5 "
                         "rms, codes within 3e-3 rms) instead of the default 3-plane bf16 split (2e-7 / 5e-5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="label-map workload: stream launches instead of a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "ffhq-256":
        run_swav(args, FFHQ)
    elif args.workload == "car-512":
        run_swav(args, CAR)
    elif args.workload == "pidray-256-labelmap":
        run_labelmap(args)
    else:
        run_kmeans(args)


if __name__ == "__main__":
    main()
