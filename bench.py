#!/usr/bin/env python
"""Benchmark of the blind-BWE sampler hot path (BASELINE.json metric:
"blind-BWE sampler steps/s at 1/2/4/8 B200; operator % of HBM roofline").

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # the reference algorithm on host cores
    python bench.py --mode operator                        # config-4 operator sweep (table, not the contract line)

Workload (config.workload): BASELINE configs[1] per-GPU slice -- 8 chains of
T=184184 samples at 22.05 kHz per GPU (64 chains over 8 GPUs), random-init
CQTDiff+ (44.5 M parameters, torch.manual_seed(0)), 35-step EDM schedule,
2nd-order sampler, filter fit max_iter=100, NFFT=4096, K=5 breakpoints,
synthetic piano-like audio low-passed at 1 kHz / -20 dB/oct.  A "step" is one
iteration of the sampling loop (testing/blind_bwe_sampler.py:685, incl. the
Heun correction): 2 x {denoiser fwd+bwd through both CQTs, hpf, filter fit,
reconstruction guidance}.  value = chains * steps / seconds, whole job.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR, AUDIO_LEN, NFFT, CHAINS_PER_GPU = 22050, 184184, 4096, 8
# The CPU arm cannot afford full-length chains (one chain-step of T=184184 costs
# ~150 s on 8 host threads, dominated by the PyTorch conv U-Net): it runs the
# SAME sampler on a quarter-length segment and scales the time by T/T' (the
# cost of the conv net and of the block transforms is linear in T).
CPU_AUDIO_LEN = AUDIO_LEN // 4


def piano_like(B, T, sr, seed):
    """SURVEY 8(d) synthetic input: decaying harmonic notes, std 0.063 (vectorised)."""
    g = torch.Generator().manual_seed(seed)
    n = torch.arange(T, dtype=torch.float32)
    out = torch.zeros(B, T)
    for b in range(B):
        k = torch.randint(20, 76, (24,), generator=g)
        f0 = 27.5 * 2.0 ** (k.float() / 12)
        n0 = torch.randint(0, T, (24,), generator=g).float()
        for p in range(24):
            tt = ((n - n0[p]).clamp(min=0)) / sr
            gate = (n >= n0[p]).float()
            for h in range(1, 13):
                if h * float(f0[p]) >= sr / 2:
                    break
                out[b] += gate * h ** -1.2 * torch.exp(-tt * h / 1.5) * torch.sin(2 * math.pi * h * float(f0[p]) * tt)
    return out / out.std(dim=1, keepdim=True) * 0.063


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                    "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in o.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
def build_world(device, chains, seed, max_iter=100):
    from babe_b200 import blind_bwe_utils as bu, denoiser, edm, sampler
    args = sampler.make_args(sample_rate=SR, audio_len=AUDIO_LEN, NFFT=NFFT, max_iter=max_iter)
    torch.manual_seed(0)
    net = denoiser.CQTDiffPlus(args, device).to(device)
    for p in net.parameters():
        p.requires_grad_(False)
    x = piano_like(chains, AUDIO_LEN, SR, 1234 + seed).to(device)
    f = torch.fft.rfftfreq(NFFT, d=1 / SR).to(device)
    y = bu.apply_filter(x, bu.design_filter([1000.0], [-20.0], f), NFFT)      # blind_bwe.yaml:135-137
    smp = sampler.BlindSamplerFused(net, edm.EDM(args), args, rid=False)
    return args, net, smp, y


def run_ours(a):
    from babe_b200 import build, distributed as bd, profiling
    build.build()
    rank, local, world = bd.init_from_env()
    if world != a.gpus and rank == 0 and a.gpus != 1:
        print(f"warning: --gpus {a.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    torch.backends.cudnn.benchmark = not a.no_autotune
    from babe_b200 import net_ops
    net_ops.AUTOTUNE_CONV = not a.no_autotune
    chains = a.chains
    args, net, smp, y = build_world(device, chains, seed=rank)
    K, W = a.steps, a.warmup
    assert W + K <= args.tester.T, "steps + warmup must fit the 35-step schedule"

    def timed_run(mode):
        """mode 'device': inputs resident, device noise, no logging.
        mode 'e2e': reference-compatible public call -- y and every step's noise come
        from pinned host memory, the step's denoised estimate and filter go back (rid)."""
        ev = {}
        state = {"i": 0}
        if mode == "device":
            smp.device_noise, smp.noise_fn = True, None
            smp.generator = torch.Generator(device=device).manual_seed(100 + rank)
            y_in = y
            h2d = d2h = 0
        else:
            g = torch.Generator().manual_seed(100 + rank)
            pool = [torch.randn(y.shape, generator=g).pin_memory() for _ in range(W + K + 1)]
            y_host = y.cpu().pin_memory()
            out_den = torch.empty((chains, AUDIO_LEN), dtype=torch.float32).pin_memory()
            out_par = torch.empty((2, 5), dtype=torch.float32).pin_memory()
            it = iter(pool)
            smp.device_noise = False
            smp.noise_fn = lambda shape, dev: next(it).to(dev, non_blocking=True)
            h2d = y.numel() * 4            # per step: that step's noise draw
            d2h = y.numel() * 4 + out_par.numel() * 4

        def hook(i, x, p):
            if mode == "e2e":
                out_den.copy_(x, non_blocking=True)     # rid=True logging of the reference tester
                out_par.copy_(p, non_blocking=True)
            if i == W - 1:
                torch.cuda.synchronize()
                bd.barrier()
                torch.cuda.synchronize()
                profiling.reset()
                ev["clk"] = ClockSampler(local).__enter__() if rank == 0 else None
                ev["t0"] = time.perf_counter()
                ev["s"] = torch.cuda.Event(enable_timing=True)
                ev["s"].record()
                # process-wide start/end range (push/pop ranges are per thread and would miss the kernels
                # launched by autograd's backward thread):  ncu --nvtx --nvtx-include "timed_device"
                ev["nvtx"] = torch.cuda.nvtx.range_start("timed_" + mode)
            if i == W + K - 1:
                ev["e"] = torch.cuda.Event(enable_timing=True)
                ev["e"].record()
                torch.cuda.nvtx.range_end(ev["nvtx"])
                ev["host"] = time.perf_counter() - ev["t0"]     # host time to enqueue the K steps
                torch.cuda.synchronize()
                ev["wall"] = time.perf_counter() - ev["t0"]
                ev["launches"] = profiling.launches()
                ev["ops"] = profiling.summary()
                if ev["clk"] is not None:
                    ev["clk"].__exit__()
                bd.barrier()

        if mode == "e2e":
            torch.cuda.synchronize()
            y_in = y_host.to(device, non_blocking=True)
        smp.predict_blind_bwe(y_in, rid=False, max_steps=W + K, step_hook=hook)
        ms = ev["s"].elapsed_time(ev["e"])
        ev["ms"] = bd.max_over_ranks(ms, device)
        return ev, h2d, d2h

    # `value`: device-resident run without per-call CUDA events; a second, identical pass with the
    # events on supplies the per-kernel table of `roofline` (the events cost host time only)
    use_graph = not (a.no_cuda_graph or a.single_pass)      # ncu launch lists profile the eager launches
    smp.cuda_graph = use_graph
    dev_run, _, _ = timed_run("device")
    graphed = smp._graphed is not None
    smp.cuda_graph = False               # the per-kernel pass needs the wrappers to run (eager)
    if a.single_pass:                    # ncu launch lists: exactly K timed steps, no second pass
        prof_run = dev_run
    else:
        profiling.enable(True)
        prof_run, _, _ = timed_run("device")
        profiling.enable(False)
    smp.cuda_graph = use_graph
    if a.skip_e2e:                       # profiling runs (ncu) only need the device-resident leg
        e2e_run, h2d, d2h = dev_run, 0, 0
    else:
        e2e_run, h2d, d2h = timed_run("e2e")

    # final gather of outputs and filter estimates: the only collective on the path
    xg = bd.gather_rows(y[:, :16].contiguous())
    _ = bd.gather_params(torch.zeros(2, 5, device=device))
    total_chains = chains * world
    value = total_chains * K / (dev_run["ms"] / 1e3)
    e2e_value = total_chains * K / (e2e_run["ms"] / 1e3)

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    ops = prof_run["ops"]
    # dominant STREAMING operator (the fit loop is a latency-bound single-CTA kernel
    # with ~50 KB of traffic: it is listed in "ops" but has no bandwidth roofline)
    stream = {k: v for k, v in ops.items() if v["bytes_avg"] > 1e6}
    dom = max(stream, key=lambda k: stream[k]["ms_total"]) if stream else None
    roof = None
    traffic = None
    tj = {}
    try:                      # per-call DRAM traffic of the dominant operator from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if dom in tj and tj[dom].get("B") == chains:
            traffic = tj[dom]["bytes"]
            if tj[dom].get("algorithmic_bytes"):      # captured at one shape: scale to the average call
                traffic = int(ops[dom]["bytes_avg"] * tj[dom]["bytes"] / tj[dom]["algorithmic_bytes"])
    except Exception:
        pass
    if dom:
        o = ops[dom]
        roof = {"bound": "hbm", "kernel": dom, "achieved": round(o["gbs"], 1), "peak": peak, "unit": "GB/s",
                "frac": round(o["gbs"] / peak, 4), "frac_of_nominal_8TBps": round(o["gbs"] / 8000.0, 4),
                "traffic": traffic,
                "algorithmic_bytes": int(o["bytes_avg"]), "peak_source": peak_src,
                "avg_ms": round(o["ms_avg"], 4), "calls": o["calls"],
                "share_of_step": round(o["ms_total"] / prof_run["ms"], 4),
                "ops": {k: {"ms_avg": round(v["ms_avg"], 4), "GBps": round(v["gbs"], 1), "calls": v["calls"],
                            "share": round(v["ms_total"] / prof_run["ms"], 4)} for k, v in ops.items()},
                "profiled_ms_per_step": round(prof_run["ms"] / K, 3)}
    if roof:
        # the same figures for the dominant operator of the SIGNAL path proper (CQT,
        # STFT filter, statistics) -- the glue kernels above belong to the denoiser body (8f-2)
        glue = ("gn_", "gate_", "resample2")
        sig = {k: v for k, v in stream.items() if not k.startswith(glue)}
        if sig:
            ds = max(sig, key=lambda k: sig[k]["ms_total"])
            o = ops[ds]
            t2 = None
            try:
                if ds in tj and tj[ds].get("B") == chains:
                    t2 = tj[ds]["bytes"]
            except Exception:
                pass
            roof["signal_path"] = {"kernel": ds, "achieved": round(o["gbs"], 1), "frac": round(o["gbs"] / peak, 4),
                                   "traffic": t2, "algorithmic_bytes": int(o["bytes_avg"]),
                                   "avg_ms": round(o["ms_avg"], 4), "calls": o["calls"],
                                   "share_of_step": round(o["ms_total"] / prof_run["ms"], 4)}
    op_roof = operator_probe(device, peak) if not a.skip_e2e else None
    cpu = None
    if not a.no_cpu_baseline:
        cpu = cpu_baseline_sample(steps=1)
    line = {
        "metric": "blind-BWE sampler chain-steps/s", "value": round(value, 4), "unit": "chain-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(dev_run["ms"] / K, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1] per-GPU slice: {chains} chains/GPU x T={AUDIO_LEN} @ {SR} Hz, "
                               "random-init CQTDiff+ (44.5M params), 35-step blind EDM sampler (order 2), "
                               "NFFT=4096, K=5, fit max_iter=100",
                   "chains_per_gpu": chains, "total_chains": total_chains, "audio_len": AUDIO_LEN,
                   "sample_rate": SR, "nfft": NFFT, "sampler_steps_per_s": round(K / (dev_run["ms"] / 1e3), 4),
                   "parallelism": f"replicas x{world} (independent chains, no data-path collective)",
                   "l2_note": "per-step working set (activations of the 44.5M-param U-Net at B=8, >10 GB) far exceeds the 126 MB L2",
                   "tf32": bool(torch.backends.cudnn.allow_tf32),
                   "cuda_graph": graphed,
                   "host_enqueue_ms_per_step": round(dev_run.get("host", 0.0) * 1e3 / K, 1)},
        "e2e": {"value": round(e2e_value, 4), "unit": "chain-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": round(e2e_run["ms"] / K, 3)},
        "gpu_launches": dev_run["launches"],
        "clocks": dev_run["clk"].summary() if dev_run.get("clk") else None,
        "roofline": roof, "operator_roofline": op_roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# "library" baseline on the same GPU: the reference's own torch formulation of the operator
# (torch.stft / torch.istft -> cuFFT, masked index ops, autograd) -- measurement only.
def torch_apply_filter(x, H, nfft):
    """utils/blind_bwe_utils.py:6-39 as the reference runs it on a GPU."""
    window = torch.hamming_window(window_length=nfft).to(x.device)
    xp = torch.cat((x, torch.zeros(*x.shape[:-1], nfft).to(x.device)), 1)
    X = torch.stft(xp, nfft, hop_length=nfft // 2, window=window, center=False, onesided=True, return_complex=True)
    X = X * H.unsqueeze(-1)
    return torch.istft(X, nfft, hop_length=nfft // 2, window=window, center=False, return_complex=False)[:, :x.shape[-1]]


def torch_design_filter(fc, A, f):
    """utils/blind_bwe_utils.py:82-119 (multi-slope branch)."""
    H = torch.zeros(f.shape).to(f.device)
    H[f < fc[0]] = 1
    H[f >= fc[0]] = 10 ** (A[0] * torch.log2(f[f >= fc[0]] / fc[0]) / 20)
    for i in range(1, len(fc)):
        H[f >= fc[i]] = 10 ** (A[i] * torch.log2(f[f >= fc[i]] / fc[i]) / 20) * H[f >= fc[i]][0]
    return H


def torch_fit_loop(xden, y, params, f, nfft, iters=100):
    """testing/blind_bwe_sampler.py:556-590 as the reference runs it (Python loop, autograd, host syncs)."""
    window = torch.hamming_window(window_length=nfft).to(y.device)

    def stft(v):
        vp = torch.cat((v, torch.zeros(*v.shape[:-1], nfft).to(v.device)), 1)
        return torch.view_as_real(torch.stft(vp, nfft, hop_length=nfft // 2, window=window, center=False,
                                             onesided=True, return_complex=True))
    Xd, Y = stft(xden), stft(y)
    Xm = torch.sqrt(Xd[..., 0] ** 2 + Xd[..., 1] ** 2)
    Ym = torch.sqrt(Y[..., 0] ** 2 + Y[..., 1] ** 2)
    w = torch.sqrt(torch.linspace(0, 1, Xm.shape[1]).to(y.device)).unsqueeze(-1)
    mu = torch.tensor([1000.0, 10.0], device=y.device)
    p = params.clone()
    for _ in range(iters):
        p.requires_grad = True
        H = torch_design_filter(p[0], p[1], f)
        norm = torch.linalg.norm((Xm * H.unsqueeze(-1) * w).reshape(-1) - (Ym * w).reshape(-1), ord=2)
        g = torch.autograd.grad(norm, p, create_graph=True)
        p = p - mu.unsqueeze(1) * g[0]
        p.detach_()
        p[0, 0] = torch.clamp(p[0, 0], min=20, max=SR // 2)
        for k in range(1, p.shape[1]):
            p[0, k] = torch.clamp(p[0, k], min=p[0, k - 1] + 1, max=SR // 2)
        p[1, 0] = torch.clamp(p[1, 0], min=-50, max=-1)
        for k in range(1, p.shape[1]):
            p[1, k] = torch.clamp(p[1, k], min=-50, max=p[1, k - 1])
    return p


def operator_probe(device, peak):
    """"operator % of HBM roofline" (second half of BASELINE.json's metric): the streaming
    operators alone at a batch that fills the GPU (config 4: B=512 x T=2^17, NFFT=4096),
    CUDA events, L2 flushed between iterations."""
    from babe_b200 import ops
    from cqt_nsgt_pytorch import CQT_nsgt
    B, T = 512, 1 << 17
    x = torch.randn(B, T, device=device) * 0.063
    y = torch.randn(B, T, device=device) * 0.063
    out = torch.empty_like(x)
    f = torch.fft.rfftfreq(NFFT, d=1 / SR).to(device)
    fc = torch.tensor([300.0, 600.0, 1000.0, 3000.0, 6000.0], device=device)
    A = torch.tensor([-10.0, -15.0, -20.0, -30.0, -40.0], device=device)
    cq = CQT_nsgt(7, 64, mode="oct", window=("kaiser", 1), fs=SR, audio_len=AUDIO_LEN, device=device)
    xc = torch.randn(64, AUDIO_LEN, device=device) * 0.063
    coefs = [None]
    cases = {
        "apply_filter": (lambda: ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, out=out), 8 * B * T),
        "apply_filter_adj": (lambda: ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, adjoint=True, out=out), 8 * B * T),
        "stft_stats": (lambda: ops.stft_stats(x, y, NFFT), 8 * B * T),
        "cqt_analysis_B64": (lambda: coefs.__setitem__(0, cq.fwd(xc.unsqueeze(1))),
                             64 * (4 * AUDIO_LEN + 8 * cq.plan.coef_per_row)),
        "cqt_synthesis_B64": (lambda: cq.bwd(coefs[0]), 64 * (4 * AUDIO_LEN + 8 * cq.plan.coef_per_row)),
        "hpf_DC_B64": (lambda: cq.apply_hpf_DC(xc), 64 * 8 * AUDIO_LEN),
    }
    Hd = ops.design_filter(fc, A, f, strict=False)
    x8, y8 = x[:8, :AUDIO_LEN // 2].contiguous(), y[:8, :AUDIO_LEN // 2].contiguous()
    p0 = torch.tensor([[280.0, 285, 290, 295, 300], [-15.0, -17, -20, -25, -30]], device=device)
    cases["torch+cuFFT apply_filter (library baseline)"] = (lambda: torch_apply_filter(x, Hd, NFFT), 8 * B * T)
    cases["torch+cuFFT hpf (rfft/irfft, library baseline) B64"] = (
        lambda: torch.fft.irfft(torch.fft.rfft(xc) * cq.plan.Hhpf, n=AUDIO_LEN), 64 * 8 * AUDIO_LEN)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    res = {"shape": f"B={B} x T={T} (STFT ops), B=64 x T={AUDIO_LEN} (CQT ops)", "l2": "flushed between iterations"}
    for name, (fn, nbytes) in cases.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(7):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ms = sorted(ts)[len(ts) // 2]
        gbs = nbytes / 1e9 / (ms / 1e3)
        res[name] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 4)}
    # filter fit: one launch vs the reference's Python loop on the same GPU (latency, 8 rows)
    from babe_b200 import sampler as _s
    fit = _s.FilterFit(nfft=NFFT, sample_rate=SR, device=device)
    for name, fn in (("fit_params 100 it (stats + 1 launch)", lambda: fit(x8, y8, p0.clone())),
                     ("torch fit loop 100 it (library baseline)", lambda: torch_fit_loop(x8, y8, p0, f, NFFT))):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        res[name] = {"ms": round((time.perf_counter() - t0) * 1e3, 3)}
    return res


# ---------------------------------------------------------------------------
def cpu_world(chains, seed, max_iter=100, audio_len=CPU_AUDIO_LEN):
    """The reference algorithm on the host: oracle operator + oracle CQT + the same
    PyTorch denoiser body on the CPU."""
    from babe_b200 import denoiser, sampler
    from oracle import blind_sampler as obs, filter_fit as ofit
    from oracle import stft_filter as sf
    from oracle.cqt_shim import OracleCQT
    args = sampler.make_args(sample_rate=SR, audio_len=audio_len, NFFT=NFFT, max_iter=max_iter)
    torch.manual_seed(0)
    cqt = OracleCQT(7, 64, window=("kaiser", 1), fs=SR, audio_len=audio_len, dtype=torch.float32)
    net = denoiser.CQTDiffPlus(args, "cpu", cqt=cqt)
    for p in net.parameters():
        p.requires_grad_(False)
    x = piano_like(chains, audio_len, SR, 1234 + seed)
    f = torch.fft.rfftfreq(NFFT, d=1 / SR)
    y = sf.apply_filter(x, sf.design_filter([1000.0], [-20.0], f), NFFT)
    cfg = obs.SamplerConfig(T=35, audio_len=audio_len)
    cfg.fit = ofit.FitConfig(nfft=NFFT, sample_rate=SR, max_iter=max_iter)
    return cfg, net, cqt, y


def cpu_time_steps(steps, warmup):
    from oracle import blind_sampler as obs
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, net, cqt, y = cpu_world(1, 0)
    marks = []
    trace = []

    class Marker(list):
        def append(self, item):
            marks.append(time.perf_counter())

    t0 = time.perf_counter()
    torch.manual_seed(42)
    obs.predict_blind_bwe(cfg, net, cqt.apply_hpf_DC, y, steps=warmup + steps, trace=Marker())
    start = t0 if warmup == 0 else marks[warmup - 1]
    # scale the quarter-length segment to the full chain length
    return (marks[warmup + steps - 1] - start) * (AUDIO_LEN / CPU_AUDIO_LEN)


def cpu_baseline_sample(steps=1):
    dt = cpu_time_steps(steps, 0)
    return {"value": round(steps / dt, 5), "unit": "chain-steps/s", "cores": torch.get_num_threads(),
            "kind": "port", "seconds": round(dt, 2),
            "sample": f"1 chain x {steps} sampler step(s) (2 denoiser fwd+bwd + 2 filter fits + 2 guidance "
                      f"evaluations each) on a T'={CPU_AUDIO_LEN} segment, time scaled x{AUDIO_LEN // CPU_AUDIO_LEN} "
                      f"to T={AUDIO_LEN}; oracle operator/CQT + PyTorch CPU denoiser"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = a.steps, a.warmup
    dt = cpu_time_steps(K, W)
    v = K / dt
    cpu = {"value": round(v, 5), "unit": "chain-steps/s", "cores": torch.get_num_threads(), "kind": "port",
           "sample": f"1 chain x {K} timed sampler steps after {W} warm-up steps on a T'={CPU_AUDIO_LEN} segment, "
                     f"time scaled x{AUDIO_LEN // CPU_AUDIO_LEN} to T={AUDIO_LEN} "
                     "(the reference is Python: the oracle port restates it; /root/reference is not on the GPU box)"}
    line = {"impl": "reference", "metric": "blind-BWE sampler chain-steps/s", "value": round(v, 5),
            "unit": "chain-steps/s", "n_gpus": a.gpus, "steps": K, "warmup": W,
            "ms_per_step": round(dt / K * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] per-GPU slice, bounded sample: 1 chain, "
                                   f"T'={CPU_AUDIO_LEN} scaled to T={AUDIO_LEN} @ {SR} Hz, random-init CQTDiff+, "
                                   "35-step blind EDM sampler, NFFT=4096, K=5, fit max_iter=100; host cores only"},
            "cpu_baseline": cpu,
            "e2e": {"value": round(v, 5), "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
def run_operator_sweep(a):
    """Config 4: operator microbenchmark, STFT -> parametric filter -> iSTFT + gradients."""
    from babe_b200 import build, ops, sampler
    build.build()
    dev = torch.device("cuda")
    peak, src = measured_peak()
    f = torch.fft.rfftfreq(NFFT, d=1 / SR).to(dev)
    fc = torch.tensor([300.0, 600.0, 1000.0, 3000.0, 6000.0], device=dev)
    A = torch.tensor([-10.0, -15.0, -20.0, -30.0, -40.0], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    shapes = [(1, 1 << 14), (8, 1 << 17), (64, 1 << 17), (8, AUDIO_LEN), (64, AUDIO_LEN), (512, 1 << 17),
              (512, 1 << 20)] if not a.shapes else [tuple(int(v) for v in s.split("x")) for s in a.shapes.split(",")]
    for B, T in shapes:
        if B * T * 4 > (8 << 30):
            continue
        x = torch.randn(B, T, device=dev) * 0.063
        y = torch.randn(B, T, device=dev) * 0.063
        out = torch.empty_like(x)
        fit = sampler.FilterFit(nfft=NFFT, sample_rate=SR, device=dev)
        p = torch.tensor([[280.0, 285, 290, 295, 300], [-15.0, -17, -20, -25, -30]], device=dev)
        fp = torch.stack((fc, A))
        spec_bytes = 8 * B * (NFFT // 2 + 1) * ops.num_frames(T, NFFT)

        def recg(x=x, y=y, fp=fp):
            xg = x.detach().requires_grad_(True)
            n = sampler.rec_guidance_norms(xg, y, f, fp, NFFT)
            torch.autograd.grad(n.sum(), xg)

        cases = {
            "apply_filter fwd (8BT)": (lambda: ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, out=out), 8 * B * T),
            "apply_filter adj (8BT)": (lambda: ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, adjoint=True, out=out), 8 * B * T),
            "fit statistics (8BT)": (lambda: ops.stft_stats(x, y, NFFT), 8 * B * T),
            "fit loop 100 it": (lambda: fit(x, y, p.clone(), abc=abc), 0),
            "rec-guidance norms + grad (20BT, r materialised)": (recg, 20 * B * T),
            "apply_stft a1 (4BT + 8BFM)": (lambda: ops.stft(x, NFFT), 4 * B * T + spec_bytes),
            "apply_filter_istft a2 (8BFM + 4BT)": (lambda: ops.istft(X, NFFT), 4 * B * T + spec_bytes),
        }
        abc = ops.stft_stats(x, y, NFFT)
        X = ops.stft(x, NFFT)
        for name, (fn, nbytes) in cases.items():
            for _ in range(3):
                fn()
            ts = []
            for _ in range(a.iters):
                if B * T * 8 < (252 << 20):
                    flush.fill_(1)
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                fn()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ms = sorted(ts)[len(ts) // 2]
            gbs = nbytes / 1e9 / (ms / 1e3) if nbytes else 0.0
            rows.append({"B": B, "T": T, "op": name, "ms": round(ms, 4), "GBps": round(gbs, 1),
                         "frac_of_peak": round(gbs / peak, 4)})
            print(json.dumps(rows[-1]))
    print(json.dumps({"mode": "operator", "peak_GBps": peak, "peak_source": src, "rows": len(rows)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="sampler", choices=["sampler", "operator"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shapes", default="")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: skip the host-buffer leg")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="run the denoiser eagerly instead of replaying captured CUDA graphs")
    ap.add_argument("--single-pass", action="store_true",
                    help="profiling aid: skip the second (per-kernel event timing) pass")
    ap.add_argument("--no-autotune", action="store_true", help="profiling aid: cudnn.benchmark off")
    a = ap.parse_args()
    a.steps = max(1, a.steps)
    if a.impl != "reference":
        a.warmup = max(1, a.warmup)          # the timed region starts after a completed (untimed) step
    if a.impl == "reference":
        return run_reference(a)
    if a.mode == "operator":
        return run_operator_sweep(a)
    return run_ours(a)


if __name__ == "__main__":
    main()
    import torch.distributed as _dist
    if _dist.is_available() and _dist.is_initialized():
        _dist.destroy_process_group()
