#!/usr/bin/env python
"""bench.py — benchmark of the B200-native R-YOLOv4 hot path (contract: task prompt ④; BASELINE.json metric
"training-step img/s at 800x800 bs=32, 1/2/4/8 B200; rotated-NMS boxes/s").

Headline (`value`, every N): the model/size of BASELINE.json configs[1] — synthetic 800x800, bs=32 PER GPU, yolov4 / csl /
nc=2, random weights (reference init, train.py:28-33) — driven through ONE FULL TRAINING STEP (train.py:184-202):
train-mode forward (batch-statistics BatchNorm), ComputeCSLLoss value + gradient, conv-stack backward (dgrad + wgrad),
the bucketed NCCL all-reduce of the fp32 gradients when N>1, SGD(momentum .937, nesterov).  It is a superset of
configs[1]'s "forward+loss", whose img/s is reported next to it (config.fwd_loss_img_s).  N>1: one rank per GPU, its own
32-image shard of the global batch (weak scaling).

The other BASELINE configs ride in `aux` of the same JSON line:
  aux.train_v7      configs[2]: yolov7 / csl / nc=16 (data/DOTA.yaml) full train step, at EVERY N (same DDP path)
  aux.kfloss        configs[3]: KFIoU loss, 50 000 pairs/image x 256 images: GB/s vs the measured HBM peak (N=1, rank 0)
  aux.post_process  configs[4]: post_process on 64 images x 100 000 rows: rows/s, front-end GB/s, IoU pairs/s (N=1)
each with its CPU baseline (reference code staged under oracle/_ref when available, else the oracle port).
configs[0] (416^2 bs=2 CPU smoke) is a parity case: tests/test_config1.py.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-cpu] [--no-aux]

--impl reference times the CPU arm of the headline: the REFERENCE'S OWN Yolo + ComputeCSLLoss + loss.backward() +
SGD step (oracle/_ref, staged by oracle/make_ref.py; oracle port if absent) on all host cores, on a bounded sample
(2 of the 32 images per step).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(anchors=[[12, 16, 19, 36, 40, 28], [36, 75, 76, 55, 72, 146], [142, 110, 192, 243, 459, 401]],
           angles=[-90, -60, -30, 0, 30, 60])                                   # data/hyp.yaml:2-7
HYP = dict(fl_gamma=0.0, box=0.05, obj=1.0, obj_pw=1.0, cls=0.5, cls_pw=1.0)   # data/hyp.yaml:11-17
S, BS, PER_IMG = 800, 32, 100
NC = 2          # headline workload (tools/ import it)
# 2*MAC over the convs @800^2: forward; dgrad = forward minus the stem; wgrad = forward
WORKLOADS = {
    "train_v4": dict(ver="yolov4", nc=2, fwd_gflop=218.14, stem_gflop=1.106,
                     name="yolov4/csl/nc2 800x800 full train step (fwd + CSL loss + bwd + SGD)"),
    "train_v7": dict(ver="yolov7", nc=16, fwd_gflop=168.38, stem_gflop=1.106,
                     name="yolov7/csl/nc16 800x800 full train step (fwd + CSL loss + bwd + SGD)"),
}
METRIC, UNIT = "training-step img/s, 800x800, bs=32 per GPU (synthetic, random weights)", "img/s"


def weights_init_normal(m):                                                   # train.py:28-33
    cn = m.__class__.__name__
    if cn.find("Conv2d") != -1:
        torch.nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif cn.find("BatchNorm2d") != -1:
        torch.nn.init.normal_(m.weight.data, 1.0, 0.02)
        torch.nn.init.constant_(m.bias.data, 0.0)


def gaussian_label(label, num_class=180, u=0, sig=6.0):
    x = np.arange(-num_class / 2, num_class / 2)
    y = np.exp(-(x - u) ** 2 / (2 * sig ** 2))
    i = int(num_class / 2 - label)
    return np.concatenate([y[i:], y[:i]], axis=0)


def make_targets(seed, bs, nc):
    g = np.random.default_rng(seed)
    rows = []
    for b in range(bs):
        for _ in range(PER_IMG):
            w = g.uniform(0.02, 0.15)
            h = min(w * g.uniform(1, 3), 0.9)
            th = g.uniform(-np.pi / 2, np.pi / 2 - 1e-3)
            rows.append([b, float(g.integers(0, nc)), g.uniform(0.05, 0.95), g.uniform(0.05, 0.95), w, h, th]
                        + list(gaussian_label(th * 180 / np.pi + 90)))
    return torch.tensor(np.array(rows), dtype=torch.float32)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}", "--format=csv,noheader",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        num = lambda s: float(s.split()[0])
        load = [r for r in rows if num(r[3]) > 300] or rows
        reasons = set()
        for r in load:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([num(r[1]) for r in load])), "sm_max_mhz": num(rows[0][2]),
                "power_w_max": max(num(r[3]) for r in rows), "samples": len(rows), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU arms
def _reference_ns():
    try:
        from oracle import make_ref
        return make_ref.load()
    except Exception as e:                                         # staging absent / broken: fall back to the port
        print(f"[bench] reference staging unavailable ({e!r}); using the oracle port", file=sys.stderr)
        return None


def cpu_train_arm(wl, steps, warmup, sample_imgs=2):
    """One reference training step (forward, CSL loss, autograd backward, SGD; train.py:195-202) on the host cores, on a
    bounded sample of the workload.  kind "reference": the reference's own nn.Modules and loss (oracle/_ref);
    kind "port": oracle/model_cpu.py + oracle/hotpath.py.  Returns (img/s, cores, kind, sample description)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    img = torch.rand(sample_imgs, 3, S, S)
    tg = make_targets(0, sample_imgs, wl["nc"])
    ns = _reference_ns()
    ts = []
    if ns is not None:
        kind = "reference"
        m = ns.Yolo(wl["nc"], CFG, "csl", wl["ver"])
        m.apply(weights_init_normal)
        m.train()
        crit = ns.ComputeCSLLoss(m, HYP)
        opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.937, nesterov=True)
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = m(img, training=True)
            loss, _ = crit(out, tg)
            loss.backward()
            opt.step()
            opt.zero_grad()
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    else:
        kind = "port"
        from oracle import hotpath as hp
        from oracle import model_cpu
        import ryolo_b200 as R
        m = R.Yolo(wl["nc"], CFG, "csl", wl["ver"])
        m.apply(weights_init_normal)
        pnames = {k for k, _ in m.named_parameters()}
        sd = {k: v.clone().requires_grad_(k in pnames) for k, v in m.state_dict().items()}
        opt = torch.optim.SGD([sd[k] for k in sd if k in pnames], lr=0.01, momentum=0.937, nesterov=True)
        an = hp.make_anchors(CFG["anchors"])
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad()
            levels, _, stats = model_cpu.forward(sd, img, wl["ver"], "csl", wl["nc"], train=True, decode=False)
            loss, _ = hp.csl_loss(levels, tg, an, wl["nc"], HYP)
            loss.backward()
            opt.step()
            with torch.no_grad():
                for k, v in stats.items():
                    sd[k].copy_(v)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    src = "the reference's own Yolo/ComputeCSLLoss (oracle/_ref)" if kind == "reference" else "oracle port"
    return sample_imgs / float(np.mean(ts)), cores, kind, \
        f"{sample_imgs} of the {BS} images per step, {steps} timed step(s), {src}, torch CPU fp32 ({cores} threads)"


def cpu_kfloss_arm():
    """BASELINE configs[3] on the host: (i) the reference's KFLoss at N = 50 000 (ONE image: its accidental [N,1]+[N]
    broadcast, lib/loss.py:114,148, allocates N x N fp32 = 10 GB, so that is the largest size it can run);
    (ii) the O(N) torch restatement (oracle/hotpath.kf_loss) on 32 images' worth of pairs."""
    from oracle import hotpath as hp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(4)

    def pairs(n):
        pr = torch.cat((torch.rand(n, 2, generator=gen) * 2 - 0.5, torch.rand(n, 2, generator=gen) * 8 + 0.5,
                        (torch.rand(n, 1, generator=gen) - 0.5) * 3.14), 1)
        tg = torch.cat((torch.rand(n, 2, generator=gen), torch.rand(n, 2, generator=gen) * 8 + 0.5,
                        (torch.rand(n, 1, generator=gen) - 0.5) * 3.14), 1)
        return pr, tg

    out = {"cores": cores}
    n = 50000 * 32
    pr, tg = pairs(n)
    p = pr.clone().requires_grad_(True)
    t0 = time.perf_counter()
    loss, _ = hp.kf_loss(p, tg)
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    out["port_O(N)"] = {"pairs": n, "fwd_s": t1 - t0, "fwd_bwd_s": t2 - t0, "fwd_gbs": n * 44 / (t1 - t0) / 1e9,
                        "fwd_bwd_gbs": n * 64 / (t2 - t0) / 1e9, "kind": "port"}
    ns = _reference_ns()
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 0
    if ns is not None and avail > 48 * 2 ** 30:                   # fwd+bwd of the N x N form peaks near 40 GB
        n1 = 50000
        pr, tg = pairs(n1)
        p = pr.clone().requires_grad_(True)
        kf = ns.KFLoss()
        t0 = time.perf_counter()
        loss, _ = kf(p, tg)
        t1 = time.perf_counter()
        loss.backward()
        t2 = time.perf_counter()
        out["reference_NxN"] = {"pairs": n1, "fwd_s": t1 - t0, "fwd_bwd_s": t2 - t0,
                                "fwd_gbs": n1 * 44 / (t1 - t0) / 1e9, "fwd_bwd_gbs": n1 * 64 / (t2 - t0) / 1e9,
                                "kind": "reference",
                                "note": "one image (50 000 pairs): the N x N broadcast of lib/loss.py:114,148 needs "
                                        "10 GB per intermediate; 32 images at once would need 10 TB"}
    else:
        out["reference_NxN"] = {"unavailable": "reference not staged" if ns is None else
                                f"host has {avail / 2 ** 30:.0f} GiB free; the reference's N x N form needs ~40 GiB"}
    return out


def make_pred(B, Rr, nc, gen, device):
    """BASELINE configs[4] candidates: boxes jittered around 200 centres/image so that NMS suppresses."""
    centres = torch.rand(B, 200, 2, device=device, generator=gen) * 800
    pick = torch.randint(0, 200, (B, Rr), device=device, generator=gen)
    xy = torch.gather(centres, 1, pick[..., None].expand(B, Rr, 2)) + \
        torch.randn(B, Rr, 2, device=device, generator=gen) * 6
    w = torch.rand(B, Rr, 1, device=device, generator=gen) * 116 + 4
    h = w * (1 + 3 * torch.rand(B, Rr, 1, device=device, generator=gen))
    th = (torch.rand(B, Rr, 1, device=device, generator=gen) - 0.5) * np.pi * 0.9999
    oc = torch.rand(B, Rr, 1 + nc, device=device, generator=gen)
    return torch.cat((xy, w, h, th, oc), 2).contiguous()


def cpu_post_process_arm(conf, iou, imgs=2):
    """post_process (lib/general.py:136-183) on `imgs` images x 100 000 rows on the host: the reference's own glue with
    the single-thread C++ restatement of detectron2 nms_rotated behind it (like upstream's CPU op), or the port."""
    gen = torch.Generator().manual_seed(5)
    pred = make_pred(imgs, 100000, 2, gen, "cpu")
    ns = _reference_ns()
    if ns is not None:
        fn, kind = ns.post_process, "reference"
    else:
        from oracle import hotpath as hp
        fn, kind = hp.post_process, "port"
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    outs = fn(pred, conf, iou)
    dt = time.perf_counter() - t0
    return {"value": imgs * 100000 / dt, "unit": "rows/s", "cores": 1, "kind": kind, "seconds": dt,
            "sample": f"{imgs} of the 64 images (100 000 rows each); python glue on torch CPU, rotated NMS single-threaded "
                      f"C++ (oracle/rotated_ops.cpp) like upstream's CPU operator",
            "survivors": [int(o.shape[0]) for o in outs]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS["train_v4"]
    steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
    v, cores, kind, sample = cpu_train_arm(wl, steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": 1e3 * 2 / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": BS, "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
class Env:
    pass


def bench_train(env, wl, steps, warm, sample_clocks, use_graph=True):
    """One training workload on this rank's GPU: resident (`value`), instrumented roofline pass, forward+loss only,
    and end to end from pinned host memory.  All timings are CUDA events, max over ranks."""
    import ryolo_b200 as R
    from ryolo_b200 import _lib as L
    from ryolo_b200 import ops
    dev, world, rank = env.dev, env.world, env.rank
    nc = wl["nc"]
    torch.manual_seed(42)
    model = R.Yolo(nc, CFG, "csl", wl["ver"])
    model.apply(weights_init_normal)
    model = model.to(dev).train()
    crit = R.ComputeCSLLoss(model, HYP)
    crit.sync_items = False
    trainer = R.TrainStep(model, crit, lr=0.01, momentum=0.937, nesterov=True)      # train.py:156
    host_imgs = [torch.rand(BS, 3, S, S).pin_memory() for _ in range(2)]
    host_tg = [make_targets(rank * 2 + i, BS, nc).pin_memory() for i in range(2)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    dev_tg = [h.to(dev) for h in host_tg]

    def resident(n):
        for i in range(n):
            trainer(dev_imgs[i & 1], dev_tg[i & 1])

    resident(warm)
    sampler = ClockSampler(env.local) if (sample_clocks and rank == 0) else None
    l0 = L.LAUNCHES[0]
    ms_total = env.timed(resident, steps)                     # headline: no per-launch instrumentation inside
    launches = L.LAUNCHES[0] - l0
    clocks = sampler.stop() if sampler else None

    # ---- roofline pass: the same steps again with a CUDA-event pair around every tensor-core launch
    psteps = min(steps, 5)
    ops.PROFILE = []
    env.timed(resident, psteps)
    prof, ops.PROFILE = ops.PROFILE, None
    tc_ms = {}
    for tag, a, b in prof:
        tc_ms[tag[0]] = tc_ms.get(tag[0], 0.0) + a.elapsed_time(b) / psteps
    # ... and once more with the weight-gradient GEMMs on the main stream: an event pair then brackets the kernel alone
    # (in the headline configuration a conv / dgrad launch also waits for the SMs a side-stream wgrad CTA still holds)
    from ryolo_b200.model import backward as BW
    tc_ser = {}
    side = BW.WGRAD_SIDE_STREAM
    try:
        BW.WGRAD_SIDE_STREAM = False
        resident(1)
        ops.PROFILE = []
        env.timed(resident, psteps)
        prof, ops.PROFILE = ops.PROFILE, None
        for tag, a, b in prof:
            tc_ser[tag[0]] = tc_ser.get(tag[0], 0.0) + a.elapsed_time(b) / psteps
    finally:
        BW.WGRAD_SIDE_STREAM = side
        ops.PROFILE = None

    # ---- forward + loss only (BASELINE configs[1] wording), same model and inputs
    def fwd_loss(n):
        for i in range(n):
            lv = model(dev_imgs[i & 1], training=True)
            crit.value_and_grad(lv, dev_tg[i & 1])

    fwd_loss(2)
    fl_steps = min(steps, 5)
    ms_fl = env.timed(fwd_loss, fl_steps)

    # ---- end-to-end: pinned host -> device copies (prefetched on a copy stream) + loss read-back, every step
    copy_stream = torch.cuda.Stream(dev)
    stage_i = [torch.empty_like(dev_imgs[0]) for _ in range(2)]
    stage_t = [torch.empty_like(dev_tg[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    out_host = torch.empty(8, dtype=torch.float32).pin_memory()

    def upload(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            stage_i[s].copy_(host_imgs[s], non_blocking=True)
            stage_t[s].copy_(host_tg[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e(n):
        cur = torch.cuda.current_stream()
        for s in range(2):
            freed[s].record(cur)
        upload(0)
        for i in range(n):
            s = i & 1
            if i + 1 < n:
                upload(i + 1)
            cur.wait_event(ready[s])
            items = trainer(stage_i[s], stage_t[s])
            freed[s].record(cur)
            out_host.copy_(items, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_host[4])

    e2e(2)
    ms_e2e = env.timed(e2e, steps)
    eager = dict(img_s=world * BS * steps / ms_total * 1e3, ms_per_step=ms_total / steps,
                 e2e_img_s=world * BS * steps / ms_e2e * 1e3)
    execution = f"eager: {launches // steps} kernel launches per step through the C ABI"
    # ---- the same step as ONE CUDA graph launch (TrainStep.capture; single-GPU path)
    if world == 1 and use_graph:
        try:
            l1 = L.LAUNCHES[0]
            trainer.capture(dev_imgs[0], dev_tg[0], target_capacity=dev_tg[0].shape[0], warmup=1)
            per_step = (L.LAUNCHES[0] - l1) // 2                      # one warm-up step + the captured one

            def resident_g(n):
                for i in range(n):
                    trainer.replay(dev_imgs[i & 1], dev_tg[i & 1])

            def e2e_g(n):
                cur = torch.cuda.current_stream()
                for s in range(2):
                    freed[s].record(cur)
                upload(0)
                for i in range(n):
                    s = i & 1
                    if i + 1 < n:
                        upload(i + 1)
                    cur.wait_event(ready[s])
                    items = trainer.replay(stage_i[s], stage_t[s])
                    freed[s].record(cur)
                    out_host.copy_(items, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            resident_g(2)
            sampler = ClockSampler(env.local) if (sample_clocks and rank == 0) else None
            ms_g = env.timed(resident_g, steps)
            if sampler:
                clocks = sampler.stop()
            e2e_g(2)
            ms_e2e_g = env.timed(e2e_g, steps)
            if ms_g < ms_total:
                ms_total, ms_e2e, launches = ms_g, ms_e2e_g, per_step * steps
                execution = f"one CUDA graph launch per step ({per_step} kernels captured; TrainStep.capture/replay)"
            eager["graph_img_s"] = world * BS * steps / ms_g * 1e3
        except Exception as e:                                        # noqa: BLE001 - report and keep the eager numbers
            eager["graph_error"] = repr(e)[:300]
    res = dict(value=world * BS * steps / ms_total * 1e3, ms_per_step=ms_total / steps, launches=launches, clocks=clocks,
               eager=eager, execution=execution,
               tc_ms=tc_ms, tc_ser=tc_ser, fwd_loss_img_s=world * BS * fl_steps / ms_fl * 1e3, fwd_loss_ms=ms_fl / fl_steps,
               e2e=dict(value=world * BS * steps / ms_e2e * 1e3, unit=UNIT,
                        h2d_bytes_per_step=int(host_imgs[0].numel() * 4 + host_tg[0].numel() * 4),
                        d2h_bytes_per_step=32, ms_per_step=ms_e2e / steps),
               grad_mb=trainer.grad.numel() * 4 / 1e6)
    del trainer, model, crit, dev_imgs, dev_tg, stage_i, stage_t, host_imgs, host_tg
    torch.cuda.empty_cache()
    return res


def roofline_of(res, wl, pk, pk_src):
    # dominant kernel = conv_fwd_kernel (forward + dgrad launches, main stream).  The wgrad GEMMs run on a side
    # stream overlapped with dgrad / BN backward; their per-launch durations include time-slicing -> listed beside.
    gflop = (2 * wl["fwd_gflop"] - wl["stem_gflop"]) * BS          # forward + dgrad (no stem), per step
    tc_ms = res["tc_ms"]
    tc_total = tc_ms.get("conv", 0.0) + tc_ms.get("dgrad", 0.0)
    tflops = gflop / max(tc_total, 1e-9)                           # GFLOP / ms == TFLOP/s
    peak = pk["bf16_tflops_sustained"]
    wg_gflop = wl["fwd_gflop"] * BS
    out = {"kernel": "conv_fwd_kernel (tcgen05 implicit GEMM: forward + dgrad launches)", "bound": "tensor",
           "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak, "traffic": None,
           "peak_source": f"{pk_src} bf16_tflops_sustained",
           "timing": "CUDA-event pair around every launch, separate pass after the headline timing",
           "ms_per_step": {k: round(v, 3) for k, v in tc_ms.items()}, "algorithmic_gflop_per_step": gflop,
           "wgrad": {"kernel": "conv_wgrad_kernel (side stream, overlapped)", "algorithmic_gflop_per_step": wg_gflop,
                     "achieved": wg_gflop / max(tc_ms.get("wgrad", 0.0), 1e-9)},
           "step_tensor_frac": 3 * wl["fwd_gflop"] * BS / res["ms_per_step"] / peak}
    ser = res.get("tc_ser") or {}
    if ser.get("conv") and ser.get("dgrad"):
        t = gflop / (ser["conv"] + ser["dgrad"])
        out["serialized"] = {"note": "same step with the wgrad GEMMs on the main stream: every event pair brackets one "
                                     "kernel alone (no wait for SMs held by a side-stream CTA)",
                             "achieved": t, "frac": t / peak, "ms_per_step": {k: round(v, 3) for k, v in ser.items()},
                             "wgrad_achieved": wg_gflop / max(ser.get("wgrad", 0.0), 1e-9)}
    for name in ("r02_conv_traffic.json", "r01_conv_traffic.json"):   # made by tools/conv_traffic.py from an ncu pass
        tp = os.path.join(ROOT, "profiles", name)
        if wl["ver"] == "yolov4" and os.path.exists(tp):
            tj = json.load(open(tp))
            out["traffic"] = tj["dram_bytes_per_launch"]
            out["traffic_source"] = f"profiles/{name}: mean DRAM bytes over {tj['launches']} launches (one step, cold " \
                                    f"cache); algorithmic {tj['algorithmic_bytes_per_launch']:.3g} B/launch"
            break
    return out


def bench_kfloss(env, pk, iters=5):
    import ryolo_b200 as R
    dev = env.dev
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    N = 50000 * 256                                              # 563 MB of pairs: larger than the 126 MB L2
    pr = torch.cat((torch.rand(N, 2, device=dev) * 2 - 0.5, torch.rand(N, 2, device=dev) * 8 + 0.5,
                    (torch.rand(N, 1, device=dev) - 0.5) * 3.14), 1).contiguous()
    tg = torch.cat((torch.rand(N, 2, device=dev), torch.rand(N, 2, device=dev) * 8 + 0.5,
                    (torch.rand(N, 1, device=dev) - 0.5) * 3.14), 1).contiguous()
    kf = R.KFLoss()
    out = {"workload": "KFLoss (lib/loss.py:81-150) on 50 000 pairs/image x 256 images", "pairs": N,
           "bytes_per_pair": {"fwd": 44, "fwd_bwd": 64}, "peak_gbs": pk["hbm_gbs"],
           "l2": "563 MB of pairs per launch (> 126 MB L2) + 256 MB flush between timed groups"}
    for grad, bpp, key in ((False, 44, "fwd"), (True, 64, "fwd_bwd")):
        p = pr.clone().requires_grad_(grad)
        ts = []
        for i in range(iters + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(10):
                kf(p, tg)
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b) / 10)
        ms = float(np.median(ts))
        out[key] = {"ms": ms, "gbs": N * bpp / ms / 1e6, "frac": N * bpp / ms / 1e6 / pk["hbm_gbs"],
                    "pairs_per_s": N / ms * 1e3}
    return out


def bench_post_process(env, pk, iters=5):
    import ryolo_b200 as R
    dev = env.dev
    gen = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"workload": "post_process (lib/general.py:136-183) on 64 images x 100 000 candidate rows, clustered boxes",
           "l2": "205 MB (nc2) / 563 MB (nc16) of rows per call (> 126 MB L2) + 256 MB flush between iterations",
           "fp32_alu_peak_tflops": 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) / 1e6}
    for nc, conf, iou, tag in ((2, 0.001, 0.65, "val_nc2"), (2, 0.7, 0.2, "detect_nc2"), (16, 0.001, 0.65, "val_nc16")):
        pred = make_pred(64, 100000, nc, gen, dev)
        for _ in range(2):
            R.post_process_device(pred, conf, iou, mutate=False)
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            d, r, n = R.post_process_device(pred, conf, iou, mutate=False)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        # candidates that reach NMS per image (<= 5000): the reference's mask kernel evaluates K(K-1)/2 pairs of them
        sc = (pred[:, :, 6:] * pred[:, :, 5:6]).amax(2)
        K = (sc > conf).sum(1).clamp(max=5000).double()
        pairs = float((K * (K - 1) / 2).sum())
        out[tag] = {"conf_thres": conf, "iou_thres": iou, "nc": nc, "ms": ms, "rows_per_s": 64 * 100000 / ms * 1e3,
                    "frontend_bytes_per_row": (6 + nc) * 4,
                    "gbs_whole_call": 64 * 100000 * (6 + nc) * 4 / ms / 1e6,
                    "frac_hbm_whole_call": 64 * 100000 * (6 + nc) * 4 / ms / 1e6 / pk["hbm_gbs"],
                    "nms_candidate_pairs": pairs, "pairs_per_s_whole_call": pairs / ms * 1e3,
                    "survivors_per_img": float(n.float().mean())}
        del pred
    out["note"] = "whole-call numbers (score/filter + select/sort + pair mask + scan in one timed region); the per-kernel " \
                  "split and DRAM bytes are in profiles/r02_nonconv_*.  The pair stage is ALU-bound by construction."
    return out


def run_gpu(args):
    import torch.distributed as dist
    from ryolo_b200 import _lib as L
    env = Env()
    env.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    env.rank = rank = int(os.environ.get("RANK", "0"))
    env.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    env.dev = dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L.check(L.lib().ryolo_check_device(local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(n)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    env.timed = timed
    warm = max(args.warmup, 3)
    wl = WORKLOADS[args.workload]
    head = bench_train(env, wl, args.steps, warm, True, not args.no_graph)
    aux = {}
    if not args.no_aux:
        other = "train_v7" if args.workload == "train_v4" else "train_v4"
        o = bench_train(env, WORKLOADS[other], args.steps, warm, False, not args.no_graph)
        pk, pk_src = peaks()
        aux[other] = {"workload": WORKLOADS[other]["name"], "value": o["value"], "unit": UNIT,
                      "ms_per_step": o["ms_per_step"], "n_gpus": world, "per_gpu_batch": BS, "e2e": o["e2e"],
                      "fwd_loss_img_s": o["fwd_loss_img_s"], "gpu_launches": o["launches"],
                      "execution": o["execution"], "eager": o["eager"],
                      "allreduce_mb": o["grad_mb"] if world > 1 else 0,
                      "roofline": roofline_of(o, WORKLOADS[other], pk, pk_src)}
        if world == 1:
            aux["kfloss"] = bench_kfloss(env, pk)
            aux["post_process"] = bench_post_process(env, pk)
    if rank == 0:
        pk, pk_src = peaks()
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["name"], "per_gpu_batch": BS, "global_batch": BS * world,
                       "targets_per_img": PER_IMG,
                       "parallelism": f"dp{world}" + (" (bucketed NCCL all-reduce of the fp32 gradients, overlapped with "
                                                      "backward)" if world > 1 else ""),
                       "optimizer": "SGD lr .01 momentum .937 nesterov (train.py:156)",
                       "execution": head["execution"], "eager": head["eager"],
                       "fwd_loss_img_s": head["fwd_loss_img_s"], "fwd_loss_ms_per_step": head["fwd_loss_ms"],
                       "l2": "inputs (246 MB images + multi-GB activations per step) exceed the 126 MB L2"},
            "e2e": head["e2e"], "gpu_launches": head["launches"], "clocks": head["clocks"],
            "roofline": roofline_of(head, wl, pk, pk_src),
        }
        if world == 1 and not args.no_cpu:
            v, cores, kind, sample = cpu_train_arm(wl, 1, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
            if "kfloss" in aux:
                aux["kfloss"]["cpu_baseline"] = cpu_kfloss_arm()
            if "post_process" in aux:
                aux["post_process"]["cpu_baseline"] = cpu_post_process_arm(0.001, 0.65)
        if aux:
            line["aux"] = aux
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step only (no CUDA-graph capture)")
    ap.add_argument("--no-aux", action="store_true", help="headline workload only (no aux.train_v7 / kfloss / post_process)")
    ap.add_argument("--workload", default="train_v4", choices=["train_v4", "train_v7"],
                    help="headline workload; the other one is reported under aux")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args)


if __name__ == "__main__":
    main()
