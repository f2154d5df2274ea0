#!/usr/bin/env python
"""bench.py — rays/s of the K-sample uncertainty render (BASELINE.json configs[1]: africa.txt architecture,
512x512 image, N=128 samples, K=32 latent samples, mean/variance/depth) on N B200s, one process per GPU.

    python bench.py --gpus N --steps K --warmup W                    # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on the host CPU

A "step" renders one full synthetic image per GPU (weak scaling: rays shard with no collective).
`value` times the device-resident path (rays in HBM -> K-field maps + mean/std/depth in HBM); `e2e` times the public
`render_rays` call with host-pinned rays in and the K fields + statistics copied back out.  `roofline` is the
dominant kernel (the tcgen05 network stage) against the measured dense bf16 peak; `cpu_baseline` is the oracle port
of the reference (faithful K-fold materialisation) on a bounded ray sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_POINT = 4708864          # 2 354 432 MAC per network evaluation of one point
H = W_IMG = 512
FOCAL = 443.4
NEAR, FAR = 1.2, 8.0
N_SAMPLES = 128
WORKLOAD = ("africa.txt full-image 512x512 uncertainty render (mean/variance/depth), N=128 samples, K=32 latent samples, "
            "W=512 D=8, random-init weights, one image per GPU per step")


def image_rays(h, w, focal, seed_pose: int):
    """get_rays (run_nerf_helpers.py:288-297) for c2w = I (rank 0) or a small yaw (other ranks), packed (B,11)."""
    from oracle import cfnerf_oracle as O
    c2w = torch.eye(4)[:3].clone()
    if seed_pose:
        th = 0.05 * seed_pose
        c2w[0, 0], c2w[0, 2], c2w[2, 0], c2w[2, 2] = torch.cos(torch.tensor(th)), torch.sin(torch.tensor(th)), \
            -torch.sin(torch.tensor(th)), torch.cos(torch.tensor(th))
    o, d = O.get_rays(h, w, focal, c2w)
    return O.pack_ray_batch(o, d, NEAR, FAR)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def k1_dram_traffic(points_per_launch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the K1 launch from the committed `ncu --set full` capture
    (profiles/r01_prof_k1_summary.csv, same 4 194 304-point launch as the bench); None when it does not apply."""
    path = os.path.join(ROOT, "profiles", "r01_prof_k1_summary.csv")
    if not os.path.isfile(path) or points_per_launch != 4194304:
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(path):
        f = line.strip().split(",")
        if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[2]) * mult.get(f[1], 1.0)
    return tot or None


def tgemm_dram_traffic(precision: str, points: int):
    """DRAM read + write of one trunk-layer forward GEMM of the training step from the committed ncu capture of that
    precision (first kernel block of the summary); None when there is none for this shape."""
    path = os.path.join(ROOT, "profiles", "r01_prof_tgemm_fwd_summary.csv" if precision == "tf32" else "r01_prof_tgemm_bf16_summary.csv")
    if not os.path.isfile(path) or points != 524288:
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, blocks = 0.0, 0
    for line in open(path):
        if line.startswith("##"):
            blocks += 1
            if blocks > 1:
                break
        f = line.strip().split(",")
        if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(f[2]) * mult.get(f[1], 1.0)
    return tot or None


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rays_per_s(n_rays: int, reps: int, threads: int | None = None):
    """The oracle port of the reference path (K-fold materialisation as in models.py:210-217) on the host."""
    from oracle import cfnerf_oracle as O
    torch.set_num_threads(threads or host_threads())   # torchrun exports OMP_NUM_THREADS=1: use every host core
    cfg = O.CfnConfig()
    p = O.make_params(cfg, 0, "default")
    sa, sr = O.make_latents(cfg, 0)
    ea, er = O.test_latents(sa, sr)
    rays = image_rays(H, W_IMG, FOCAL, 0)[:: (H * W_IMG) // n_rays][:n_rays].contiguous()
    times = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            for i in range(0, n_rays, 512):  # netchunk = 65536 points (main:604)
                O.render_rays(p, cfg, rays[i:i + 512], ea, er, False, faithful=True)
            times.append(time.perf_counter() - t0)
    return n_rays / min(times), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_rays = args.cpu_rays
    times = []
    torch.set_num_threads(host_threads())   # torchrun exports OMP_NUM_THREADS=1: the reference gets every host core
    from oracle import cfnerf_oracle as O
    cfg = O.CfnConfig()
    p = O.make_params(cfg, 0, "default")
    sa, sr = O.make_latents(cfg, 0)
    ea, er = O.test_latents(sa, sr)
    rays = image_rays(H, W_IMG, FOCAL, 0)[:: (H * W_IMG) // n_rays][:n_rays].contiguous()
    with torch.no_grad():
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            for i in range(0, n_rays, 512):
                out = O.render_rays(p, cfg, rays[i:i + 512], ea, er, False, faithful=True)
                O.k_reduce(out["rgb_map"], out["depth_map"], cfg.K)
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    tot = sum(times)
    v = n_rays * len(times) / tot
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": "rays/sec (K-sample uncertainty render)", "value": v, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": H * W_IMG,
                   "reference_sample": f"each step times {n_rays} evenly spaced rays of the image on the host CPU"},
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": f"{n_rays} rays of the 512x512 image per step, oracle port of the reference "
                                   "(faithful K-fold conditioning), torch CPU fp32"},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp32"],
                    help="operand type of the tensor-core GEMM chain (bf16 and fp16 run at the same rate)")
    ap.add_argument("--chunk", type=int, default=32768, help="rays per render_rays call (bounds the flow-parameter buffer)")
    ap.add_argument("--cpu-rays", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE.json configs[2])")
    ap.add_argument("--train-rays", type=int, default=4096, help="rays per GPU per optimisation step")
    ap.add_argument("--train-precision", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="GEMM engine of the training leg: bf16 storage + kind::f16 (BASELINE configs[2] wording), "
                         "tf32 over fp32 storage, or the fp32 CUDA-core check engine")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist

    import cfnerf_b200 as cf
    from oracle import cfnerf_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback); use --impl reference for the host baseline")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = O.CfnConfig()
    params = O.make_params(cfg, 0, "default")          # random-init weights of the africa.txt architecture
    sa, sr = O.make_latents(cfg, 0)
    net = cf.NeRFFlowsParams.from_oracle_params(cfg, params, sa, sr).to(dev)
    eng = cf.engine_for(net, dev, args.precision)
    rays_host = image_rays(H, W_IMG, FOCAL, rank).pin_memory()
    rays_dev = rays_host.to(dev)
    B = rays_dev.shape[0]
    chunk = args.chunk
    ea, er = cf.test_latents(net, dev)
    t_vals = cf.reference_t_schedule(N_SAMPLES, dev)
    n_chunks = (B + chunk - 1) // chunk
    k1_events = []

    def render_device(rays, record_k1=False):
        """device-resident step: z schedule -> network (K1) -> flows+compositing (K2); 3 launches per chunk."""
        outs = []
        for i in range(0, B, chunk):
            r = rays[i:i + chunk]
            z = eng.zvals(r, t_vals, None, False)
            if record_k1:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            fp = eng.network(r.shape[0], N_SAMPLES, rays=r, z_vals=z)
            if record_k1:
                e1.record()
                k1_events.append((e0, e1))
            o = eng.flow_composite(fp, z, r[:, 3:6], 11, ea, er, False, want_kstats=True)
            outs.append(o)
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng.pack()
    for _ in range(max(args.warmup, 3)):
        render_device(rays_dev)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        render_device(rays_dev, record_k1=True)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    k1_ms = sum(a.elapsed_time(b) for a, b in k1_events)

    # ---- end to end through the public API: pinned host rays in, K fields + statistics out ----
    out_host = None

    def render_e2e():
        nonlocal out_host
        res = []
        for i in range(0, B, chunk):
            r = rays_host[i:i + chunk].to(dev, non_blocking=True)
            o = cf.render_rays(r, net, None, N_SAMPLES, False, False, K_samples=cfg.K, precision=args.precision,
                               want_kstats=True)
            res.append(o)
        cat = {k: torch.cat([o[k] for o in res], 0) for k in ("rgb_map", "disp_map", "depth_map", "kstats")}
        if out_host is None:
            out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in cat.items()}
        for k, v in cat.items():
            out_host[k].copy_(v, non_blocking=True)
        torch.cuda.synchronize()
        return cat

    for _ in range(2):
        render_e2e()
    barrier()
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(args.steps):
        cat = render_e2e()
    ee1.record()
    barrier()
    e2e_ms = ee0.elapsed_time(ee1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = rays_host.numel() * 4
    d2h = sum(v.numel() * 4 for v in cat.values())

    # ---- the other half of the metric: one optimisation step (BASELINE.json configs[2]) ----
    # render (train mode, saved activations) -> K-mean + KDE-NLL + 0.01 * entropy -> backward (flow/composite backward,
    # dgrad chain, split-K wgrad) -> one all-reduce of the flat gradient bucket (N > 1) -> Adam.  Weak scaling like the
    # render leg: --train-rays rays per GPU.  Runs last (it moves the weights).
    train_ms = 0.0
    if not args.no_train:
        from cfnerf_b200 import dist as D
        # the repo's own trainer step: a straight chain of C-ABI calls (no autograd graph), flat gradient buffer =
        # all-reduce bucket, fused Adam
        trainer = D.FusedTrainStep(net, lr=5e-4, precision=args.train_precision)
        gt = torch.Generator().manual_seed(100 + rank)
        t_rays = rays_dev[torch.randperm(B, generator=gt)[:args.train_rays].to(dev)].contiguous()
        t_target = torch.rand(t_rays.shape[0], 3, generator=gt).to(dev)
        torch.manual_seed(100 + rank)
        for _ in range(3):
            trainer.step(t_rays, t_target)
        barrier()
        te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        te0.record()
        for _ in range(args.steps):
            t_out = trainer.step(t_rays, t_target)
        te1.record()
        barrier()
        train_ms = te0.elapsed_time(te1)

    # roofline of the training GEMMs (HBM-bound by their activation traffic): one trunk-layer forward GEMM on the step's
    # shape (points x W x W, bias + ReLU + ReLU bit mask) timed alone with CUDA events; algorithmic bytes = X in + Y out
    tg_ms = 0.0
    if not args.no_train and args.train_precision in ("tf32", "bf16") and rank == 0:
        Mp = t_rays.shape[0] * N_SAMPLES
        bf = args.train_precision == "bf16"
        dt = torch.bfloat16 if bf else torch.float32
        Xg = torch.randn(Mp, cfg.W, device=dev).to(dt)
        Wg = (torch.randn(cfg.W, cfg.W, device=dev) * 0.05).to(dt)
        bg = torch.randn(cfg.W, device=dev)
        Yg = torch.empty(Mp, cfg.W, device=dev, dtype=dt)
        mb = torch.zeros(Mp, (cfg.W + 31) // 32, dtype=torch.int32, device=dev)

        def one_gemm():
            if bf:
                cf.gemm_bf16(Xg, Wg.t(), bias=bg, epilogue="relu", mask_out=mb, out=Yg)
            else:
                cf.gemm(Xg, Wg.t(), engine="tf32", bias=bg, epilogue="relu", out=Yg, round_out=True)

        for _ in range(3):
            one_gemm()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(10):
            one_gemm()
        g1.record()
        torch.cuda.synchronize()
        tg_ms = g0.elapsed_time(g1) / 10
        del Xg, Yg, mb
    barrier()

    t = torch.tensor([ms, e2e_ms, k1_ms, train_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, k1_ms, train_ms = [float(x) for x in t.cpu()]

    if rank == 0:
        peaks, which = measured_peaks()
        rays_total = B * world * args.steps
        value = rays_total / (ms * 1e-3)
        e2e = rays_total / (e2e_ms * 1e-3)
        pts_per_launch = min(chunk, B) * N_SAMPLES
        n_k1 = len(k1_events)
        k1_avg_s = (k1_ms * 1e-3) / n_k1
        achieved = (B * N_SAMPLES * FLOP_PER_POINT / n_chunks) / k1_avg_s / 1e12
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        line = {
            "metric": "rays/sec (K-sample uncertainty render)", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": B, "chunk_rays": chunk,
                       "l2": "no flush needed: each step streams a 9.7 GB flow-parameter buffer (>> 126 MB L2)"},
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 3 * n_chunks * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": k1_dram_traffic(pts_per_launch),
                         "traffic_note": "DRAM bytes per K1 launch from profiles/r01_prof_k1_summary.csv (ncu --set full); "
                                         "algorithmic HBM bytes = 288 B/point of flow parameters written = 1.208e9",
                         "kernel": "mlp_tc_kernel (network stage K1)",
                         "peak_source": f"{which} bf16_tflops_sustained",
                         "k1_share_of_step": k1_ms / ms, "points_per_launch": pts_per_launch},
            "clocks": sampler.summary(),
        }
        if not args.no_train:
            n_t = t_rays.shape[0]
            step_s = train_ms * 1e-3 / args.steps
            line["train_step"] = {
                "value": n_t * world / step_s, "unit": "rays/s", "ms_per_step": step_s * 1e3, "rays_per_gpu": n_t,
                "precision": args.train_precision, "loss": float(t_out["loss"]),
                "achieved_tflops_per_gpu": n_t * N_SAMPLES * FLOP_PER_POINT * 3 / step_s / 1e12,
                "note": "forward + loss + backward + all-reduce + Adam through cfnerf_b200.dist.FusedTrainStep; GEMMs = TMA-fed tcgen05: kind::f16 over "
                        "bf16-stored activations / gradients (bf16), kind::tf32 over fp32 storage (tf32), or CUDA-core fp32 "
                        "FMA (fp32); fp32 accumulation, master weights and weight gradients throughout; flops = 3 x forward"}
            if tg_ms > 0:
                Mp = n_t * N_SAMPLES
                bf = args.train_precision == "bf16"
                esz = 2 if bf else 4
                gb = 2.0 * Mp * cfg.W * esz / 1e9          # activations in + out (the weight matrix is L2-resident)
                tf = 2.0 * Mp * cfg.W * cfg.W / (tg_ms * 1e-3) / 1e12
                hbm = peaks.get("hbm_gbs", 6545.9)
                tpeak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
                # arithmetic intensity W / esz FLOP per byte: 256 (bf16 storage, above the 209 FLOP/B ridge of this part:
                # tensor-bound) or 128 (fp32 storage: HBM-bound)
                roof = ({"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                         "hbm_gbs": gb / (tg_ms * 1e-3), "hbm_frac": gb / (tg_ms * 1e-3) / hbm} if bf else
                        {"bound": "hbm", "achieved": gb / (tg_ms * 1e-3), "peak": hbm, "unit": "GB/s",
                         "frac": gb / (tg_ms * 1e-3) / hbm, "tflops": tf})
                roof.update({
                    "traffic": tgemm_dram_traffic(args.train_precision, Mp),
                    "traffic_note": "DRAM read + write of this GEMM from the committed ncu --set full capture "
                                    "(profiles/r01_prof_tgemm_*_summary.csv) at 524288 points; algorithmic bytes = "
                                    f"2 x points x 512 x {esz} = {2 * 524288 * 512 * esz:.4g}",
                    "kernel": "tgemm_kernel (one trunk-layer forward GEMM of the training step: bias + ReLU + bit mask)",
                    "ms": tg_ms})
                line["train_step"]["roofline"] = roof
        if not args.no_cpu_baseline:
            v, cores = cpu_reference_rays_per_s(args.cpu_rays, 2)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_rays} rays of the same image, oracle port of the reference "
                                              "(faithful K-fold conditioning), torch CPU fp32, best of 2"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
