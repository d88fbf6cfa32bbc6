#!/usr/bin/env python
"""bench.py - headline benchmark of the SFD2 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--precision exact|mixed|fast|fp32]
    python bench.py --impl reference [...]        # the reference's CPU path (oracle port) on host cores

Metric (BASELINE.json): images/s of extract @1600x1200, top-4096 keypoints
(configs[1]); the 4096x4096 mutual-NN matcher (configs[2]) is reported beside it.

A "step" is one pass of the extract path over a batch of B synthetic 1600x1200
images.  `value` times the C-ABI device entry point with the batch already
resident in HBM; `e2e` times the reference-facing call
`extract_resnet_return(model, img_cpu, topK=4096, ...)` per image with pinned
HOST buffers (H2D of the image and D2H of keypoints/scores/descriptors inside the
timed region).  Inputs cycle through a pool larger than L2 and every step
rewrites ~GBs of activations, so nothing is served from a warm L2.

One JSON line on stdout (rank 0).  Under torchrun each rank owns its own batch
(weak scaling, no data-path collective); one NCCL all_gather collects
(x, y, score) + counts for the benchmark table outside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
WEIGHTS = os.path.join(REPO, "weights", "ressegnetv2_wapv2.npz")

H, W, TOPK, CONF = 1200, 1600, 4096, 0.001
GFLOP_PER_IMAGE = 831.114        # reference's dense count, SURVEY.md A.1 (415.557 GMAC)
GFLOP_CONV1A, GFLOP_STA = 6.636, 0.184   # the two layers that do not run in tc_conv_kernel
GFLOP_PER_PAIR = 4.295           # 2 * 4096 * 4096 * 128
METRIC = "images/sec extract@1600x1200 top-4096"


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling guide recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_extract_rate(n_images, warmup=1, flush_denormal=False):
    """images/s of the oracle port (same op sequence as the reference) on all host threads.
    flush_denormal=True is NOT what the reference does: the checkpoint's dead BatchNorm channels
     make its fp32 convolutions run on denormals, which Intel hosts execute
    through microcode assists; the flag shows what the same code does without that penalty."""
    import torch
    torch.set_flush_denormal(bool(flush_denormal))
    from oracle import sfd2_oracle as orc
    from sfd2_b200.synth import synth_image
    st = orc.load_state(WEIGHTS)
    imgs = [synth_image(s, H, W) for s in range(max(1, min(n_images, 4)))]
    for i in range(warmup):
        orc.extract(st, imgs[0], topK=TOPK, conf_th=CONF)
    t = []
    for i in range(n_images):
        t0 = time.perf_counter()
        orc.extract(st, imgs[i % len(imgs)], topK=TOPK, conf_th=CONF)
        t.append(time.perf_counter() - t0)
    return n_images / sum(t), min(t), torch.get_num_threads()


def cpu_match_rate(n_pairs):
    import torch
    from oracle import sfd2_oracle as orc
    from sfd2_b200.synth import synth_descriptors
    d0, d1 = synth_descriptors(0, 4096, 4096)
    a, b = torch.from_numpy(d0.T.copy())[None], torch.from_numpy(d1.T.copy())[None]
    orc.match_hloc(a, b)
    t0 = time.perf_counter()
    for _ in range(n_pairs):
        orc.match_hloc(a, b)
    return n_pairs / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path = the oracle port
    (the reference is pure Python and /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    steps, warm = args.steps, args.warmup
    rate, best, threads = cpu_extract_rate(steps, warmup=max(1, warm))
    pairs = cpu_match_rate(10)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1000.0 / rate, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "extract 1600x1200 top-4096, 1 image per step, oracle port of the reference "
                                   "PyTorch-CPU path (torch %s, %d threads)" % (torch.__version__, threads)},
            "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} images of 1600x1200 (+{max(1, warm)} warm-up), best {best:.3f} s"},
            "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "match": {"pairs_per_s": pairs, "unit": "4096x4096 pairs/s", "sample": "10 pairs, hloc NearestNeighbor port"},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from sfd2_b200 import Extractor, extract_resnet_return
    from sfd2_b200.matchers import match_dev, _ctx as match_ctx
    from sfd2_b200.synth import synth_image_u8, synth_descriptors

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, K, Wm = args.batch, args.steps, args.warmup
    pool_n = max(args.pool, B)
    # distinct synthetic images per rank (images[rank::world] of the C4-style list)
    seeds = [rank + world * i for i in range(pool_n)]
    pool_u8 = np.stack([synth_image_u8(s, H, W) for s in seeds])
    pool = torch.from_numpy(pool_u8).to(dev).float().div_(255.0).permute(0, 3, 1, 2).contiguous()  # [P,3,H,W] f32
    host_imgs = [torch.from_numpy((pool_u8[i].astype(np.float32) / np.float32(255)).transpose(2, 0, 1)[None].copy()).pin_memory()
                 for i in range(min(pool_n, 8))]

    ex = Extractor(WEIGHTS, use_stability=True, precision=args.precision, topk=TOPK, conf_th=CONF, device=dev)
    ctx = ex.model.ctx

    def step(i):
        j = (i * B) % pool_n
        idx = [(j + t) % pool_n for t in range(B)]
        batch = pool[idx] if idx != list(range(j, j + B)) else pool[j:j + B]
        return ex(batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        out = step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ctx.profile(True)
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        out = step(Wm + i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launch_count() - launches0
    # the same K steps again without per-launch events: this is the reported time
    barrier()
    e0.record()
    for i in range(K):
        out = step(Wm + i)
    e1.record()
    barrier()
    ms_clean = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the other tcgen05 precision modes, same device-resident workload, K steps each (reported beside the main number) ----
    other = {}
    if not args.no_other_modes:
        for prec in ("exact", "mixed", "fast"):
            if prec == args.precision:
                continue
            ex2 = Extractor(WEIGHTS, use_stability=True, precision=prec, topk=TOPK, conf_th=CONF, device=dev)
            for i in range(3):
                ex2(pool[(i % 2) * B:(i % 2) * B + B] if pool_n >= 2 * B else pool[:B])
            barrier()
            e0.record()
            for i in range(K):
                j = (i * B) % pool_n
                ex2(pool[j:j + B] if j + B <= pool_n else pool[:B])
            e1.record()
            barrier()
            other[prec] = e0.elapsed_time(e1)
            del ex2

    # ---- e2e: HOST buffers in, HOST features out, copies inside the timed region ----
    # (a) the batched C-ABI call sfd2_extract_host on B pinned host images per step (H2D of image i+1 overlaps
    #     the kernels of image i); (b) the reference-facing single-image call extract_resnet_return.
    model = ex.model
    host_batch = torch.cat(host_imgs[:min(B, len(host_imgs))] * ((B + len(host_imgs) - 1) // len(host_imgs)))[:B].pin_memory()
    for i in range(2):
        ex.extract_host(host_batch)
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        r = ex.extract_host(host_batch)
    e2e_s = time.perf_counter() - t0
    n_e2e = B * K
    n_single = max(4, min(B * K, 24))
    for i in range(2):
        extract_resnet_return(model, host_imgs[i % len(host_imgs)], topK=TOPK, conf_th=CONF, scales=[1.0])
    barrier()
    t0 = time.perf_counter()
    for i in range(n_single):
        r = extract_resnet_return(model, host_imgs[i % len(host_imgs)], topK=TOPK, conf_th=CONF, scales=[1.0])
    single_s = time.perf_counter() - t0

    # ---- matcher (configs[2]): 4096 x 4096 x 128 mutual NN, device-resident ----
    d0, d1 = synth_descriptors(rank, 4096, 4096)
    a, b = torch.from_numpy(d0).to(dev), torch.from_numpy(d1).to(dev)
    for _ in range(3):
        match_dev(a, b, precision=args.precision)
    mctx = match_ctx(local)
    mctx.profile(True); mctx.profile_read()
    n_pairs = 50
    barrier()
    e0.record()
    for _ in range(n_pairs):
        m0, s0 = match_dev(a, b, precision=args.precision)
    e1.record()
    barrier()
    match_ms = e0.elapsed_time(e1)
    mprof = mctx.profile_read(); mctx.profile(False)

    # ---- localizer pattern: one query (4096) against 50 db images (2000 descriptors each), grouped launch ----
    from sfd2_b200.matchers import match_one_to_many
    rngm = np.random.RandomState(100 + rank)
    dbs = rngm.randn(50 * 2000, 128).astype(np.float32)
    dbs /= np.linalg.norm(dbs, axis=1, keepdims=True)
    dbt = torch.from_numpy(dbs).to(dev)
    offs = np.arange(51, dtype=np.int32) * 2000
    for _ in range(2):
        match_one_to_many(a, dbt, offs, precision=args.precision)
    barrier()
    e0.record()
    for _ in range(5):
        match_one_to_many(a, dbt, offs, precision=args.precision)
    e1.record()
    barrier()
    o2m_ms = e0.elapsed_time(e1) / 5
    e0.record()
    for k in range(50):
        match_dev(a, dbt[k * 2000:(k + 1) * 2000], precision=args.precision)
    e1.record()
    barrier()
    loop_ms = e0.elapsed_time(e1)

    # ---- max over ranks, all-gather for the table ----
    t = torch.tensor([ms_clean, ms, e2e_s, match_ms, float(launches)] + [other.get(k, 0.0) for k in ("exact", "mixed", "fast")],
                     device=dev, dtype=torch.float64)
    gather_ms = 0.0
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        table = torch.cat([out["keypoints"], out["scores"][..., None]], -1).contiguous()   # [B, K, 3]
        allk = torch.empty(world * table.shape[0], TOPK, 3, device=dev)
        allc = torch.empty(world * B, dtype=torch.int32, device=dev)
        barrier()
        g0 = time.perf_counter()
        dist.all_gather_into_tensor(allk, table)
        dist.all_gather_into_tensor(allc, out["counts"])
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - g0) * 1e3
        ms_clean, ms, e2e_s, match_ms = [float(x) for x in tmax[:4]]
        other = {k: float(tmax[5 + i]) for i, k in enumerate(("exact", "mixed", "fast")) if k in other}
        launches = int(tsum[4].item())
        kpts_total = int(allc.sum().item())
    else:
        kpts_total = int(out["counts"].sum().item())

    if rank == 0:
        P = peaks()
        imgs = world * B * K
        value = imgs / (ms_clean / 1e3)
        tc = {k: v for k, v in prof.items() if k.startswith("tc_conv:") or k.startswith("conv_f32:")}
        conv_ms = sum(v[1] for v in tc.values())
        conv_launches = sum(v[0] for v in tc.values())
        alg_gflop = GFLOP_PER_IMAGE - GFLOP_CONV1A - GFLOP_STA
        n_img_prof = B * K
        achieved = alg_gflop * n_img_prof / conv_ms if conv_ms > 0 else 0.0       # GFLOP/ms = TFLOP/s
        traffic = None
        tp = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))       # written by tools/ncu_summary.py from the committed ncu --set full capture
            traffic = tj.get(args.precision, tj).get("tc_conv_bytes_per_launch")
        step_ms_prof = ms / K
        per_kernel = {}
        for k, (cnt, tot) in prof.items():
            name = k.split(":")[0]
            e = per_kernel.setdefault(name, [0, 0.0])
            e[0] += cnt; e[1] += tot
        mk = [v for k, v in mprof.items() if k.startswith("match")]
        match_kernel_ms = sum(v[1] for v in mk) / max(1, sum(v[0] for v in mk))
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_clean / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"exact": "f16x3->f32", "mixed": "f16x3->f32 (descriptor head f16x1)", "fast": "f16->f32", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": f"extract 1600x1200 top-4096, batch {B} images/step/GPU resident in HBM (f32 NCHW), "
                                   f"precision={args.precision}", "image": [H, W], "topk": TOPK, "batch_per_gpu": B,
                       "l2": f"inputs cycle through a {pool_n}-image pool ({pool_n * H * W * 12 / 1e6:.0f} MB > 126 MB L2); "
                             "every image rewrites >1 GB of activations", "parallelism": f"dp{world} (images sharded, no data-path collective)"},
            "e2e": {"value": world * n_e2e / e2e_s, "unit": "images/s", "h2d_bytes_per_step": B * H * W * 3 * 4,
                    "d2h_bytes_per_step": B * (TOPK * (2 + 1 + 128) * 4 + 4),
                    "note": f"sfd2_extract_host (C ABI) on {B} pinned host images per step, results to pinned host buffers",
                    "single_image_call": {"value": world * n_single / single_s, "unit": "images/s",
                                          "note": "extract_resnet_return(model, pinned host image): one synchronous call per image"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "tc_conv_kernel" if args.precision != "fp32" else "conv_f32_kernel",
                         "achieved": achieved, "peak": P["tflops"], "unit": "TFLOP/s", "frac": achieved / P["tflops"],
                         "peak_source": P["src"] + " (bf16 sustained)", "traffic": traffic,
                         "launches": conv_launches, "avg_launch_ms": conv_ms / max(1, conv_launches),
                         "algorithmic_gflop_per_image": alg_gflop, "share_of_step": conv_ms / ms if ms else None},
            "kernels_ms_per_image": {k: v[1] / n_img_prof for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])},
            "layers_ms_per_image": {k.split(":")[1]: v[1] / n_img_prof for k, v in tc.items()},
            "match": {"pairs_per_s": world * n_pairs / (match_ms / 1e3), "unit": "4096x4096x128 pairs/s",
                      "ms_per_pair": match_ms / n_pairs, "kernel_ms": match_kernel_ms,
                      "tflops": GFLOP_PER_PAIR / match_kernel_ms if match_kernel_ms else None,
                      "frac_of_peak": (GFLOP_PER_PAIR / match_kernel_ms / P["tflops_burst"]) if match_kernel_ms else None,
                      "one_to_many": {"workload": "4096 query x 50 db sets of 2000 descriptors", "grouped_ms": o2m_ms,
                                      "per_pair_loop_ms": loop_ms, "pairs_per_s_grouped": 50 / (o2m_ms / 1e3)}},
            "table": {"keypoints_last_step": kpts_total, "allgather_ms": gather_ms},
            "other_modes": {k: {"value": world * B * K / (v / 1e3), "unit": "images/s", "note": "device-resident, same workload"}
                            for k, v in other.items()},
        }
        if world == 1 and not args.no_cpu:
            rate, best, threads = cpu_extract_rate(args.cpu_images)
            line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": f"{args.cpu_images} images of 1600x1200 top-4096 through the oracle port "
                                              f"(1 warm-up), best {best:.3f} s/image"}
            r2, b2, _ = cpu_extract_rate(3, flush_denormal=True)
            line["cpu_baseline_flush_denormal"] = {"value": r2, "unit": "images/s", "cores": threads, "kind": "port",
                                                   "sample": f"3 images, torch.set_flush_denormal(True) (not the reference's "
                                                             f"setting), best {b2:.3f} s/image"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--pool", type=int, default=16)
    # mixed = the 3-pass fp16 split (fp32-grade) on everything that feeds the heat-map, single-pass fp16 on the descriptor
    # head: keypoints and scores identical to `exact`, descriptors within the north-star 1e-3 (tests/test_gpu_parity.py)
    ap.add_argument("--precision", default="mixed", choices=["exact", "mixed", "fast", "fp32"])
    ap.add_argument("--no-other-modes", action="store_true", help="skip the short exact / fast timings reported beside the main one")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-images", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
