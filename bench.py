#!/usr/bin/env python
"""bench.py -- BPR-MF training throughput on the BASELINE.json workload (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--shape ml-20m]

One "step" = one synchronous BPR-MF training step over one batch of ``--batch`` triples (u,i,j)
(the reference's zero_grad + calc_loss + backward + optimizer.step, AbstractRecommender.py:119-128).
Workload at N=1: BASELINE.json configs[1] -- MF + BPR, synthetic ML-20M shape (138,493 x 26,744,
20 M interactions, num_ng=4 -> 80 M triples/epoch), factors=64, fp32, SGD lr .01, reg .001/.001.
Metric: BPR user-item pairs (= training triples) per second, whole job.

Own arm (default):
  value      K steps timed with CUDA events around the persistent step-kernel launches, index planes and
             tables resident in HBM.  Steps walk the epoch's batches; launches cover one epoch's worth of
             steps at most (what MF.fit does), so K steps = ceil(K / steps_per_epoch) launches.
  e2e        the same metric through the public API with HOST batches: MF.fit_host_batches(pinned host index planes)
             -- every step's index arrays are copied H2D and every step's loss is read back D2H inside the timed
             region, the copy of batch s+1 overlapping the kernel of batch s; the blocking per-batch
             MF.train_step loop (the reference's loop shape) is reported beside it.
  roofline   algorithmic bytes (24*F+12 per triple) / event-timed launch duration vs the
             measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  oracle/torch_port.py (the reference's algorithm on PyTorch-CPU) on this host's cores.
Reference arm (--impl reference): the same port timed alone on the host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bpr_train_triples_per_sec"
UNIT = "triples/s"


# ------------------------------------------------------------------------------- helpers
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi sampler running beside the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.rows, self.proc, self.thr = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(dev)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, n = [], None, set(), 0
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, mxc = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = mxc
            if t0 - 0.05 <= ts <= t1 + 0.05:
                n += 1
                sm.append(clk)
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:                                      # region shorter than the sampling period
            sm = [float(x.split(",")[1]) for _, x in self.rows[-3:] if len(x.split(",")) > 2]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples_in_region": n}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def profiled_traffic(factors, batch):
    """dram bytes per STEP from the committed ncu capture (profiles/traffic.json), else None; the caller scales it to
    the steps of its average launch."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            if t.get("factors") == factors and t.get("batch") == batch:
                return t.get("dram_bytes_per_step")
        except Exception:  # noqa: BLE001
            pass
    return None


def build_workload(shape, device, num_ng, seed, sampler):
    """Synthetic interactions -> (data dict, triples int32 [T,3] on `device`).  sampler: 'cuda' | 'oracle'."""
    from daisyrec_b200.utils.synthetic import SHAPES, make_interactions
    U, I, nnz = SHAPES[shape]
    d = make_interactions(U, I, nnz, seed=seed, device=device)
    row_ptr_h = d["row_ptr"].cpu().numpy()
    if sampler == "cuda":
        from daisyrec_b200 import ops
        st = ops.mt19937_seed(seed)
        draws = ops.sampler_draw_mt19937(st, row_ptr_h, U, I, num_ng)
        js = ops.sampler_kth_complement(d["row_ptr"], d["col"], torch.from_numpy(draws).to(device), I)
        triples = ops.sampler_explode(d["coo_u"], d["coo_i"], js)
    else:
        from oracle import oracle as orc
        js = orc.sample_negatives(orc.mt_seed(seed), row_ptr_h, d["col"].cpu().numpy(), U, I, num_ng)
        triples = torch.from_numpy(orc.explode_triples(d["coo_u"].cpu().numpy(), d["coo_i"].cpu().numpy(), js))
    return d, triples


def time_cpu_port(P0, Q0, planes_cpu, batch, hyper, budget_s, max_steps, warmup=1):
    """Time oracle/torch_port.py steps of `batch` triples on the host cores within ~budget_s."""
    from oracle.torch_port import TorchMFBaseline
    m = TorchMFBaseline(P0, Q0, hyper["lr"], hyper["reg_1"], hyper["reg_2"], "sgd")
    n = planes_cpu[0].numel()
    nb = max(1, n // batch)

    def run(s):
        lo = (s % nb) * batch
        b = [p[lo:lo + batch].to(torch.int64) for p in planes_cpu]
        return m.step(*b)

    for s in range(warmup):
        run(s)
    t0 = time.perf_counter()
    done = 0
    while done < max_steps:
        run(warmup + done)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * batch / dt, done, dt


# ------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank, local, world = dist_env()
    if rank != 0:
        return 0
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    dev = "cuda" if torch.cuda.is_available() else "cpu"     # data generation only; the timed path is CPU
    from daisyrec_b200.utils.synthetic import SHAPES, init_tables
    U, I, _ = SHAPES[args.shape]
    d, triples = build_workload(args.shape, dev, args.num_ng, args.seed, "oracle")
    T = triples.shape[0]
    g = torch.Generator(); g.manual_seed(args.seed)
    perm = torch.randperm(T, generator=g)
    tr = triples.cpu()[perm]
    planes = [tr[:, k].contiguous() for k in range(3)]
    P0, Q0 = init_tables(U, I, args.factors, args.seed, "cpu")
    hyper = dict(lr=0.01, reg_1=0.001, reg_2=0.001)
    # calibrate: shrink the per-step sample so that K+W steps fit the time budget
    budget = args.ref_budget
    batch = args.batch
    tps1, _, dt1 = time_cpu_port(P0, Q0, planes, batch, hyper, 1e9, 1, warmup=1)
    per_step = batch / tps1
    total = per_step * (args.steps + args.warmup)
    if total > budget:
        batch = max(256, int(batch * budget / total) // 256 * 256)
    from oracle.torch_port import TorchMFBaseline
    m = TorchMFBaseline(P0, Q0, hyper["lr"], hyper["reg_1"], hyper["reg_2"], "sgd")
    nb = max(1, T // batch)

    def step(s):
        lo = (s % nb) * batch
        return m.step(*[p[lo:lo + batch].to(torch.int64) for p in planes])

    for s in range(args.warmup):
        step(s)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(args.warmup + s)
    dt = time.perf_counter() - t0
    value = args.steps * batch / dt
    sample = (f"{args.steps} steps x {batch} triples of the same epoch (own arm batch {args.batch}"
              f"{'' if batch == args.batch else ', shrunk to fit the time budget'}), dense fp32 autograd + SGD, "
              f"torch {torch.__version__} CPU, {cores} threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, d, T, batch),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, d, T, batch, extra=None):
    cfg = {"workload": f"MF+BPR synthetic {args.shape} shape ({d['user_num']}x{d['item_num']}, nnz={d['nnz']}, "
                       f"num_ng={args.num_ng} -> {T} triples/epoch), factors={args.factors}, SGD lr=0.01 reg=0.001/0.001",
           "batch_size": batch, "factors": args.factors, "triples_per_epoch": T, "optimizer": "sgd",
           "l2": "index planes (12 B/triple, all K steps) exceed L2 and are streamed once; the factor tables "
                 "(42 MB at F=64) are persistent model state reused by every step and stay L2-resident by design"}
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------- own arm
def run_own(args):
    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from daisyrec_b200.parallel import run_sharded_bench
        return run_sharded_bench(args, rank, local, world, dev)
    from daisyrec_b200 import ops
    from daisyrec_b200.model.MFRecommender import MF
    from daisyrec_b200.utils.synthetic import init_tables
    import logging

    d, triples = build_workload(args.shape, dev, args.num_ng, args.seed, "cuda")
    U, I, F, B = d["user_num"], d["item_num"], args.factors, args.batch
    T = triples.shape[0]
    g = torch.Generator(device=dev); g.manual_seed(args.seed)
    perm = torch.randperm(T, generator=g, device=dev)
    bu, bi, bj = ops.gather_triples(triples, perm)
    del perm
    spe = (T + B - 1) // B                                           # steps per epoch

    cfg = dict(gpu="", logger=logging.getLogger("bench"), lr=0.01, reg_1=0.001, reg_2=0.001, epochs=1, topk=50,
               user_num=U, item_num=I, factors=F, loss_type="BPR", optimizer="default", init_method="default",
               early_stop=False, progress=False)
    model = MF(cfg)
    P0, Q0 = init_tables(U, I, F, args.seed, dev)
    model.load_state_dict({"embed_user.weight": P0, "embed_item.weight": Q0})
    model._begin_fit("sgd")
    P, Q, ws, hp = model.embed_user.weight, model.embed_item.weight, model._ws, model._hp

    def run_steps(first, k, timed):
        """k steps starting at global step `first`, walking the epoch cyclically; one launch per epoch segment."""
        evs, launches, s = [], 0, first
        while k > 0:
            pos = s % spe
            seg = min(k, spe - pos)
            if timed:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            ops.mf_bpr_train_steps(P, Q, ws, bu, bi, bj, B, pos, seg, hp, check=False)
            if timed:
                e1.record()
                evs.append((e0, e1, seg, pos))
            launches += 1
            s += seg
            k -= seg
        return evs, launches

    clocks = ClockSampler(local)
    run_steps(0, args.warmup, False)
    torch.cuda.synchronize()
    t_region0 = time.time()
    evs, launches = run_steps(args.warmup, args.steps, True)
    torch.cuda.synchronize()
    t_region1 = time.time()
    ms = sum(e0.elapsed_time(e1) for e0, e1, _, _ in evs)
    done_triples = 0
    for _, _, seg, pos in evs:
        done_triples += min(T, (pos + seg) * B) - pos * B
    value = done_triples / ms * 1e3
    nan_check = ops.mf_bpr_loss(P, Q, ws, bu[:B], bi[:B], bj[:B], hp).item()
    if not np.isfinite(nan_check):
        raise RuntimeError("bench: loss became non-finite during the timed steps")

    # ---- end to end through the public API with HOST batches (pinned): every step's index arrays are copied
    #      H2D inside the timed region and every step's loss is read back D2H.
    #      e2e      = MF.fit_host_batches: the copy of batch s+1 overlaps the kernel of batch s (pipelined)
    #      e2e_sync = MF.train_step per batch with a blocking loss read, exactly the reference's loop shape
    ke = max(1, min(args.steps, args.e2e_steps, spe - 1))
    planes = [t[:ke * B].cpu().pin_memory() for t in (bu, bi, bj)]
    wk = min(3, ke)
    model.fit_host_batches(*[p_[:wk * B] for p_ in planes], B, wk)   # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rounds = max(1, args.steps // ke)
    for _ in range(rounds):
        host_losses = model.fit_host_batches(*planes, B, ke)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = e0.elapsed_time(e1)
    e2e_value = rounds * ke * B / e2e_ms * 1e3
    assert bool(torch.isfinite(host_losses).all())
    ks = min(ke, 16)
    host = [[p_[s * B:(s + 1) * B] for p_ in planes] for s in range(ks)]
    model.train_step(host[0])
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for s in range(ks):
        model.train_step(host[s])
    f1.record()
    torch.cuda.synchronize()
    t_e1 = time.time()
    e2e_sync_value = ks * B / f0.elapsed_time(f1) * 1e3
    clk = clocks.stop(t_region0, t_e1)

    # ---- CPU baseline (bounded sample) on this host's cores
    torch.set_num_threads(os.cpu_count() or 1)
    planes_cpu = [t[:min(T, 4 * B)].cpu() for t in (bu, bi, bj)]
    P0c, Q0c = init_tables(U, I, F, args.seed, "cpu")
    cpu_tps, cpu_steps, cpu_dt = time_cpu_port(P0c, Q0c, planes_cpu, B, dict(lr=0.01, reg_1=0.001, reg_2=0.001),
                                               args.cpu_budget, 8)
    peak, peak_src = measured_peaks()
    bytes_per_triple = 24 * F + 12
    avg_launch_ms = ms / len(evs)
    avg_launch_triples = done_triples / len(evs)
    achieved = avg_launch_triples * bytes_per_triple / (avg_launch_ms * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, d, T, B, {"parallelism": "single GPU", "steps_per_epoch": spe}),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 8,
                    "steps": rounds * ke,
                    "api": "MF.fit_host_batches(pinned host index planes): per step H2D of the batch + step kernel + "
                           "D2H of the loss; copy of batch s+1 overlaps the kernel of batch s",
                    "per_step_blocking": {"value": e2e_sync_value, "steps": ks,
                                          "api": "MF.train_step(host batch) with a blocking loss read per step"}},
            "gpu_launches": launches,
            "gpu_launches_note": "persistent cooperative kernel: one launch runs up to steps_per_epoch synchronous steps",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (lambda t_: None if t_ is None else t_ * (args.steps / len(evs)))(profiled_traffic(F, B)),
                         "traffic_note": "dram__bytes_read+write per step from profiles/r01b (ncu --set full) x steps per launch",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_triple": bytes_per_triple,
                         "kernel": "mf_bpr_steps_kernel", "avg_launch_ms": avg_launch_ms},
            "cpu_baseline": {"value": cpu_tps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{cpu_steps} steps x {B} triples of the same workload in {cpu_dt:.1f} s "
                                       f"(oracle/torch_port.py: dense fp32 autograd + SGD, torch {torch.__version__} CPU)"}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--factors", type=int, default=64)
    ap.add_argument("--shape", default="ml-20m")
    ap.add_argument("--num-ng", dest="num_ng", type=int, default=4)
    ap.add_argument("--seed", type=int, default=2022)
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=64)
    ap.add_argument("--cpu-budget", dest="cpu_budget", type=float, default=12.0)
    ap.add_argument("--ref-budget", dest="ref_budget", type=float, default=120.0)
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 8 if args.steps is None else args.steps
        args.warmup = 2 if args.warmup is None else args.warmup
        return run_reference(args)
    args.steps = 760 if args.steps is None else args.steps
    args.warmup = 76 if args.warmup is None else max(args.warmup, 3)
    return run_own(args)


if __name__ == "__main__":
    sys.exit(main())
