"""bench.py -- SAC grad-steps/sec (batch 256, 64x64 depth) on N B200s.

  python bench.py --gpus N --steps K --warmup W             (N>1: launched under torchrun)
  python bench.py --impl reference ...                      (CPU arm: the oracle restatement of
                                                             SB2.10.1/TF1.14's SAC step, all host cores)

A "step" = one SAC minibatch gradient step (replay sample -> VecNormalize -> 3 CNN fwd, 2 CNN bwd,
heads, losses -> [all-reduce] -> 3x Adam -> Polyak) at batch 256 per GPU on synthetic 64x64x2 depth
observations (BASELINE.json configs[1]).  `value` times K steps with the replay already resident in
HBM (CUDA events, max over ranks); `e2e` times the same step through the C-ABI parity entry point
with HOST (pinned) batches, i.e. host->device copies of the batch and device->host read of the
losses inside the timed region.  Weak scaling: every rank processes its own 256-sample minibatch
and one NCCL all-reduce averages the gradients, so N ranks = one step on a global batch of N*256;
value = N x synchronous steps/s, in 256-sample step equivalents.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
FLOP_PER_STEP_B256 = 2 * 256 * 18_923_328          # 9.689 GFLOP
LR = 3e-4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["bf16_tflops"], d["hbm_gbs"], "measured (MEASURED_PEAKS.json, burst)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed (-lms 100) DURING the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.proc = dev, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            self.rows = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.count(",") >= 5]

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def load_case():
    from oracle import sac_ref as R  # only used by the cpu_baseline / reference legs
    vn = dict(np.load(os.path.join(GOLD, "vecnorm_sac_depth.npz")))
    raw = dict(np.load(os.path.join(GOLD, "sac_depth_params.npz")))
    cfg = R.SACConfig(obs_shape=(64, 64, 2))
    params = {n: raw[n] for n, _ in R.param_specs(cfg)}
    return cfg, params, vn


def cpu_reference_steps(batch, seconds, threads=None):
    """Times the CPU restatement of the SB2 SAC step (oracle/sac_ref.py, PyTorch-CPU fp32).

    The thread count is calibrated first (one timed step per candidate): on many-core hosts the
    small per-layer ops of this graph run far slower with every hardware thread than with a
    moderate pool, and TF1's own intra-op pool would be tuned the same way.  Returns
    (steps/s, steps, seconds, threads_used)."""
    import torch
    from oracle import sac_ref as R
    from b200grasp import synth
    cfg, params, vn = load_case()
    raw = synth.make_transitions(batch, vn["obs_mean"], vn["obs_var"])
    norm = dict(obs=R.normalize_obs(raw["obs"], vn["obs_mean"], vn["obs_var"]),
                next_obs=R.normalize_obs(raw["next_obs"], vn["obs_mean"], vn["obs_var"]),
                act=raw["act"], rew=R.normalize_reward(raw["rew"], float(vn["ret_var"])), done=raw["done"])
    eps = synth.make_eps(batch)
    p, opt = params, R.OptState.zeros(params)
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    if threads is None:
        cands = sorted({c for c in (avail, 64, 32, 16, 8) if c <= avail}, reverse=True)
        best, best_t = cands[-1], float("inf")
        for c in reversed(cands):                  # small pools first: they bound the calibration time
            torch.set_num_threads(c)
            R.sac_step(p, opt, norm, eps, LR, cfg, torch.float32)
            t0 = time.perf_counter()
            R.sac_step(p, opt, norm, eps, LR, cfg, torch.float32)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
            if dt > 4 * best_t:
                break
        threads = best
    torch.set_num_threads(threads)
    _, _, p, opt = R.sac_step(p, opt, norm, eps, LR, cfg, torch.float32)      # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        _, _, p, opt = R.sac_step(p, opt, norm, eps, LR, cfg, torch.float32)
        n += 1
        el = time.perf_counter() - t0
        if el >= seconds or n >= 400:
            break
    return n / el, n, el, threads


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path cannot run here
    (stable-baselines 2.10.1 / TF 1.14 are not installable), so this arm times
    the oracle port on all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, n, el, cores = cpu_reference_steps(256, min(60.0, 3.0 * max(1, args.steps)))
    line = {
        "impl": "reference", "metric": "SAC grad-steps/sec (batch 256, 64x64 depth)", "value": rate, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SAC depth CNN (config/gripper_grasp.yaml), batch 256, 64x64x2 obs, trained-weight init",
                   "note": "CPU restatement of SB2.10.1/TF1.14 SAC step (oracle/sac_ref.py, PyTorch-CPU fp32), not TF itself"},
        "cpu_baseline": {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n} full B=256 gradient steps in {el:.1f}s; torch intra-op threads calibrated to {cores} of {os.cpu_count()}"},
        "e2e": {"value": rate, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--replay-filled", type=int, default=8192, help="transitions resident in HBM (2 x 32 KiB each: 512 MiB > L2)")
    ap.add_argument("--buffer-size", type=int, default=1_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"],
                    help="bf16x3 = tcgen05 BF16 hi/lo split, the mode that passes the 1e-4 parity tests (default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import b200grasp
    from b200grasp import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(b200grasp.Learner.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    vn = dict(np.load(os.path.join(GOLD, "vecnorm_sac_depth.npz")))
    raw_params = dict(np.load(os.path.join(GOLD, "sac_depth_params.npz")))
    B = args.batch
    prec = {"fp32": 0, "bf16x3": 1, "bf16": 2}[args.precision]
    L = b200grasp.Learner((64, 64, 2), n_act=5, batch_size=B, buffer_size=args.buffer_size, seed=1234, device=local,
                          rank=rank, nranks=world, nccl_id=nccl_id, precision=prec)
    L.load_parameters(raw_params)      # identical replicas on every rank
    L.set_norm_stats(vn["obs_mean"], vn["obs_var"], float(vn["ret_var"]), float(vn["clip_obs"]), float(vn["clip_reward"]),
                     float(vn["epsilon"]))
    # replay shard of this rank (different data per rank), resident in HBM before the timed region
    chunk = 2048
    for i in range(0, args.replay_filled, chunk):
        tr = synth.make_transitions(min(chunk, args.replay_filled - i), vn["obs_mean"], vn["obs_var"], seed=synth.DATA_SEED + 1000 * rank + i)
        L.replay_add(tr["obs"], tr["act"], tr["rew"], tr["next_obs"], tr["done"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K graph replays between CUDA events on the learner's stream
    L.step(args.warmup, lr=LR)
    barrier()
    with ClockSampler(local) as clk:
        time.sleep(0.25)                       # let the sampler stream before the timed region starts
        L.step(args.steps, lr=LR)              # the timed region: exactly K steps, CUDA events on the learner's stream
        ms = L.last_step_ms()
        t_end = time.time() + 1.0              # short regions: repeat the same K-step region so that several clock
        while time.time() < t_end:             # samples fall under load; the fastest K-step pass is reported
            L.step(args.steps, lr=LR)
            ms = min(ms, L.last_step_ms())
    barrier()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    sync_steps_per_s = args.steps / (ms * 1e-3)
    value = world * sync_steps_per_s

    # ---- end to end through the C ABI with pinned HOST batches (H2D of the batch + D2H of the losses per step)
    tr = synth.make_transitions(B, vn["obs_mean"], vn["obs_var"], seed=77 + rank)
    eps = synth.make_eps(B, seed=78 + rank)
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() for k, v in dict(tr, eps=eps).items()}
    h2d = sum(v.nbytes for v in pin.values())
    d2h = 7 * B * 4 + B * 5 * 4 + 16 * 4 + 64 + 8
    e2e_steps = max(10, min(args.steps, 200))
    for _ in range(3):
        L.step_host_pipelined(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"], pin["eps"], lr=LR)
    L.pipeline_flush()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):        # every step: H2D of its batch (pinned) + D2H of a step's losses; copy k overlaps compute k-1
        L.step_host_pipelined(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"], pin["eps"], lr=LR)
    last = L.pipeline_flush()
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    assert np.isfinite(last["qf1_loss"])
    t = torch.tensor([el], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e = world * e2e_steps / float(t.item())
    # un-pipelined variant of the same call (b2g_sac_step_explicit: copy, step, read back, return)
    t0 = time.perf_counter()
    for _ in range(max(10, e2e_steps // 4)):
        L.step_explicit(pin["obs"], pin["act"], pin["rew"], pin["next_obs"], pin["done"], pin["eps"], lr=LR)
    e2e_sync = world * max(10, e2e_steps // 4) / (time.perf_counter() - t0)
    # the learn() loop of the SB-shaped front end: one new transition enters the device replay per gradient step
    one = {k: v[:1] for k, v in pin.items() if k != "eps"}
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        L.replay_add(one["obs"], one["act"], one["rew"], one["next_obs"], one["done"])
        L.step(1, lr=LR)
    e2e_learn = world * e2e_steps / (time.perf_counter() - t0)

    # ---- dominant-kernel roofline: per-launch device time of every group of ONE step (CUDA events on
    # the learner's stream between launches), on rank 0
    line = None
    prof = None
    for _ in range(3):              # every rank: the step contains the all-reduce
        prof = L.profile_step(lr=LR)
    if rank == 0:
        gemm_groups = {k: v for k, v in prof.items() if k.startswith("conv") or k.startswith("cnn_") or k.startswith("fc1_") or k.startswith("heads_fc0")
                       or k in ("heads_wgrad", "heads_dgrad")}
        gemm_ms = sum(gemm_groups.values())
        peak_tf, peak_hbm, peak_src = peaks()
        flops = FLOP_PER_STEP_B256 * B / 256
        achieved = flops / (gemm_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.exists(tpath):          # dram read+write per gg_tc_kernel launch from the committed ncu --set full capture
            traffic = json.load(open(tpath)).get("gg_tc_kernel_dram_bytes_per_launch")
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": traffic, "traffic_unit": "bytes per gg_tc_kernel launch (ncu dram read+write averaged over the step's tensor-engine launches, profiles/ncu_launches_r1.csv)", "kernel": "gg_tc_kernel (tcgen05 gather-GEMM, cp.async-fed BF16 hi/lo planes: convs, cnn_fc1 and head fc0 layers, fwd/wgrad/dgrad; 11 launches per step)" if prec else "gg_simt_kernel (fp32 FFMA engine)",
                    "peak_source": peak_src, "launch_ms": gemm_ms, "launches": len(gemm_groups),
                    "step_share": gemm_ms / sum(prof.values()), "per_group_ms": {k: round(v, 4) for k, v in prof.items()},
                    "whole_step_frac": value / world * flops / 1e12 / peak_tf}
        cpu = None
        if not args.no_cpu_baseline:
            rate, n, cel, cores = cpu_reference_steps(B, args.cpu_seconds)
            cpu = {"value": rate, "unit": "steps/s", "cores": cores, "kind": "port",
                   "sample": f"{n} full B={B} gradient steps in {cel:.1f}s (oracle/sac_ref.py, PyTorch-CPU fp32)"}
        line = {
            "metric": "SAC grad-steps/sec (batch 256, 64x64 depth)", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (fp32-faithful split, f32 accumulate)", "bf16": "bf16"}[args.precision], "data": "synthetic",
            "config": {"workload": "SAC depth CNN (config/gripper_grasp.yaml), batch 256/GPU, 64x64x2 obs, 1M-slot replay",
                       "global_batch": B * world, "replay_capacity": args.buffer_size, "replay_filled": args.replay_filled,
                       "l2": "replay working set 512 MiB > 126 MB L2; minibatch indices are random per step",
                       "precision": {"fp32": "fp32 FFMA (B2G_PREC_FP32_SIMT)", "bf16x3": "tcgen05 BF16 hi/lo split x3, fp32 TMEM accumulate (B2G_PREC_BF16X3; passes 1e-4 parity)", "bf16": "tcgen05 single-pass BF16 (fast mode, ~5e-4 on Q)"}[args.precision], "parallelism": f"dp{world}",
                       "sync_steps_per_s": sync_steps_per_s},
            "clocks": clk.summary(),
            "e2e": {"value": e2e, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16 * 4 + 8 + 64, "steps": e2e_steps,
                    "api": "b2g_sac_step_host_pipelined: full 256-sample batch from pinned host memory every step, losses read back every step (one step late)",
                    "unpipelined_steps_per_s": e2e_sync,
                    "learn_loop_steps_per_s": e2e_learn, "learn_loop_h2d_bytes_per_step": int(sum(v.nbytes for v in one.values()))},
            "gpu_launches": L.launches_per_step() * args.steps,
            "roofline": roofline, "cpu_baseline": cpu,
        }
    L.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


if __name__ == "__main__":
    main()
