#!/usr/bin/env python
"""bench.py -- collision checks/s (and plans/s) of the Prius planner hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the CPU arm)

Headline workload = BASELINE.json configs[2], the configuration the metric's
"collision checks/sec at 1/2/4/8 B200" is quoted on: 2^24 oriented-footprint poses per GPU
against a 4096 x 4096 synthetic grid at 0.1 m (weak scaling: every rank checks its own
2^24 poses; flags gathered with one NCCL all-gather inside the timed region).  A "step" is
one pass of the batch through pp_collide_batch_dev.  The single-query / batched-plan
figures (configs[1] and [4]) are reported in the same JSON line under "plans".

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POSES = 1 << 24
GRID = 4096
RES = 0.1
ALG_BYTES_ORIENTED = 12 + 1 + 48 * 17   # pose + flag + one byte per lattice sample = 829 B
ALG_BYTES_AABB = 12 + 1 + 47 * 15       # 718 B
# ncu --set full, collide_oriented_kernel, 2^22 poses: dram__bytes_read.sum + dram__bytes_write.sum
DRAM_BYTES_PER_POSE_NCU = (51.002880e6 + 39.424e3) / (1 << 22)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs: one `nvidia-smi -lms 20`
    child streams timestamped rows; only rows stamped inside [start(), stop()] are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.t0 = self.t1 = None
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        except Exception:
            self.proc = None

    def _run(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def start(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def stop(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def close(self):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def summary(self):
        import datetime
        inside, allrows = [], []
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
                float(r[1])
            except Exception:
                continue
            allrows.append(r)
            if self.t0 and self.t1 and self.t0 <= ts <= self.t1 + datetime.timedelta(milliseconds=30):
                inside.append(r)
        use = inside if inside else allrows[-3:]
        sm = sorted(float(r[1]) for r in use)
        mx = [float(r[2]) for r in use]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in use:
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(r[3]) for r in use if r[3].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(inside), "power_w_max": max(pw) if pw else None}


def bind_to_gpu_cpus(local):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end leg are first-touched on the GPU's NUMA node (8 ranks pulling 52 GB/s each through one node's
    memory controller is what limits e2e otherwise).  Best effort; returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        idx = local
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local < len(ids) and ids[local].isdigit():
                idx = int(ids[local])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def make_grid():
    from autonomouscar_b200 import worlds
    return worlds.block_grid(GRID, GRID, frac=0.01, block=16, seed=20181)


def params(mode):
    from autonomouscar_b200 import default_params
    return default_params(costmap_res=RES, costmap_width=GRID, costmap_height=GRID, collision_mode=mode)


# -----------------------------------------------------------------------------------------
# the CPU arm.  oracle/_ref = the reference's own local_planner.cpp compiled unmodified against
# ROS test doubles (oracle/Makefile); when it is present the reference's OWN collision check
# (checkCollision + calcCollisionMatrix, LP:443-496: axis-aligned cell box, the only footprint
# test the reference has -- it has no oriented one) is what gets timed, on the same poses and
# grid, one node instance per host thread.  The oracle's ORIENTED port is reported next to it.
# -----------------------------------------------------------------------------------------
def cpu_collide_rate(grid, mode, n_sample, threads, seed=1):
    from autonomouscar_b200 import worlds
    from oracle import binding as orc
    p = params(mode)
    box = worlds.pose_box(p)
    poses = orc.sample_poses(seed, 0, 0, n_sample, *box)
    _, secs = orc.collide_batch(p, grid, poses, n_threads=threads)
    return n_sample / secs, secs


def grid_as_msg(grid):
    """nav_msgs/OccupancyGrid.data whose LP:439 flip gives `grid` (planner order)."""
    import numpy as np
    return np.ascontiguousarray(grid.reshape(-1)[::-1])


class ReferenceCollider:
    """The reference's checkCollision on `threads` node instances (oracle/_ref, AABB variant)."""

    def __init__(self, grid, threads):
        from autonomouscar_b200 import PP_COLLISION_REF_AABB, worlds
        from oracle import ref_binding as ref
        self.p = params(PP_COLLISION_REF_AABB)
        self.box = worlds.pose_box(self.p)
        self.pool = ref.RefPool(self.p, grid_as_msg(grid), threads, ref.AABB)

    def rate(self, n_sample, seed=1):
        from oracle import binding as orc
        poses = orc.sample_poses(seed, 0, 0, n_sample, *self.box)
        flags, secs = self.pool.check_collision(poses)
        assert int((flags == 2).sum()) == 0      # the bench poses stay clear of the reference's UB
        return n_sample / secs, secs

    def close(self):
        self.pool.close()


def ref_available():
    try:
        from oracle import ref_binding as ref
        return ref.available()
    except Exception:
        return False


WORKLOAD = "BASELINE configs[2]: 2^24 oriented-footprint poses per GPU vs 4096^2 grid @0.1 m, 1% area in 16x16 blocks"


def headline_config(n_poses):
    """The static description of the headline workload -- identical in both arms (ours and --impl reference)."""
    return {"workload": WORKLOAD, "poses_per_gpu": int(n_poses), "collision_mode": "ORIENTED",
            "grid": "4096x4096 int8 @0.1 m", "footprint_m": [4.7, 1.586], "lattice": "48x17 samples",
            "l2_policy": "inputs larger than L2 (201 MB poses per step)"}


def run_reference(args):
    """The CPU arm on the SAME workload as ours (oriented footprint, same grid, same Philox poses): the reference
    has no oriented test (its checkCollision is an axis-aligned cell box, LP:443-496), so the headline value is the
    oracle's port of the ORIENTED test on all host threads (kind "port"); the reference's OWN collision code
    (oracle/_ref, REF_AABB semantics) is timed next to it on the same poses and reported inside cpu_baseline,
    where it pairs with our arm's roofline.ref_aabb figures."""
    from autonomouscar_b200 import PP_COLLISION_ORIENTED
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid = make_grid()
    cores = os.cpu_count() or 1
    # ~2^25 poses of CPU work for the whole run
    n_sample = max(1 << 15, min(1 << 21, (1 << 25) // max(1, args.steps)))
    one = lambda n: cpu_collide_rate(grid, PP_COLLISION_ORIENTED, n, cores)[0]      # noqa: E731
    for _ in range(args.warmup):
        one(1 << 16)
    rates = []
    t0 = time.time()
    for _ in range(args.steps):
        rates.append(one(n_sample))
    wall = time.time() - t0
    v = sum(rates) / len(rates)
    v1, _ = cpu_collide_rate(grid, PP_COLLISION_ORIENTED, 1 << 17, 1)
    cb = {"value": v, "unit": "poses/s", "cores": cores, "kind": "port",
          "sample": f"{n_sample} of the 2^24 poses per step, oracle port of the ORIENTED test, {cores} threads, g++ -O2",
          "one_thread_poses_per_s": v1}
    if ref_available():
        rc = ReferenceCollider(grid, cores)
        rc.rate(1 << 16)
        va, _ = rc.rate(1 << 21)
        rc.close()
        cb["ref_aabb"] = {"poses_per_s": va, "cores": cores, "kind": "reference",
                          "what": "the reference's own checkCollision (LP:443-496, :445 dropped) from oracle/_ref, REF_AABB semantics"}
    line = {
        "impl": "reference", "metric": "collision_checks_per_sec", "value": v, "unit": "poses/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 fixed-point Q7.24 + f64 set-up", "data": "synthetic",
        "config": headline_config(args.poses),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ncu --set full captures of this round's kernels at 2^22 poses (profiles/README.md): warp instructions and DRAM bytes per pose
NCU_ORIENTED = {"warp_inst_per_pose": 76065443 / (1 << 22), "dram_bytes_per_pose": (53.518592e6 + 157.952e3) / (1 << 22),
                "issue_active_pct_ncu": 63.5, "source": "profiles/r2_ncu_collide_oriented_v14.csv"}
NCU_DIRECT = {"warp_inst_per_pose": 1404516194 / (1 << 22), "issue_active_pct_ncu": 85.1,
              "source": "profiles/r2_ncu_collide_oriented_v12_direct.csv"}
NCU_AABB = {"warp_inst_per_pose": 31759359 / (1 << 24), "dram_bytes_per_launch_2p24": 208.604416e6 + 13.38752e6,
            "source": "profiles/r2_ncu_collide_aabb_stream_v3.csv"}


# -----------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from autonomouscar_b200 import (PP_COLLISION_ORIENTED, PP_COLLISION_REF_AABB, Planner, default_params,
                                    make_query, worlds)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    multi = world > 1
    numa = bind_to_gpu_cpus(local) if multi else None   # pinned host buffers of the e2e leg land next to the GPU
    if multi:
        args.skip_cpu = True                            # CPU-side figures are reported at N = 1 only
    if multi:
        # result gathers run on a high-priority stream so that their (few) CTAs are scheduled as soon as
        # collision-kernel CTAs retire instead of queueing behind the whole next launch
        opts = None
        if os.environ.get("PP_BENCH_NCCL_PRIORITY", "1") == "1":
            try:
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            except Exception:
                opts = None
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)

    grid = make_grid()
    p = params(PP_COLLISION_ORIENTED)
    pl = Planner(p, device=local)
    pl.set_grid(grid)
    box = worlds.pose_box(p)
    n = args.poses
    n_words = (n + 31) // 32
    # inputs resident in HBM before the timed region: the Philox sampler fills them (row S1), stream = rank
    poses = pl.sample_poses_dev(2018, rank, 0, n, *box)
    # two result buffers: the NCCL gather of step k overlaps the kernel of step k+1.  What is gathered is one BIT
    # per pose (pp_pack_flags_dev): 2 MiB per rank instead of 16 MiB
    flags2 = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    gather_done = [None, None]
    # How the bits travel (sharding.ResultGather).  "symm": every rank PULLS its peers' 2 MiB over NVLink with the
    # copy engines (peer memory mapped through torch's symmetric memory, two signal-pad barriers around the copies)
    # -- no SM is involved, so the gather runs beside the next step's persistent kernel, which leaves no SM free for
    # a collective's kernel.  "nccl": one ncclAllGather (its kernel ends up queued behind the next step's kernel:
    # measured 0.937 weak scaling at N = 8).  The symmetric-memory set-up falls back to NCCL when the platform
    # refuses it.
    gather_kind, rg = None, None
    bits_kernel = False
    if multi:
        from autonomouscar_b200.sharding import ResultGather
        rg = ResultGather(n_words, torch.int32, dev, transport=args.gather, slots=2)
        gather_kind = rg.transport
        if args.gather == "symm" and gather_kind != "symm":
            print(f"[bench] symmetric memory unavailable ({rg.why_not_symm}); gathering with NCCL", file=sys.stderr)
        bits_kernel = gather_kind == "symm" and not args.pack_pass
    bits2 = rg.send if multi else [None, None]
    gathered2 = [a.view(-1) for a in rg.all] if multi else [None, None]
    flags = flags2[0]
    step_no = [0]
    pl.sync()

    def step():
        i = step_no[0] & 1
        step_no[0] += 1
        if multi and gather_done[i] is not None:
            pl.stream.wait_event(gather_done[i])        # the buffer's previous gather has finished
        if multi and bits_kernel:
            pl.collide_bits_dev(poses, out=bits2[i])            # the batch kernel writes the bit per pose itself
        else:
            pl.collide_dev(poses, out=flags2[i])
        if multi:
            # result gather (north_star: NCCL only to gather results), ordered after the kernels.  (With NCCL the
            # kernel-written bits, pp_collide_batch_bits_dev, made the step 12 us longer at N = 2 and 27 us at N = 8:
            # the next step's persistent kernel then takes every SM before the all-gather kernel of this step gets
            # one.  The copy-engine gather needs no SM, and there the bits kernel saves the pack pass.)
            if not bits_kernel:
                pl.pack_flags_dev(flags2[i], out=bits2[i])
            ev = torch.cuda.Event()
            ev.record(pl.stream)
            gather_done[i] = rg.gather(i, after=ev)
        return i

    def barrier():
        pl.sync()
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
            torch.cuda.synchronize()

    kernel_ms = []
    clk = ClockSampler(local)
    for _ in range(max(3, args.warmup)):    # (re-)warm with the sampler child already streaming
        step()
    barrier()
    launches0 = pl.launch_count()
    clk.start()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(pl.stream)
    last = 0
    for _ in range(args.steps):
        last = step()
    if multi:
        pl.stream.wait_event(gather_done[last])
    e1.record(pl.stream)
    barrier()
    wall = time.perf_counter() - t0
    clk.stop()
    clk.close()
    dev_ms = e0.elapsed_time(e1)
    launches = pl.launch_count() - launches0
    # ---- N > 1: the gathered result equals what ONE GPU computes for every rank's shard (bit for bit)
    parity = None
    if multi:
        ok = True
        if rank == 0:
            got = gathered2[last].view(world, n_words)
            tmp_p = torch.empty_like(poses)
            tmp_f = torch.empty(n, dtype=torch.uint8, device=dev)
            for r in range(world):
                pl.sample_poses_dev(2018, r, 0, n, *box, out=tmp_p)     # rank r's pose stream, regenerated here
                pl.collide_dev(tmp_p, out=tmp_f)
                mine = pl.pack_flags_dev(tmp_f)
                pl.sync()
                ok = ok and bool(torch.equal(mine, got[r]))
            del tmp_p, tmp_f
        parity = {"gathered_flag_bits_equal_single_gpu_recompute": ok if rank == 0 else None, "ranks_checked": world}
    # per-launch duration of the dominant kernel (events recorded inside the library, on its stream)
    pl.sync()
    torch.cuda.synchronize()
    for _ in range(3):
        pl.collide_dev(poses, out=flags)
        kernel_ms.append(pl.last_phase_ms()[2])
    narrow_units = pl.last_narrow_count()
    hit_rate = float(flags.float().mean().item())
    t_ms = torch.tensor([dev_ms], device=dev)
    if multi:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / args.steps
    value = n * world / (ms_per_step * 1e-3)

    # ---- end to end through the host-pointer C-ABI call: pinned host poses in, host flags out
    h_poses = torch.empty((n, 3), dtype=torch.float32, pin_memory=True)
    h_poses.copy_(poses)
    h_flags = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    for _ in range(2):
        pl.collide(h_poses, out=h_flags)
    if multi:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pl.collide(h_poses, out=h_flags)
    e2e_s = (time.perf_counter() - t0) / args.steps
    t_e = torch.tensor([e2e_s], device=dev)
    if multi:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = n * world / float(t_e.item())
    assert torch.equal(h_flags, flags.cpu()), "host-pointer path and device path disagree"
    e2e_aabb = None
    # the same through the compact host format (pp_collide_batch_q16: int16 triples, 6 B per pose over PCIe)
    from autonomouscar_b200.capi import quantise_poses
    q_np, quant, back_np = quantise_poses(h_poses.numpy(), *box)
    h_q16 = torch.empty((n, 3), dtype=torch.int16, pin_memory=True)
    h_q16.copy_(torch.from_numpy(q_np))
    h_flags_q = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for _ in range(2):
        pl.collide_q16(h_q16, quant, out=h_flags_q)
    if multi:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pl.collide_q16(h_q16, quant, out=h_flags_q)
    t_q = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev)
    if multi:
        dist.all_reduce(t_q, op=dist.ReduceOp.MAX)
    e2e_q16 = n * world / float(t_q.item())
    back_dev = torch.from_numpy(back_np).to(dev)
    torch.cuda.synchronize()
    f_back = pl.collide_dev(back_dev)
    pl.sync()                                               # (the call is asynchronous on the handle's stream)
    q16_ok = bool(torch.equal(h_flags_q, f_back.cpu()))
    del back_dev, q_np, back_np

    # ---- secondary figures on rank 0 only: direct (no broad phase), REF_AABB, plans
    extra, roof_extra = {}, {}
    if rank == 0:
        def timed(fn, reps=5):
            fn(); pl.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(pl.stream)
            for _ in range(reps):
                fn()
            b.record(pl.stream)
            pl.sync()
            return a.elapsed_time(b) / reps
        pl.set_broad_phase(False)
        ms_direct = timed(lambda: pl.collide_dev(poses, out=flags))
        direct_units = pl.last_narrow_count()
        pl.set_broad_phase(True)
        pl.set_collision_mode(PP_COLLISION_REF_AABB)
        timed(lambda: pl.collide_dev(poses, out=flags), reps=3)
        ms_aabb_list = []
        for _ in range(10):                       # kernel-only time: events recorded by the library around the launch
            pl.collide_dev(poses, out=flags)
            ms_aabb_list.append(pl.last_phase_ms()[2])
        ms_aabb = sorted(ms_aabb_list)[len(ms_aabb_list) // 2]
        ms_aabb_loop = timed(lambda: pl.collide_dev(poses, out=flags), reps=20)   # back-to-back launches, gaps included
        aabb_hit = float(flags.float().mean().item())
        pl.collide(h_poses, out=h_flags)
        t0 = time.perf_counter()
        for _ in range(3):
            pl.collide(h_poses, out=h_flags)
        e2e_aabb = 3 * n / (time.perf_counter() - t0)
        pl.set_collision_mode(PP_COLLISION_ORIENTED)
        ms_sampler = timed(lambda: pl.sample_poses_dev(2018, rank, 0, n, *box, out=poses))
        ms_pack = timed(lambda: pl.pack_flags_dev(flags), reps=10)
        l2_gbps = pl.bench_l2_gather(16 << 20)      # random 32 B sectors of a 16 MiB (L2-resident) buffer
        # row S3 stand-alone: 2^22 random 0.45 m edges (one 150 ms step at 1.5 m/s), oriented footprint at every
        # search_res / 2 along each
        n_edges = 1 << 22
        e_xy = poses[:n_edges]
        edges = torch.empty((n_edges, 5), dtype=torch.float32, device=dev)
        edges[:, 0], edges[:, 1] = e_xy[:, 0], e_xy[:, 1]
        edges[:, 2] = e_xy[:, 0] + 0.45 * torch.cos(e_xy[:, 2])
        edges[:, 3] = e_xy[:, 1] + 0.45 * torch.sin(e_xy[:, 2])
        edges[:, 4] = e_xy[:, 2]
        e_flags = torch.empty(n_edges, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ms_edges = timed(lambda: pl.collide_edges_dev(edges, out=e_flags), reps=3)
        # row S2 stand-alone: brute-force nearest neighbour, 2^20 tree nodes (a config-4 sized tree) x 16384 queries
        nn_nodes, nn_q = 1 << 20, 1 << 14
        gen = torch.Generator(device=dev).manual_seed(7)
        node_xy = (torch.rand((nn_nodes, 2), generator=gen, device=dev) - 0.5) * 800.0
        query_xy = (torch.rand((nn_q, 2), generator=gen, device=dev) - 0.5) * 800.0
        torch.cuda.synchronize()
        ms_nn = timed(lambda: pl.nearest_dev(node_xy, query_xy), reps=3)
        peak_hbm = _peaks()[0]
        stream_bytes = 13 * n                      # what REF_AABB has to move: 12 B pose in + 1 B flag out
        roof_extra = {
            # the only variant that does SURVEY 8(d)'s per-pose work: every pose runs all 816 lattice samples
            "oriented_direct": {"poses_per_s": n / (ms_direct * 1e-3), "kernel_ms": ms_direct,
                                "algorithmic_GBps_829B_per_pose": ALG_BYTES_ORIENTED * n / (ms_direct * 1e-3) / 1e9,
                                "frac_of_hbm_peak": ALG_BYTES_ORIENTED * n / (ms_direct * 1e-3) / 1e9 / peak_hbm,
                                "frac_of_l2_gather_peak": ALG_BYTES_ORIENTED * n / (ms_direct * 1e-3) / 1e9 / l2_gbps,
                                "bound": "issue", "narrow_units": direct_units,
                                "warp_inst_per_pose": NCU_DIRECT["warp_inst_per_pose"], "inst_source": NCU_DIRECT["source"],
                                "issue_achieved_Gwarp_inst_per_s": NCU_DIRECT["warp_inst_per_pose"] * n / (ms_direct * 1e-3) / 1e9},
            "l2_gather_peak_GBps_measured": l2_gbps,
            # REF_AABB (rows A10/A11, the reference-faithful test): one map look-up per pose, HBM-stream bound
            "ref_aabb": {"kernel": "collide_aabb_stream_kernel", "poses_per_s": n / (ms_aabb * 1e-3), "kernel_ms": ms_aabb,
                         "bound": "hbm", "bytes_per_launch": stream_bytes,
                         "achieved_GBps": stream_bytes / (ms_aabb * 1e-3) / 1e9, "peak_GBps": peak_hbm,
                         "frac": stream_bytes / (ms_aabb * 1e-3) / 1e9 / peak_hbm, "hit_rate": aabb_hit,
                         "traffic": NCU_AABB["dram_bytes_per_launch_2p24"] * n / (1 << 24), "traffic_source": NCU_AABB["source"],
                         "ms_per_call_back_to_back": ms_aabb_loop,
                         "e2e_poses_per_s_host_buffers": e2e_aabb},
        }
        extra = {
            "philox_sampler_poses_per_s": n / (ms_sampler * 1e-3),
            "philox_sampler_write_GBps": 12 * n / (ms_sampler * 1e-3) / 1e9,
            "pack_flags_ms": ms_pack,
            "edge_collision": {"edges": n_edges, "steps_per_edge": 3, "ms": ms_edges, "edges_per_s": n_edges / (ms_edges * 1e-3),
                               "hit_rate": float(e_flags.float().mean().item())},
            "nearest_neighbour": {"nodes": nn_nodes, "queries": nn_q, "ms": ms_nn,
                                  "distance_evals_per_s": nn_nodes * nn_q / (ms_nn * 1e-3),
                                  "algorithmic_GBps_8B_per_node_per_query": 8.0 * nn_nodes * nn_q / (ms_nn * 1e-3) / 1e9},
        }
        del edges, e_flags, node_xy, query_xy
    plans = bench_plans(pl, rank, world, dev, multi, args) if not args.skip_plans else None
    costmap = bench_costmap(args) if (rank == 0 and not args.skip_plans) else None

    if rank == 0:
        peak, peak_kind = _peaks()
        clocks = clk.summary()
        # average launch duration of the dominant kernel: at N = 1 the timed region holds exactly one launch of it per
        # step and nothing else, so the region's own CUDA events (on the launch stream) give it; at N > 1 the region
        # also holds the gather, and the kernel is timed by the library's events around three launches after it
        k_iso = sum(kernel_ms) / len(kernel_ms)
        k_ms = dev_ms / args.steps if not multi else k_iso
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        issue_peak = sm_count * 4 * sm_mhz * 1e6 / 1e9          # G warp-instructions/s: 4 schedulers per SM, 1 per clock
        issue_achieved = NCU_ORIENTED["warp_inst_per_pose"] * n / (k_ms * 1e-3) / 1e9
        cpu = None
        if not args.skip_cpu and world == 1:   # the CPU baseline is reported at N = 1 only
            cores = os.cpu_count() or 1
            v_port, secs_port = cpu_collide_rate(grid, PP_COLLISION_ORIENTED, 1 << 22, cores)
            v1, _ = cpu_collide_rate(grid, PP_COLLISION_ORIENTED, 1 << 20, 1)
            cpu = {"value": v_port, "unit": "poses/s", "cores": cores, "kind": "port",
                   "sample": f"{1 << 22} of the 2^24 poses ({secs_port:.2f} s), oracle port of the ORIENTED test, g++ -O2",
                   "one_thread_poses_per_s": v1}
            if ref_available():
                n_sample = 1 << 22
                rc = ReferenceCollider(grid, cores)
                rc.rate(1 << 16)
                v, secs = rc.rate(n_sample)
                rc.close()
                rc1 = ReferenceCollider(grid, 1)
                v_1, _ = rc1.rate(1 << 17)
                rc1.close()
                cpu["ref_aabb"] = {"poses_per_s": v, "cores": cores, "one_thread_poses_per_s": v_1, "kind": "reference",
                                   "what": "the reference's own checkCollision (LP:443-496) from oracle/_ref; pairs with roofline.ref_aabb"}
        roofline = {"bound": "issue", "achieved": issue_achieved, "peak": issue_peak, "unit": "Gwarp-inst/s",
                    "frac": issue_achieved / issue_peak,
                    "how": "ncu warp instructions per pose x poses / CUDA-event kernel time vs SMs x 4 x SM clock",
                    "warp_inst_per_pose": NCU_ORIENTED["warp_inst_per_pose"], "inst_source": NCU_ORIENTED["source"],
                    "traffic": NCU_ORIENTED["dram_bytes_per_pose"] * n, "traffic_unit": "bytes/launch",
                    "kernel": "collide_oriented_kernel", "kernel_ms": k_ms, "kernel_ms_three_launches_after_the_region": k_iso,
                    "narrow_phase_poses": narrow_units,
                    # the byte view of the same launch: what it really moves (12 B pose + 1 B flag) against HBM
                    "hbm": {"bytes_per_launch": 13 * n, "achieved_GBps": 13 * n / (k_ms * 1e-3) / 1e9, "peak_GBps": peak,
                            "frac": 13 * n / (k_ms * 1e-3) / 1e9 / peak, "peak_kind": peak_kind},
                    "survey_829B_per_pose_GBps": ALG_BYTES_ORIENTED * n / (k_ms * 1e-3) / 1e9}
        roofline.update(roof_extra)
        if "oriented_direct" in roofline:
            roofline["oriented_direct"]["issue_frac"] = roofline["oriented_direct"]["issue_achieved_Gwarp_inst_per_s"] / issue_peak
        if plans:
            roofline["plans"] = {k: plans[k] for k in ("batch_queries", "plan_kernel_ms_rank0", "plans_per_s_kernel_rank0")
                                 if k in plans}
        e2e = {"value": e2e_value, "unit": "poses/s", "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": n,
               "h2d_GBps_per_rank": 12 * n / float(t_e.item()) / 1e9,
               "compact_int16_poses": {"value": e2e_q16, "h2d_bytes_per_step": 6 * n, "d2h_bytes_per_step": n,
                                       "flags_equal_float32_path_on_dequantised_poses": q16_ok,
                                       "what": "pp_collide_batch_q16: poses quantised to 6.3 mm / 9.6e-5 rad"}}
        if plans:
            e2e["plans"] = {k: plans[k] for k in ("batch_queries", "plans_per_s", "plans_per_s_with_host_copy_on_rank0",
                                                  "gathered_records_equal_single_gpu", "record_bytes", "single_query_ms",
                                                  "found") if k in plans}
        if parity:
            roofline["parity_at_n_gpus"] = parity
        cfg = headline_config(n)
        line = {
            "metric": "collision_checks_per_sec", "value": value, "unit": "poses/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 fixed-point Q7.24 + f64 set-up", "data": "synthetic",
            "config": cfg,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "wall_s_timed_region": wall,
            "run": {"hit_rate": hit_rate, "broad_phase": True, "rank_cpu_affinity": numa,
                    "gather": ({"symm": "1 bit per pose (" + ("pp_collide_batch_bits_dev" if bits_kernel else "pp_pack_flags_dev") + "), pulled from the peers over NVLink by the copy engines (symmetric memory)", "nccl": "1 bit per pose (pp_pack_flags_dev), ncclAllGather"}[gather_kind] if multi else None), "parity": parity},
            "extra": extra,
            "plans": plans,
            "costmap": costmap,
        }
        emit(line)
    pl.close()
    if multi:
        dist.destroy_process_group()


def bench_costmap(args):
    """Row N2: grids/s of the local costmap builder (pc_build, launch-file sizes: 300x300 @0.25 m grid,
    2 x 640-ray scans, 16 x 512 cloud, 12000^2 global road map resident in HBM) through the C ABI with
    host buffers, next to the reference's own local_costmap.cpp (oracle/_ref) on one host thread --
    the reference node is single-threaded and builds one grid per centre-laser scan (LC:366-370)."""
    import numpy as np

    from autonomouscar_b200 import LocalCostmap, worlds
    from autonomouscar_b200.costmap_capi import default_costmap_params
    G = 12000
    p = default_costmap_params(global_height_cells=float(G), global_width_cells=float(G))
    car = (30.0, -12.0, 0.7)
    tfs = worlds.costmap_tfs(*car)
    gg = worlds.road_like_global_grid(G, 17, zero_at=worlds.costmap_corner_global_cells(p, *car))
    scenes = [worlds.costmap_scene(s) for s in range(4)]
    have_ref = ref_available() and not args.skip_cpu
    if have_ref:
        from oracle import ref_binding as ref
        have_ref = ref.costmap_available()
    if have_ref:
        ref.costmap_set_tfs(tfs)                 # only the CPU leg (the reference node reads tf itself)
    for sc in scenes:                            # the GPU arm: planar tf numbers computed directly, no reference code
        worlds.fill_costmap_tf(sc, car)
    cm = LocalCostmap(p, device=0)
    cm.set_global_grid(gg)
    out = {"config": "300x300 @0.25 m, 2x640 rays, 16x512 cloud, 12000^2 global map (full_sim.launch sizes)"}
    grids = {}
    for want_host, key in ((True, "e2e_grids_per_s"), (False, "device_resident_grids_per_s")):
        for sc in scenes:
            grids[id(sc)] = cm.build(sc, want_host=want_host)
        reps, ms = 300, []
        t0 = time.perf_counter()
        for k in range(reps):
            cm.build(scenes[k % 4], want_host=want_host)
            ms.append(cm.last_build_ms())
        out[key] = reps / (time.perf_counter() - t0)
        out["device_ms_per_grid"] = float(np.median(ms))
    out["kernels_per_grid"] = 5
    # the query generator's region scan on the device grid (row N4, local_navigator.py:81-257)
    t0 = time.perf_counter()
    for k in range(300):
        mask = cm.front_regions()
    out["front_regions_call_us"] = (time.perf_counter() - t0) / 300 * 1e6
    out["front_regions_mask"] = int(mask)
    # scenario sweep: 256 scenes (one per simulated car) through one set of launches, masks back only
    batch = [scenes[k % 4] for k in range(256)]
    cm.build_batch(batch, want_host=False)
    t0 = time.perf_counter()
    for _ in range(5):
        cm.build_batch(batch, want_host=False)
    out["batch256_grids_per_s_masks_only"] = 5 * 256 / (time.perf_counter() - t0)
    out["batch256_device_ms"] = cm.last_build_ms()
    out["h2d_bytes_per_grid"] = int(4 * (len(scenes[0].right) + len(scenes[0].left) + scenes[0].cloud.size))
    out["d2h_bytes_per_grid"] = cm.n_cells
    if not args.skip_cpu:
        from oracle import binding as orc
        last = cm.build(scenes[0])
        det, _ = orc.costmap_build(p, scenes[0].inputs, gg, math_mode=orc.MATH_DET)
        out["bit_exact_vs_oracle"] = bool(np.array_equal(last, det))
        from oracle import navigator_ref
        cm.build(scenes[0], want_host=False)
        out["front_regions_equal_oracle"] = bool(cm.front_regions() == navigator_ref.region_mask(det, int(p.width_cells)))
        secs = [orc.costmap_build(p, sc.inputs, gg, math_mode=orc.MATH_LIBM)[1] for sc in scenes]
        out["cpu_oracle_port_grids_per_s_1thread"] = len(secs) / sum(secs)
        if have_ref:
            node = ref.RefCostmap(p)
            node.set_global(gg)
            node.build(scenes[0])
            rs = []
            for sc in scenes * 3:
                g_ref, sec = node.build(sc)
                rs.append(sec)
            out["cpu_reference_grids_per_s_1thread"] = len(rs) / sum(rs)
            out["cpu_reference_kind"] = "reference: local_costmap.cpp compiled from the reference sources (oracle/_ref), g++ -O2, 1 thread (the node is single-threaded)"
            # parity leg: the reference node reads its tf numbers through quaternion round trips; hand the CUDA builder
            # exactly those (a last-ulp difference in the map <- base_link transform moves road-overlay samples across
            # cell borders) and compare the two grids
            chk = worlds.costmap_scene(0)
            ref.costmap_derive_inputs(chk)
            out["cells_differing_from_reference"] = int((cm.build(chk) != node.build(chk)[0]).sum())
            node.close()
    cm.close()
    # row N3: the static global road map from the reference's own 303 road polylines (12000^2 cells)
    edges = os.path.join(ROOT, "tests", "golden", "road_edges.npz")
    if os.path.exists(edges):
        from autonomouscar_b200 import GlobalRoadMap
        z = np.load(edges)
        ms = []
        for _ in range(4):
            gm = GlobalRoadMap(z["xy"], z["offsets"], device=0, want_host=False)
            ms.append(gm.build_ms)
            gm.close()
        t0 = time.perf_counter()
        gm = GlobalRoadMap(z["xy"], z["offsets"], device=0, want_host=True)
        wall = time.perf_counter() - t0
        g = {"config": "12000x12000 @0.25 m from the 303 road polylines of edges.csv (10403 segments)",
             "device_ms": float(np.median(ms[1:])), "kernels": gm.launches, "e2e_ms_with_144MB_d2h": 1e3 * wall}
        gold = os.path.join(ROOT, "tests", "golden", "ref_global_costmap.npz")
        if os.path.exists(gold):
            import hashlib
            zz = np.load(gold)
            sha = np.frombuffer(hashlib.sha256(gm.host.tobytes()).digest(), dtype=np.uint8)
            g["identical_to_reference_published_grid"] = bool(np.array_equal(sha, zz["full_sha"]))
            g["cpu_reference_ms_recorded_in_container"] = 1e3 * float(zz["full_ref_seconds"][0])
        if not args.skip_cpu and have_ref and ref.global_costmap_available():
            import tempfile
            from autonomouscar_b200.costmap_capi import default_global_costmap_params
            with tempfile.TemporaryDirectory() as d:
                _, secs = ref.global_costmap_build(default_global_costmap_params(), z["xy"], z["offsets"], d)
            g["cpu_reference_ms_1thread"] = 1e3 * secs
            g["cpu_reference_kind"] = "reference: global_costmap.cpp compiled from the reference sources (oracle/_ref), g++ -O2, 1 thread"
        gm.close()
        out["global_road_map"] = g
        out["stack_rollout"] = bench_stack(args, z)
        out["stack_rollout_oriented_corridor"] = bench_stack(args, z, oriented=True)
        out["sweep_256_cars_own_costmaps"] = bench_sweep(args)
    out["route_rollouts"] = bench_rollouts(args)
    return out


def bench_rollouts(args, cars=1024, ticks=20):
    """Row N4: `cars` independent closed-loop drives along the reference route, one pp_plan_batch call per
    20 Hz tick (way-point rule of local_navigator.py:586-609, profile follower of controller.cpp:144-158,
    ideal tracker).  The host loop around the call is plain Python; the planning share is timed separately."""
    from autonomouscar_b200 import PP_COLLISION_AS_SHIPPED, Planner, default_params, rollout
    pl = Planner(default_params(collision_mode=PP_COLLISION_AS_SHIPPED, max_nodes=20000), device=0)
    starts = [(0.12 * k, 15.0 + 0.05 * (k % 5), 0.02 * ((k % 7) - 3), 0.1 * (k % 12)) for k in range(cars)]
    ro = rollout.RouteRollout(starts)
    inner = rollout.plan_batch_cuda(pl)
    t_plan = [0.0]

    def timed(qs):
        t0 = time.perf_counter()
        r = inner(qs)
        t_plan[0] += time.perf_counter() - t0
        return r
    ro.tick(timed)                      # warm-up tick
    t_plan[0] = 0.0
    t0 = time.perf_counter()
    for _ in range(ticks):
        ro.tick(timed)
    wall = time.perf_counter() - t0
    pl.close()
    sim_s = ticks / 20.0
    native = None
    try:    # the same loop as native host code (autonomouscar_b200/host/route_rollout_b200.hpp), no interpreter in between
        import subprocess
        import tempfile
        pkg = os.path.join(ROOT, "autonomouscar_b200")
        with tempfile.TemporaryDirectory() as d:
            exe = os.path.join(d, "route_rollout")
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host"),
                                   os.path.join(ROOT, "tests", "cpp", "test_route_rollout.cpp"), "-o", exe, "-L", pkg,
                                   "-lprius_planner_b200", f"-Wl,-rpath,{pkg}"], stderr=subprocess.DEVNULL)
            first = subprocess.check_output([exe, str(cars), str(ticks), "0.12", "1"]).decode().splitlines()[0].split()
        native = {"wall_s": float(first[0]), "plans": int(first[1]) - cars, "real_time_factor_whole_loop": sim_s / float(first[0]),
                  "car_ticks_per_s": cars * ticks / float(first[0])}
    except Exception as e:      # no compiler on the box: the Python figures stand alone
        native = {"unavailable": str(e)[:100]}
    return {"cars": cars, "ticks": ticks, "plans": cars * ticks, "found": int(ro.found - cars), "native_host_loop": native,
            "plans_per_s_in_plan_calls": cars * ticks / t_plan[0], "car_ticks_per_s_whole_loop": cars * ticks / wall,
            "real_time_factor_planning_only": sim_s / t_plan[0], "real_time_factor_whole_loop": sim_s / wall,
            "mean_metres_driven": float((ro.x - [s_[0] for s_ in starts]).mean())}


def bench_stack(args, edges, cycles=150, oriented=False):
    """Rows N1-N3 chained with the hot path, everything resident on the GPU: the global road map is built
    once (pg_build) and adopted by the costmap builder; then, per 20 Hz cycle along the reference's own route
    (global_path_planner.py:101-102, densified and looked ahead as local_navigator.py:27-46,586-609 do):
    pc_build (synthetic laser messages) -> pc_grid_dev -> pp_set_grid_dev -> pp_plan (as shipped) -> the car
    advances along the planned path for one 50 ms period (kinematic stand-in for Gazebo)."""
    import numpy as np

    from autonomouscar_b200 import (PP_COLLISION_AS_SHIPPED, PP_COLLISION_ORIENTED, PP_PLAN_FOUND, GlobalRoadMap,
                                    LocalCostmap, Planner, default_params, make_query, worlds)
    from autonomouscar_b200.costmap_capi import default_costmap_params
    gm = GlobalRoadMap(edges["xy"], edges["offsets"], device=0, want_host=False)
    cm = LocalCostmap(default_costmap_params(), device=0)
    cm.adopt_global_grid(gm)
    mode = PP_COLLISION_ORIENTED if oriented else PP_COLLISION_AS_SHIPPED
    pl = Planner(default_params(collision_mode=mode, max_nodes=20000), device=0)
    route = worlds.densified_route()
    # as shipped: random laser returns (the grid does not influence the search, LP:445); oriented: two walls
    # 2.5 m either side of the car, which the footprint test has to respect
    scenes = [worlds.corridor_scene(50 + k, half_width=2.5) if oriented else worlds.costmap_scene(50 + k) for k in range(8)]
    x, y, yaw, v, wp = 0.0, 12.6 if oriented else 15.0, 0.0, 0.0, 0
    rejected = 0
    t_cost, t_ingest, t_plan, found, dist = [], [], [], 0, 0.0
    for k in range(cycles + 5):
        sc = scenes[k % 8]
        worlds.fill_costmap_tf(sc, (x, y, yaw))
        t0 = time.perf_counter()
        cm.build(sc, want_host=False)
        t1 = time.perf_counter()
        cm.feed_planner(pl)
        pl.sync()
        t2 = time.perf_counter()
        wp, (wx, wy) = worlds.next_waypoint(route, wp, x, y)
        c, s_ = np.cos(yaw), np.sin(yaw)
        gx, gy = c * (wx - x) + s_ * (wy - y), -s_ * (wx - x) + c * (wy - y)     # convertGoalToLocalFrame
        r = pl.plan(make_query(gx, gy, start_speed=v), cap_nodes=8, want_nodes=False)
        t3 = time.perf_counter()
        rejected += int(r.n_collisions)
        if k >= 5:
            t_cost.append(t1 - t0); t_ingest.append(t2 - t1); t_plan.append(t3 - t2)
        if r.status == PP_PLAN_FOUND and r.n_profile > 0:
            found += k >= 5
            v = float(r.speeds[0])
            yaw += float(r.yaw_rates[0]) * 0.05
            x += v * 0.05 * np.cos(yaw)
            y += v * 0.05 * np.sin(yaw)
            dist += v * 0.05
    cm.close()
    pl.close()
    med = lambda a: 1e3 * float(np.median(a))                                       # noqa: E731
    tot = np.array(t_cost) + np.array(t_ingest) + np.array(t_plan)
    return {"cycles": cycles, "collision_mode": "ORIENTED, corridor scenes" if oriented else "AS_SHIPPED, random scenes",
            "plans_found": int(found), "candidate_poses_rejected": rejected, "metres_driven": dist,
            "ms_costmap": med(t_cost), "ms_grid_ingest": med(t_ingest), "ms_plan": med(t_plan),
            "ms_cycle": med(tot), "cycles_per_s": float(len(tot) / tot.sum()),
            "note": "20 Hz is the reference's plan rate (local_navigator.py:529); its own nodes need ~20 ms for the "
                    "costmap and 3-5 ms for a free-space plan per cycle on one core each"}


def bench_sweep(args, cars=256, ticks=8):
    """Scenario sweep with sensors in the loop, everything batched: per 20 Hz tick ONE pc_build_batch (all cars'
    costmaps + region masks), ONE pp_set_grid_bank_dev (all grids' derived maps, fed on the device) and ONE
    pp_plan_batch_banked (car k plans on grid k with the ORIENTED footprint test)."""
    import numpy as np

    from autonomouscar_b200 import PP_COLLISION_ORIENTED, PP_PLAN_FOUND, LocalCostmap, Planner, default_params, make_query, worlds
    from autonomouscar_b200.capi import pp_query
    from autonomouscar_b200.costmap_capi import default_costmap_params
    from autonomouscar_b200.results import PlanBatch
    cm = LocalCostmap(default_costmap_params(), device=0)
    pl = Planner(default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=20000), device=0)
    base = []
    for k in range(16):
        sc = worlds.corridor_scene(300 + k, half_width=2.5 + 0.1 * (k % 4))       # walls 2.5 .. 2.8 m either side
        worlds.fill_costmap_tf(sc, (2.0 * k, 0.0, 0.0))
        base.append(sc)
    scenes = [base[k % 16] for k in range(cars)]
    qs = (pp_query * cars)(*[make_query(18.0, 1.4 - 0.7 * (k % 5), start_speed=0.5, grid_index=k, query_id=k) for k in range(cars)])
    buf = PlanBatch(cars, cap_path=256)
    t_cost, t_bank, t_plan = [], [], []
    for k in range(ticks + 2):
        t0 = time.perf_counter()
        cm.build_batch(scenes, want_host=False, want_masks=True)
        t1 = time.perf_counter()
        cm.feed_planner_bank(pl)
        t2 = time.perf_counter()
        pl.plan_batch_into(qs, buf, banked=True)
        t3 = time.perf_counter()
        if k >= 2:
            t_cost.append(t1 - t0); t_bank.append(t2 - t1); t_plan.append(t3 - t2)
    found = int((buf.field("status") == PP_PLAN_FOUND).sum())
    rejected = int(buf.field("n_collisions").sum())
    cm.close()
    pl.close()
    med = lambda a: 1e3 * float(np.median(a))                                       # noqa: E731
    tick = med(np.array(t_cost) + np.array(t_bank) + np.array(t_plan))
    return {"cars": cars, "collision_mode": "ORIENTED, corridor scenes, one costmap per car", "plans_found": found,
            "candidate_poses_rejected": rejected, "ms_costmaps": med(t_cost), "ms_grid_bank": med(t_bank),
            "ms_plans": med(t_plan), "ms_tick": tick, "car_ticks_per_s": cars / (tick / 1e3),
            "x_real_time_at_20hz": cars / (tick / 1e3) / 20.0 / cars}


def bench_plans(pl_unused, rank, world, dev, multi, args):
    """configs[1] (single query, 10k-node cap) and configs[4] (independent queries, sharded
    contiguously over the ranks, fixed-stride result records gathered with NCCL).
    Launch-file grid (300x300 @0.25 m)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from autonomouscar_b200 import (PP_COLLISION_ORIENTED, PP_PLAN_FOUND, PlanBatch, PlanRecords, Planner, default_params,
                                    sharding, worlds)
    from autonomouscar_b200.capi import pp_query
    from autonomouscar_b200.results import record_stride
    p = default_params(collision_mode=PP_COLLISION_ORIENTED, max_nodes=10000)
    grid = worlds.block_grid(300, 300, frac=0.01, block=4, seed=9)
    worlds.clear_disc(grid, p, 0.0, 0.0, 4.0)
    pl = Planner(p, device=dev.index)
    pl.set_grid(grid)
    n_total = args.queries * world
    qs_all = worlds.forward_queries(n_total, seed=11)
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    qs = qs_all[lo:hi]
    q_arr = (pp_query * len(qs))(*qs)
    CAP = 96                                                    # path slots per record (longest path here: 73 poses)
    stride = record_stride(CAP)
    warm = PlanBatch(64, cap_path=256)
    pl.plan_batch_into((pp_query * 64)(*qs[:64]), warm)                       # warm-up
    buffers = {}
    sharding.plan_batch_gather(pl, q_arr, n_total, cap_path=CAP, buffers=buffers)
    h_rec = torch.empty((sharding.padded_block(n_total, world) * world, stride), dtype=torch.uint8, pin_memory=True) if rank == 0 else None
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    # BASELINE configs[4]: every rank plans its block, the fixed-stride records (summary + path + profile) stay on
    # the device, ONE all-gather over NVLink; the rank that wants the results (rank 0) copies them to the host once
    reps, dts, dts_host, kms = 5, [], [], []
    for with_host in (False, True):
        for _ in range(reps):
            if multi:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            allrec = sharding.plan_batch_gather(pl, q_arr, n_total, cap_path=CAP, buffers=buffers)
            if with_host and rank == 0:
                h_rec.copy_(allrec, non_blocking=True)        # the rank that wants the results on the host copies them, once
            pl.sync()
            torch.cuda.synchronize()
            (dts_host if with_host else dts).append(sharding.max_over_ranks(time.perf_counter() - t0, dev))
            kms.append(pl.last_phase_ms()[4])
    dt = sorted(dts)[len(dts) // 2]
    dt_h = sorted(dts_host)[len(dts_host) // 2]
    rows = sharding.gathered_rows(n_total, world)
    rec_view = PlanRecords(h_rec.numpy()[rows], CAP) if rank == 0 else None
    # the same queries through the host-pointer call (pp_plan_batch: results scattered into caller arrays)
    buf = PlanBatch(len(qs), cap_path=256)
    pl.plan_batch_into(q_arr, buf)
    t0 = time.perf_counter()
    pl.plan_batch_into(q_arr, buf)
    dt_host = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    records_equal = None
    if rank == 0 and multi:
        # N > 1 parity: ONE GPU plans all n_total queries; the gathered records must equal that run byte for byte
        one = pl.plan_batch_dev((pp_query * n_total)(*qs_all), n_total, cap_path=CAP)
        pl.sync()
        records_equal = bool(np.array_equal(one.cpu().numpy(), h_rec.numpy()[rows]))
        del one
    # single-query latency (config 2)
    lat, lat_k = [], []
    one = PlanBatch(1, cap_path=256)
    q_one = (pp_query * 1)(qs[0])
    for k in range(7):
        t1 = time.perf_counter()
        pl.plan_batch_into(q_one, one)          # the C-ABI call with host buffers, nothing else
        lat.append(time.perf_counter() - t1)
        lat_k.append(pl.last_phase_ms()[4])
    r1 = pl.plan(qs[0], cap_nodes=16384)
    # configs[0] / [1]: the reference's first plan request on the car_demo road world @0.1 m
    cfg1 = None
    fixture = os.path.join(ROOT, "tests", "golden", "road_world_4096.npz")
    if rank == 0 and os.path.exists(fixture):
        from autonomouscar_b200 import PP_COLLISION_REF_AABB, make_query as mq
        rgrid, rres, spawn = worlds.road_world(fixture)
        gx, gy = worlds.route_queries(spawn)
        p1 = default_params(costmap_res=rres, costmap_width=4096, costmap_height=4096,
                            collision_mode=PP_COLLISION_REF_AABB, max_nodes=10000)
        pl1 = Planner(p1, device=dev.index)
        pl1.set_grid(rgrid)
        q1 = mq(gx, gy, start_speed=0.0)
        pl1.plan(q1)
        l1 = []
        for _ in range(5):
            t1 = time.perf_counter()
            g1 = pl1.plan(q1, cap_nodes=16384)
            l1.append(time.perf_counter() - t1)
        cfg1 = {"world": "car_demo road polylines, off-road = occupied, 4096x4096 @0.1 m", "mode": "REF_AABB",
                "goal_vehicle_frame": [gx, gy], "gpu_ms": 1e3 * sorted(l1)[2], "nodes": int(g1.n_nodes),
                "path_poses": int(g1.n_path), "status": int(g1.status)}
        if not args.skip_cpu:
            from oracle import binding as orc
            t1 = time.perf_counter()
            c1 = orc.plan(p1, rgrid, q1, math_mode=orc.MATH_LIBM)
            cfg1["cpu_oracle_ms_1thread"] = 1e3 * (time.perf_counter() - t1)
            if ref_available():
                # BASELINE configs[0]: the reference planner itself, headless (oracle/_ref = its own local_planner.cpp
                # with the collision code of LP:446-496 enabled), one thread like the node
                from oracle import ref_binding as ref
                pool1 = ref.RefPool(p1, grid_as_msg(rgrid), 1, ref.AABB)
                pool1.plan_batch([q1])
                r_secs, r_st, r_nn = pool1.plan_batch([q1])
                pool1.close()
                cfg1["cpu_reference_ms_1thread"] = 1e3 * r_secs
                cfg1["cpu_reference_kind"] = "reference: local_planner.cpp from /root/reference (oracle/_ref), g++ -O2"
                cfg1["reference_same_outcome_and_tree_size"] = bool((int(r_st[0]) != 0) == (int(g1.status) != 0) and int(r_nn[0]) == int(g1.n_nodes))
            d1 = orc.plan(p1, rgrid, q1, math_mode=orc.MATH_DET)
            cfg1["bit_exact_vs_oracle"] = bool(d1.summary() == g1.summary() and np.array_equal(d1.path_xy, g1.path_xy)
                                               and np.array_equal(d1.node_parent, g1.node_parent))
        pl1.close()
    # sampling planner (extension rows S1-S3): batch of independent queries on a 1024^2 grid
    rp = default_params(costmap_res=0.1, costmap_width=1024, costmap_height=1024, collision_mode=PP_COLLISION_ORIENTED,
                        time_step_ms=1000.0, rrt_samples_per_round=64, rrt_max_samples=10000, max_nodes=10000,
                        rrt_goal_tolerance=1.5)
    rg = worlds.block_grid(1024, 1024, frac=0.01, block=16, seed=2, noise=False)
    worlds.clear_disc(rg, rp, -30.0, -30.0, 5.0)
    rng = np.random.default_rng(5)
    from autonomouscar_b200 import make_query
    n_rrt = 256
    rqs = []
    for k in range(n_rrt):
        gx, gy = rng.uniform(-20, 40, 2)
        worlds.clear_disc(rg, rp, gx, gy, 4.0)
        rqs.append(make_query(gx, gy, start_x=-30.0, start_y=-30.0, start_heading=0.7, start_speed=1.0, seed=7,
                              query_id=rank * n_rrt + k))
    rpl = Planner(rp, device=dev.index)
    rpl.set_grid(rg)
    rq_arr = (pp_query * n_rrt)(*rqs)
    rbuf = PlanBatch(n_rrt, cap_path=2048)
    rpl.plan_batch_into(rq_arr, rbuf, rrt=True)
    t0 = time.perf_counter()
    rpl.plan_batch_into(rq_arr, rbuf, rrt=True)
    rrt_dt = sharding.max_over_ranks(time.perf_counter() - t0, dev)
    rrt_samples = int(rbuf.field("n_candidates").sum())
    rrt_found = int((rbuf.field("status") == PP_PLAN_FOUND).sum())
    # BASELINE configs[3]: 8192^2 grid @0.1 m, 1 % of the area in 16x16-cell blocks, one 1M-sample query corner to corner
    cfg4 = None
    if rank == 0:
        N4 = 8192
        g4 = worlds.block_grid(N4, N4, frac=0.01, block=16, seed=4, noise=False)
        half = N4 * 0.1 / 2
        sx, sy, gx4, gy4 = -half + 10.0, -half + 10.0, half - 10.0, half - 10.0
        cfg4 = {"grid": "8192x8192 @0.1 m, 1% area in 16x16-cell blocks, start / goal at opposite corners inset 10 m"}
        for K4 in (256, 1024):
            p4 = default_params(costmap_res=0.1, costmap_width=N4, costmap_height=N4, collision_mode=PP_COLLISION_ORIENTED,
                                time_step_ms=1000.0, rrt_samples_per_round=K4, rrt_max_samples=1 << 20,
                                max_nodes=(1 << 20) + 1, rrt_goal_tolerance=2.0, rrt_goal_bias=0.05)
            if K4 == 256:
                worlds.clear_disc(g4, p4, sx, sy, 6.0)
                worlds.clear_disc(g4, p4, gx4, gy4, 6.0)
            pl4 = Planner(p4, device=dev.index)
            pl4.set_grid(g4)
            q4 = make_query(gx4, gy4, start_x=sx, start_y=sy, start_heading=0.785, start_speed=1.0, max_speed=1.5,
                            seed=2018, query_id=0)
            pl4.rrt_plan(q4, cap_nodes=8, cap_path=65536, want_nodes=False)
            t0 = time.perf_counter()
            r4 = pl4.rrt_plan(q4, cap_nodes=8, cap_path=65536, want_nodes=False)
            w4 = time.perf_counter() - t0
            cfg4[f"K{K4}"] = {"status": int(r4.status), "samples": int(r4.n_candidates), "nodes": int(r4.n_nodes),
                              "rounds": int(r4.n_expansions), "path_nodes": int(r4.n_path), "wall_s": w4,
                              "samples_per_s": r4.n_candidates / w4}
            if K4 == 256:
                # the whole 2^20-sample budget (goal box switched off so the run cannot end early): ~980 k nodes
                p4f = default_params(costmap_res=0.1, costmap_width=N4, costmap_height=N4, collision_mode=PP_COLLISION_ORIENTED,
                                     time_step_ms=1000.0, rrt_samples_per_round=K4, rrt_max_samples=1 << 20,
                                     max_nodes=(1 << 20) + 1, rrt_goal_tolerance=0.0, rrt_goal_bias=0.05)
                pl4f = Planner(p4f, device=dev.index)
                pl4f.set_grid(g4)
                pl4f.rrt_plan(q4, cap_nodes=8, cap_path=64, want_nodes=False)
                t0 = time.perf_counter()
                rf = pl4f.rrt_plan(q4, cap_nodes=8, cap_path=64, want_nodes=False)
                wf = time.perf_counter() - t0
                cfg4["K256_full_budget"] = {"samples": int(rf.n_candidates), "nodes": int(rf.n_nodes), "rounds": int(rf.n_expansions),
                                            "wall_s": wf, "kernel_ms": pl4f.last_phase_ms()[4], "samples_per_s": rf.n_candidates / wf,
                                            "us_per_round": 1e3 * pl4f.last_phase_ms()[4] / max(1, rf.n_expansions),
                                            "tree_digest": int(rf.tree_digest)}
                pl4f.close()
            cfg4[f"K{K4}"]["kernel_ms"] = pl4.last_phase_ms()[4]
            cfg4[f"K{K4}"]["us_per_round"] = 1e3 * pl4.last_phase_ms()[4] / max(1, r4.n_expansions)
            if K4 == 256 and not args.skip_cpu:
                # parity + CPU rate on the SAME run, whole tree: the oracle answers its nearest-neighbour queries through an
                # exact bucket grid (oracle/sampling_ref.h), which lets it replay the 265 k-node tree in seconds
                from oracle import binding as orc
                t0 = time.perf_counter()
                c4r = orc.rrt_plan(p4, g4, q4, cap_nodes=8, cap_path=65536, want_nodes=False)
                c4 = time.perf_counter() - t0
                cfg4["whole_run_vs_oracle"] = {"cpu_oracle_samples_per_s_1thread": c4r.n_candidates / c4,
                                               "bit_exact_vs_oracle": bool(r4.summary() == c4r.summary()
                                                                           and np.array_equal(r4.path_xy, c4r.path_xy))}
            pl4.close()
        del g4
    out = None
    if rank == 0:
        found = int((rec_view.field("status") == PP_PLAN_FOUND).sum())
        k_ms = sorted(kms)[len(kms) // 2]
        out = {"batch_queries": n_total, "plans_per_s": n_total / dt, "batch_wall_s": dt,
               "what": "host queries in, pp_plan_batch_dev per rank, ONE all-gather of fixed-stride records: all results on every GPU",
               "plans_per_s_with_host_copy_on_rank0": n_total / dt_h, "host_copy_bytes": int(h_rec.numel()),
               "record_bytes": stride, "gathered_records_equal_single_gpu": records_equal,
               "plans_per_s_host_pointer_call": n_total / dt_host,
               "plan_kernel_ms_rank0": k_ms, "plans_per_s_kernel_rank0": len(qs) / (k_ms * 1e-3), "found": found,
               "mean_nodes": float(rec_view.field("n_nodes").mean()),
               "longest_path": int(rec_view.field("n_path").max()),
               "single_query_ms": 1e3 * sorted(lat)[len(lat) // 2], "single_query_kernel_ms": sorted(lat_k)[len(lat_k) // 2],
               "single_query_nodes": int(r1.n_nodes),
               "grid": "300x300 @0.25 m (full_sim.launch), ORIENTED, node cap 10000",
               "config1_road_world": cfg1, "config4_dense_clutter_1M_samples": cfg4,
               "rrt": {"queries_per_gpu": n_rrt, "plans_per_s": n_rrt * world / rrt_dt, "found_rank0": rrt_found,
                       "samples_per_s_rank0": rrt_samples / rrt_dt, "grid": "1024x1024 @0.1 m, 10k-sample cap, K=64"}}
        if not args.skip_cpu:
            from oracle import binding as orc
            cores = os.cpu_count() or 1
            nq = min(len(qs_all), 256)
            secs, st, nn, dg = orc.plan_batch_timed(p, grid, qs_all[:nq], math_mode=orc.MATH_LIBM, n_threads=cores)
            out["cpu_plans_per_s"] = nq / secs
            out["cpu_cores"] = cores
            secs1, *_ = orc.plan_batch_timed(p, grid, qs_all[:32], math_mode=orc.MATH_LIBM, n_threads=1)
            out["cpu_plans_per_s_1thread"] = 32 / secs1
            # parity of the sample against the deterministic oracle
            secs2, st2, nn2, dg2 = orc.plan_batch_timed(p, grid, qs_all[:nq], math_mode=orc.MATH_DET, n_threads=cores)
            out["parity_sample_ok"] = bool((rec_view.field("status")[:nq] == st2).all() and (rec_view.field("n_nodes")[:nq] == nn2).all()
                                           and (rec_view.field("tree_digest")[:nq] == dg2.astype(np.uint64)).all())
            if ref_available():
                # the reference node itself (oracle/_ref), REF_AABB = its own collision code enabled:
                # same queries through the CUDA path in that mode and through the reference binary
                from autonomouscar_b200 import PP_COLLISION_REF_AABB
                from oracle import ref_binding as ref
                p_a = default_params(collision_mode=PP_COLLISION_REF_AABB, max_nodes=10000)
                pl_a = Planner(p_a, device=dev.index)
                pl_a.set_grid(grid)
                qa = qs_all[:nq]
                qa_arr = (pp_query * nq)(*qa)
                buf_a = PlanBatch(nq, cap_path=256)
                pl_a.plan_batch_into(qa_arr, buf_a)
                t1 = time.perf_counter()
                pl_a.plan_batch_into(qa_arr, buf_a)
                gpu_dt = time.perf_counter() - t1
                pool = ref.RefPool(p_a, grid_as_msg(grid), cores, ref.AABB)
                r_secs, r_st, r_nn = pool.plan_batch(qa)
                pool.close()
                pool1 = ref.RefPool(p_a, grid_as_msg(grid), 1, ref.AABB)
                r1_secs, _, _ = pool1.plan_batch(qa[:32])
                pool1.close()
                g_rec = buf_a.records()
                out["reference_node"] = {
                    "what": "Prius::LocalPlanner compiled from the reference sources (oracle/_ref), collision code "
                            "of LP:446-496 enabled, one node instance per thread; same queries on the GPU in REF_AABB mode",
                    "queries": nq, "cpu_plans_per_s": nq / r_secs, "cpu_cores": cores,
                    "cpu_plans_per_s_1thread": 32 / r1_secs, "gpu_plans_per_s_same_queries": nq / gpu_dt,
                    "same_outcome_and_tree_size_as_reference": bool(((g_rec[:, 0] != 0) == (r_st != 0)).all()
                                                                    and (g_rec[:, 2] == r_nn).all())}
                pl_a.close()
    pl.close()
    rpl.close()
    return out


_JSON_FD = None


def emit(line):
    """The ONE stdout line of the contract.  Everything else any library prints to fd 1 during the run
    (NCCL's version banner, for one) has been diverted to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--poses", type=int, default=N_POSES)
    ap.add_argument("--queries", type=int, default=1024, help="plan queries per GPU")
    ap.add_argument("--gather", default="symm", choices=["symm", "nccl"], help="N > 1: transport of the result bits")
    ap.add_argument("--pack-pass", action="store_true", help="N > 1 with --gather symm: byte flags + pp_pack_flags_dev instead of the bits kernel")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-plans", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
