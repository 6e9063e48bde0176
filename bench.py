#!/usr/bin/env python
"""Benchmark of the Nearl featurisation hot path on B200 (contract: see DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A step = featurise ONE synthetic trajectory of BASELINE.json configs[1]: 20 000 atoms x 100
frames, 32^3 grid at 0.5 A, window 10, focus ':ARG', four dynamic features (existence,
distinct_count(atomic_number), mean_distance(mass), density_flow(mass); mean over the window)
= 40 (feature, window, focus) tasks consuming 100 frames.  metric = featurised frames / s.

  value : nrl_featurize_batch with the trajectory already resident in HBM and the results left in
          HBM; the K independent steps are issued round-robin over --lanes lanes of the library
          (stream + workspace each), forked off and joined back into torch's stream between the
          two CUDA events that time them (max over ranks).
  e2e   : Featurizer.run() on host trajectories in pinned memory -- host lowering, H2D of the
          frames, kernels, D2H of the 40 grids -- nothing cached between steps.
Eight different trajectories per rank are cycled (8 x 24 MB of coordinates > L2).

--impl reference runs the reference's own CUDA kernels (compiled unmodified from
/root/reference/src into oracle/_ref by oracle/build_ref.sh) behind a numpy restatement of the
reference's per-task host pipeline (oracle/host_ref.py), i.e. "Nearl as shipped" on the same
GPU; when that module is unavailable it falls back to the CPU oracle port on the host cores.
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

METRIC = "featurised frames/sec"
UNIT = "frames/s"
N_ROTATE = 8
WORKLOADS = {
  # the headline: BASELINE.json configs[1]
  "dynamic": ("configs[1]: dynamic marching-observer features (existence, distinct_count, mean_distance) + density_flow, "
              "32^3 @0.5A, time_window=10, cutoff 5.0, sigma 1.5, focus ':ARG', synthetic 20k-atom x 100-frame trajectory"),
  # extra, for the record (profiles/): BASELINE.json configs[0] and configs[2]
  "readme": ("configs[0]: README quick start, Mass, 32^3 @0.5A, time_window=10, sigma 1.5, cutoff 5.0, focus ':ARG', "
             "synthetic 5k-atom x 100-frame trajectory"),
  "pdbbind": ("configs[2]: PDBBind-style static featurisation, 8 atom-property channels at 24^3 over 20 A, cutoff 2.5, sigma 1.5, "
              "focus ':LIG', synthetic 3k-atom complexes (one frame each; a step = one complex)"),
  "misato": ("configs[3]: MISATO-style batch, 3 static + 2 marching-observer + 2 density-flow channels at 32^3 over 20 A, window 10, "
             "cutoff 2.5 static / 5.0 dynamic, sigma 1.5, focus ':ARG', synthetic 6k-atom x 100-frame trajectories"),
  "misato_production": ("the feature set of the reference's MISATO production script (scripts/featurize_misato.py:107-217, labels left "
                        "out): 10 static atom-type channels on protein / ligand selections, 2 distinct_count observers and 4 density "
                        "flows (uniformed weights, mean / standard deviation), 32^3 over 20 A, window 10, cutoff 2.5, sigma 1.5, focus on "
                        "the ligand, synthetic 6k-atom x 100-frame trajectories"),
  "stress": ("configs[4]: 128^3 grid at 0.5 A, synthetic 200k-atom x 200-frame trajectory, time_window=100, cutoff 5.0, observers existence / "
             "distinct_count / mean_distance + 2 density flows, mean and standard deviation (reference's 9999-atom cap kept)"),
}
WORKLOAD = WORKLOADS["dynamic"]


def make_workload(name, n_traj, seed_offset=0, pin=True):
  from nearl_b200 import workloads
  if name == "dynamic":
    return workloads.dynamic_observers(n_traj=n_traj, seed_offset=seed_offset, pin=pin)
  if name == "readme":
    return workloads.readme_mass(n_traj=n_traj, seed_offset=seed_offset, pin=pin)
  if name == "pdbbind":
    return workloads.pdbbind_static(n_complex=n_traj * 8, seed_offset=seed_offset, pin=pin)
  if name == "misato":
    return workloads.misato_mixed(n_traj=n_traj, seed_offset=seed_offset, pin=pin)
  if name == "misato_production":
    return workloads.misato_production(n_traj=n_traj, seed_offset=seed_offset, pin=pin)
  if name == "stress":
    return workloads.stress_large(n_traj=min(n_traj, 2), seed_offset=seed_offset, pin=pin)
  raise SystemExit(f"unknown workload {name}")


def atoms_in_focus_box(fz, traj):
  """Atoms inside the focus box, from the same crop the kernels see (mean over the first three windows)."""
  from nearl_b200 import features as _features
  fz.traj = traj
  fz.parse_focus()
  f0 = fz.FEATURESPACE[0]
  f0.cache(traj)
  return int(np.mean([np.count_nonzero(_features.Feature.query(f0, traj.top, traj.xyz[i * fz.time_window], fz.FOCALPOINTS[i, 0])[0])
                      for i in range(min(3, fz.SLICENUMBER))]))


def workload_config(fz, trajs, workload, n_rot=N_ROTATE, by_task=False):
  """The `config` object of the JSON line: what is featurised, identical for the native and the reference arm."""
  traj = trajs[0]
  n_frames = traj.n_frames
  total = n_rot * traj.xyz.nbytes
  return {"workload": workload, "tasks_per_step": fz.FEATURENUMBER * (n_frames // fz.time_window), "frames_per_step": n_frames,
          "atoms": traj.n_atoms, "atoms_in_focus_box": atoms_in_focus_box(fz, traj), "grid": [int(v) for v in fz.FEATURESPACE[0].dims],
          "l2": f"inputs larger than L2: {n_rot} trajectories x {traj.xyz.nbytes / 1e6:.1f} MB cycled" if total > 126e6
                else f"{n_rot} trajectories x {traj.xyz.nbytes / 1e6:.2f} MB cycled (fits L2; extra workload, not the headline)",
          "sharding": "one process per GPU, every trajectory split by (window, focal point, grid slab) over the ranks, parts gathered "
                      "by rank 0; no collective on the data path" if by_task
                      else "one process per GPU, trajectories sharded by rank, no collective on the data path"}


class ClockSampler:
  """SM clock and throttle reasons sampled DURING a timed region: NVML from a thread every 10 ms
  (the timed regions are tens of milliseconds), `nvidia-smi -lms` as the fallback."""
  FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
            "clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index = index
    self.proc = None
    self.lines = []
    self.samples = []          # (sm_mhz, reasons bitmask)
    self.max_mhz = None
    self.thread = None
    self.stop = threading.Event()
    self.nvml = None

  def _nvml_loop(self, handle):
    nv = self.nvml
    while not self.stop.is_set():
      try:
        self.samples.append((float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)),
                             int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                             if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons")
                             else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle))))
      except Exception:
        break
      self.stop.wait(0.01)

  def __enter__(self):
    try:
      import pynvml as nv
      nv.nvmlInit()
      # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
      vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
      phys = self.index
      if vis and all(tok.strip().isdigit() for tok in vis.split(",")):
        phys = int(vis.split(",")[self.index])
      handle = nv.nvmlDeviceGetHandleByIndex(phys)
      self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
      self.nvml = nv
      self.thread = threading.Thread(target=self._nvml_loop, args=(handle,), daemon=True)
      self.thread.start()
      return self
    except Exception:
      self.nvml = None
    try:
      self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                    "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
      threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
    except Exception:
      self.proc = None
    return self

  def __exit__(self, *exc):
    self.stop.set()
    if self.thread is not None:
      self.thread.join(timeout=1)
    if self.proc is not None:
      time.sleep(0.25)
      self.proc.terminate()
      try:
        self.proc.wait(timeout=2)
      except Exception:
        pass

  def summary(self):
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    if self.nvml is not None and self.samples:
      nv = self.nvml
      bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
              "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
              "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
              "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
      reasons = sorted({name for _, mask in self.samples for name, bit in bits.items() if mask & bit})
      return {"sm_mhz": float(np.median([m for m, _ in self.samples])), "sm_max_mhz": self.max_mhz, "reasons": reasons,
              "samples": len(self.samples), "source": "nvml, 10 ms period"}
    sm, mx, reasons = [], [], set()
    for line in self.lines:
      parts = [p.strip() for p in line.split(",")]
      if len(parts) < 8:
        continue
      try:
        sm.append(float(parts[0])); mx.append(float(parts[1]))
      except ValueError:
        continue
      for name, val in zip(names, parts[4:8]):
        if val.lower().startswith("active"):
          reasons.add(name)
    if not sm:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
            "source": "nvidia-smi -lms 200"}


def dist_env():
  return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def run_native(args):
  import torch
  import torch.distributed as dist
  import nearl_b200
  from nearl_b200 import _lib, workloads
  rank, local_rank, world = dist_env()
  nearl_b200.update_config(quiet=True)
  if not torch.cuda.is_available():
    raise SystemExit("bench.py: no CUDA device -- the native arm has no CPU fallback")
  torch.cuda.set_device(local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  stream = torch.cuda.current_stream()
  _lib.check(_lib.lib().nrl_set_device(local_rank))
  L = _lib.lib()
  import ctypes
  stream_ptr = ctypes.c_void_p(stream.cuda_stream)
  n_lanes = max(1, min(int(args.lanes), 8))

  # every rank featurises its own trajectories: weak scaling, no data-path collective
  global WORKLOAD
  WORKLOAD = WORKLOADS[args.workload]
  # --shard trajectories (default): every rank featurises its own trajectories, weak scaling, no data-path collective
  # --shard tasks: every rank holds the SAME trajectories and computes its share of each one's (window, focal point[, grid
  #                slab]) parts, rank 0 assembles the grids (sharding.run_task_sharded): strong scaling of ONE system
  by_task = args.shard == "tasks" and world > 1
  fz, trajs = make_workload(args.workload, N_ROTATE, seed_offset=0 if by_task else rank * N_ROTATE)
  fz.keep_results = True
  n_frames = trajs[0].n_frames
  n_tasks = fz.FEATURENUMBER * (n_frames // fz.time_window)

  # ---- device-resident arm: lowered batches + trajectories in HBM -------------------------------
  lowered, dev_xyz = [], []
  for tid, traj in enumerate(trajs):
    fz.traj = traj
    assert fz.parse_focus()
    for feat in fz.FEATURESPACE:
      feat.cache(traj)
    specs = [f._native_spec() for f in fz.FEATURESPACE]
    task_list = None
    if by_task:
      keys, slabs = fz.shard_task_list((rank, world, None), specs)
      task_list = tuple(k + sl for k, sl in zip(keys, slabs))
    lowered.append(fz.lower_batch(list(range(fz.FEATURENUMBER)), specs, traj.n_atoms, False, task_list=task_list))
    dev_xyz.append(torch.from_numpy(traj.xyz).cuda())
  # The steps are independent trajectories: they are issued round-robin over `n_lanes` lanes of the library (a lane = a
  # stream + a workspace), so the crop / binning / kernel tails of one step run under the sweeps of another, exactly as
  # Featurizer.run() pipelines them.  The lanes are forked off torch's stream after the start event and joined back
  # into it before the stop event: the CUDA events on torch's stream bracket every kernel of the K steps.
  dev_out = [torch.empty(max(max(b.out_floats for b in lowered), 1), dtype=torch.float32, device="cuda") for _ in range(n_lanes)]

  n_rot = len(trajs)

  def step_device(i):
    k, lane = i % n_rot, i % n_lanes
    if lowered[k].n_tasks_total == 0:
      return
    lowered[k].execute(dev_xyz[k].data_ptr(), n_frames, dev_out[lane].data_ptr(), xyz_on_device=True, out_on_device=True,
                       no_sync=True, lane=lane)

  def lanes_fork():
    for lane in range(n_lanes):
      _lib.check(L.nrl_lane_fork(lane, stream_ptr))

  def lanes_join():
    for lane in range(n_lanes):
      _lib.check(L.nrl_lane_join(lane, stream_ptr))

  # e2e: the public call, Featurizer.run() over a list of host trajectories (it pipelines them over
  # `pipeline_depth` lanes internally); every result grid is read on the host by the sink
  touched = [0.0]

  def sink(outkey, tid, bid, pid, grid):
    touched[0] += float(grid[0, 0, 0]) + float(grid[-1, -1, -1])
  fz.keep_results = False
  fz.result_sink = sink

  def run_e2e(first, count):
    indices = [(first + i) % n_rot for i in range(count)]
    if by_task:
      from nearl_b200 import sharding
      for c0 in range(0, count, 16):          # parts are gathered per chunk of trajectories (bounded memory on rank 0)
        sharding.run_task_sharded(fz, rank=rank, world=world, device=local_rank, trajectory_indices=indices[c0:c0 + 16])
    else:
      fz.run(trajectory_indices=indices)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(step, steps, warmup, profile=False, whole=None):
    if whole is not None:
      # the public call pipelines / groups its trajectories: warm it up with a run of the timed length, so that the
      # page-locked staging buffers and device arenas it grows on the way exist before the clock starts
      whole(0, max(warmup, min(steps, 128)))
    if whole is None:
      for i in range(n_lanes):       # every lane grows its workspace before the W warm-up steps (untimed)
        step(i)
      for i in range(warmup):
        step(i)
      lanes_join()
    barrier()
    if profile:
      _lib.profile(True)
      _lib.profile_read(reset=True)
    launches0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
      e0.record(stream)
      t_host = time.perf_counter()
      if whole is not None:
        whole(warmup, steps)
      else:
        lanes_fork()
        for i in range(steps):
          step(warmup + i)
        lanes_join()
      e1.record(stream)
      barrier()
      t_host = (time.perf_counter() - t_host) * 1e3
    # the e2e arm works on the library's own lanes (not torch's stream): its span is the host wall
    # clock of the bracketed region, which contains every copy and kernel (the call returns after
    # the last result has landed in host memory)
    ms = e0.elapsed_time(e1) if whole is None else t_host
    prof = _lib.profile_read(reset=True) if profile else {}
    if profile:
      _lib.profile(False)
    launches = _lib.kernel_launches() - launches0
    if world > 1:
      t = torch.tensor([ms], device="cuda")
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms, launches, prof, clk.summary()

  ms_dev, launches, prof, clocks = timed(step_device, args.steps, args.warmup, profile=True)
  ms_e2e, _, _, clocks_e2e = timed(None, args.steps, args.warmup, whole=run_e2e)

  # untimed extra pass: the same steps one after the other on ONE lane with the per-mode launches not forked onto side
  # streams, so that the event span of a launch is the kernel alone (in the timed region the spans of concurrent launches
  # overlap: they add up to more than the step)
  prof_alone, ms_alone = {}, None
  if rank == 0:
    _lib.set_option("no_fork", 1)
    try:
      def serial_steps(count):
        for i in range(count):
          k = i % n_rot
          if lowered[k].n_tasks_total:
            lowered[k].execute(dev_xyz[k].data_ptr(), n_frames, dev_out[0].data_ptr(), xyz_on_device=True, out_on_device=True,
                               no_sync=True, lane=0)
        _lib.check(L.nrl_lane_synchronize(0))
      serial_steps(3)
      _lib.profile(True)
      _lib.profile_read(reset=True)
      n_alone = min(args.steps, 16)
      t0 = time.perf_counter()
      serial_steps(n_alone)
      ms_alone = (time.perf_counter() - t0) * 1e3 / n_alone
      prof_alone = _lib.profile_read(reset=True)
      _lib.profile(False)
    finally:
      _lib.set_option("no_fork", 0)

  # untimed probe of the host link with this step's own transfers (rank 0): one trajectory up from its page-locked array,
  # one step's result grids down -- the floor the copies alone put under an end-to-end step
  link = None
  if rank == 0:
    try:
      host_xyz = torch.from_numpy(trajs[0].xyz)
      host_out = torch.empty(max(n_tasks * int(np.prod(fz.FEATURESPACE[0].dims)), 1), dtype=torch.float32).pin_memory()
      dev_res = torch.empty_like(host_out, device="cuda")
      def copy_ms(dst, src, reps=10):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(reps):
          dst.copy_(src, non_blocking=True)
        c1.record(stream)
        torch.cuda.synchronize()
        return c0.elapsed_time(c1) / reps
      up_ms, down_ms = copy_ms(dev_xyz[0], host_xyz), copy_ms(host_out, dev_res)
      link = {"h2d_gbs": host_xyz.numel() * host_xyz.element_size() / up_ms / 1e6, "h2d_ms_per_step": up_ms,
              "d2h_gbs": host_out.numel() * 4 / down_ms / 1e6, "d2h_ms_per_step": down_ms, "page_locked_source": bool(host_xyz.is_pinned()),
              "note": "untimed probe after the runs: cudaMemcpyAsync of one trajectory / one step's grids alone on the link; "
                      "the two directions overlap in the pipelined run, so the larger one is the floor of an end-to-end step"}
      del host_out, dev_res
    except Exception as exc:      # the probe is evidence, not part of the measurement
      link = {"error": str(exc)}

  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  frames_total = n_frames * (1 if by_task else world) * args.steps
  value = frames_total / (ms_dev * 1e-3)
  e2e = frames_total / (ms_e2e * 1e-3)

  # ---- roofline of the dominant kernel (per-launch CUDA-event times collected over the timed region)
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:
    pass
  hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
  peak_kind = "measured (MEASURED_PEAKS.json, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
  kernels = {k: {"ms_total": v[0], "launches": v[1], "ms_per_launch": v[0] / max(v[1], 1)} for k, v in prof.items()}
  top = max(prof.items(), key=lambda kv: kv[1][0])[0] if prof else None
  roof = None
  config = workload_config(fz, trajs, WORKLOAD, n_rot, by_task)
  n_in = config["atoms_in_focus_box"]
  f0 = fz.FEATURESPACE[0]
  from nearl_b200 import features as _features
  G = int(np.prod(f0.dims))
  # ncu --set full capture of the same kernels on the same workload (profiles/kernel_table.json, written by
  # scripts/ncu_kernel_table.py from the committed capture): DRAM traffic and warp instructions per launch
  table = {}
  try:
    table = json.load(open(os.path.join(ROOT, "profiles", "kernel_table.json"))).get("kernels", {}) if args.workload == "dynamic" else {}
  except Exception:
    pass
  if top is not None:
    t_launch = kernels[top]["ms_per_launch"] * 1e-3
    windows = n_frames // fz.time_window
    # algorithmic bytes of one (feature, window) task = F*N*16 (packed atoms read) + 4*G (grid written);
    # one launch of a tile kernel covers the `windows` tasks of its mode -- of every observer it fuses, for the fused sweep
    n_fused = sum(isinstance(f, _features.MarchingObservers) for f in fz.FEATURESPACE) if top == "tile_fused" else 1
    alg_bytes = windows * max(n_fused, 1) * (fz.time_window * n_in * 16 + 4 * G)
    achieved = alg_bytes / t_launch / 1e9
    alone = prof_alone.get(top)
    alone_ms = alone[0] / max(alone[1], 1) if alone else None
    alone_sum = sum(v[0] / max(v[1], 1) for v in prof_alone.values()) if prof_alone else None
    roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": table.get(top, {}).get("dram_bytes"), "peak_source": peak_kind, "algorithmic_bytes_per_launch": alg_bytes,
            "launch_ms": kernels[top]["ms_per_launch"],
            # the same kernel alone (untimed extra pass: one lane, launches not forked): its span, its share of the sum of the
            # spans of a step's kernels -- the figure the ncu launch list (profiles/*_launches.csv) shows --, the bytes over it
            "launch_ms_alone": alone_ms, "share_of_step_alone": (alone_ms / alone_sum) if alone_ms and alone_sum else None,
            "achieved_alone": (alg_bytes / (alone_ms * 1e-3) / 1e9) if alone_ms else None, "ms_per_step_alone": ms_alone,
            "note": "the gather kernels are FP32-issue bound, not HBM bound (see roofline_compute); the per-mode launches of a "
                    "step run concurrently on side streams, so a launch's event span includes the time it shares the SMs"}
    # compute side: voxel x candidate pair tests per second against the FP32 issue rate
    sm_mhz = clocks.get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
    pair_tests = windows * fz.time_window * G * n_in     # upper bound actually swept: every voxel x in-box atom would be brute force
    issue_peak = 148 * 4 * sm_mhz * 1e6               # warp instructions / s: 4 schedulers per SM, one issue per clock
    step_inst = sum(table.get(k, {}).get("warp_instructions", 0.0) * v["launches"] for k, v in kernels.items()) / max(args.steps, 1)
    roof_c = {"bound": "fp32-issue", "kernel": top, "brute_force_pairs_per_launch": pair_tests,
              "brute_force_pairs_per_s": pair_tests / t_launch, "sm_mhz": sm_mhz,
              "fp32_lane_ops_per_s_peak": 148 * 128 * sm_mhz * 1e6,
              "issue_peak_warp_inst_per_s": issue_peak,
              # whole step: warp instructions of all its kernels (ncu) / device time of the step / issue peak
              "step_warp_instructions": step_inst or None,
              "step_issue_frac": (step_inst / (ms_dev / args.steps * 1e-3) / issue_peak) if step_inst else None,
              "kernel_issue_active_pct_ncu": table.get(top, {}).get("issue_active_pct")}
    # algorithmic pair tests of one observer task = F * G * n_c with n_c = rho * (4/3) pi c^3 atoms truly
    # in range of a voxel, 9 flop each; a fused launch sweeps once for all observer channels of its work items
    n_obs = sum(isinstance(f, _features.MarchingObservers) for f in fz.FEATURESPACE)
    if top.startswith("tile_") and top != "tile_voxelize" and n_obs:
      c = float(f0.cutoff)
      box = float(np.prod(np.asarray(f0.dims) * f0.spacing + 2 * c))
      n_c = n_in / box * (4.0 / 3.0) * np.pi * c ** 3
      per_task = fz.time_window * G * n_c * 9.0
      tasks_per_launch = windows * (n_obs if top == "tile_fused" else 1)
      fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6
      roof_c.update({"algorithmic_flop_per_task": per_task, "atoms_in_range_per_voxel": n_c,
                     "algorithmic_flop_per_launch": per_task * tasks_per_launch,
                     "achieved_tflops": per_task * tasks_per_launch / t_launch / 1e12,
                     "fp32_peak_tflops": fp32_peak / 1e12,
                     "frac_of_fp32_peak": per_task * tasks_per_launch / t_launch / fp32_peak,
                     "note": "launch_ms is the launch's span while it shares the SMs with the voxeliser launch of the same step; "
                             "alone (ncu, profiles/kernel_table.json) the fused sweep takes %.0f us" % table.get(top, {}).get("duration_us", float("nan"))})
  else:
    roof_c = None

  # ---- CPU baseline: oracle port on the host cores, bounded sample of the same workload -----------
  if args.workload == "stress":
    cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
           "sample": "skipped: one window of this workload is ~2e12 brute-force pair tests per feature on the CPU oracle (hours)"}
  else:
    cpu = cpu_baseline_sample(fz, trajs) if world == 1 else None      # rank 0 at N = 1 only

  line = {
    "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
    "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if by_task else "weak", "vs_baseline": None,
    "dtype": "f32", "data": "synthetic",
    "config": config,
    "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
            "h2d_bytes_per_step": int(n_frames * trajs[0].n_atoms * 12 + 4 * trajs[0].n_atoms * 4),
            "d2h_bytes_per_step": int(n_tasks * G * 4),
            "api": "nearl_b200.featurizer.Featurizer.run() over K pinned host trajectories -> host result grids (read by a sink)",
            "timer": "host wall clock around the K-step call, barrier + synchronize on both sides, max over ranks",
            "link": link},
    "gpu_launches": int(launches),
    "clocks": clocks, "clocks_e2e": clocks_e2e,
    "roofline": roof, "roofline_compute": roof_c, "kernels": kernels,
    "kernels_alone": {k: {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1]} for k, v in prof_alone.items()},
    "cpu_baseline": cpu,
  }
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def cpu_baseline_sample(fz, trajs, budget_s=10.0):
  """Oracle port (CPU restatement of the reference kernels + numpy host pipeline) on the host cores: whole trajectories of
  the workload, one after the other, until about `budget_s` seconds of CPU work are spent (at least one, at most all)."""
  from nearl_b200 import workloads
  from oracle import host_ref
  cores = os.cpu_count() or 1
  os.environ.setdefault("OMP_NUM_THREADS", str(cores))
  frames, dt, n_done = 0, 0.0, 0
  for traj in trajs:
    fz.traj = traj
    fz.parse_focus()
    for f in fz.FEATURESPACE:
      f.cache(traj)
    descs = workloads.describe(fz)
    t0 = time.perf_counter()
    host_ref.featurize(traj.xyz, fz.FEATURESPACE[0].resids, descs, fz.FOCALPOINTS, fz.time_window)
    dt += time.perf_counter() - t0
    frames += fz.SLICENUMBER * fz.time_window
    n_done += 1
    if dt >= budget_s:
      break
  return {"value": frames / dt, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{n_done} whole trajector{'y' if n_done == 1 else 'ies'} ({frames} frames, {fz.FEATURENUMBER} features each) of the same "
                    f"workload, brute-force oracle (oracle/nearl_oracle.c, OpenMP over voxels) in {dt:.2f} s"}


def run_reference(args):
  rank, local_rank, world = dist_env()
  if rank != 0:
    return
  import nearl_b200
  from nearl_b200 import workloads
  from oracle import host_ref
  nearl_b200.update_config(quiet=True)
  global WORKLOAD
  WORKLOAD = WORKLOADS[args.workload]
  # ordinary pageable trajectories, as a Nearl user has them: page-locked host memory in the process makes the
  # reference's per-call cudaMalloc/cudaFree several times slower and would bias the comparison against it
  # (the same N_ROTATE seeded trajectories as rank 0 of the native arm, so that `config` is the same object)
  fz, trajs = make_workload(args.workload, N_ROTATE, pin=False)
  kernels, kind, device = None, "port", f"{os.cpu_count()} host cores (CPU oracle port)"
  warm_ctx = False
  if not args.cold_context:
    # A live torch allocation keeps the driver's sub-allocator warm: the reference cudaMalloc's and
    # cudaFree's 4-6 buffers per call, which is ~10x cheaper in a context that already holds memory
    # (a Nearl process that also trains) than in a bare one.  The favourable case is the default.
    try:
      import torch
      if torch.cuda.is_available():
        _keep = torch.zeros(1 << 20, device="cuda")
        torch.cuda.synchronize()
        warm_ctx = True
    except Exception:
      pass
  try:
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import all_actions as ref_actions
    probe = ref_actions.frame_voxelize(np.full((4, 3), 4.0, dtype=np.float32), np.ones(4, dtype=np.float32),
                                       np.array([8, 8, 8]), 1.0, 2.0, 1.0, 0)
    if np.any(probe):     # without a GPU the reference silently returns zeros
      kernels, kind = host_ref.ModuleKernels(ref_actions), "reference"
      device = "reference CUDA kernels (unmodified /root/reference/src, sm_100 build) on GPU 0 + 1 host thread for its per-task numpy pipeline"
  except Exception as exc:
    device += f" [oracle/_ref unavailable: {exc}]"
  n_frames = trajs[0].n_frames
  per_traj = []
  for traj in trajs:
    fz.traj = traj
    fz.parse_focus()
    for f in fz.FEATURESPACE:
      f.cache(traj)
    per_traj.append((traj, fz.FOCALPOINTS.copy(), workloads.describe(fz), fz.FEATURESPACE[0].resids))
  # the CPU port is ~100x slower: bound its sample to one window per step
  max_slices = None if kind == "reference" else 1

  def step(i):
    traj, focal, descs, resids = per_traj[i % len(per_traj)]
    host_ref.featurize(traj.xyz, resids, descs, focal, fz.time_window, kernels=kernels, max_slices=max_slices)

  for i in range(args.warmup):
    step(i)
  t0 = time.perf_counter()
  for i in range(args.steps):
    step(args.warmup + i)
  dt = time.perf_counter() - t0
  frames = (n_frames if max_slices is None else max_slices * fz.time_window) * args.steps
  value = frames / dt
  line = {
    "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
    "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": workload_config(fz, trajs, WORKLOAD, len(trajs)),
    "reference_arm": {"device": device,
                      "note": "the reference has no multi-GPU path: one process (rank 0) on one GPU whatever --gpus says",
                      "cuda_context": "warm (torch allocation alive: cheap cudaMalloc/cudaFree)" if warm_ctx
                                      else "bare (as a stand-alone Nearl process)"},
    "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1 if kind == "reference" else (os.cpu_count() or 1), "kind": kind,
                     "sample": ("whole trajectory (40 tasks) per step" if max_slices is None else "1 window (4 tasks) per step")},
    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line))


def main():
  # libraries (NCCL's version banner ...) print to stdout: everything but the result line goes to stderr
  real_stdout = os.dup(1)
  os.dup2(2, 1)
  sys.stdout = os.fdopen(real_stdout, "w")
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=100)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="native", choices=["native", "reference"])
  ap.add_argument("--workload", default="dynamic", choices=sorted(WORKLOADS), help="default: the headline configs[1]")
  ap.add_argument("--lanes", type=int, default=6, help="native arm, device-resident steps: lanes of the library the steps are spread over")
  ap.add_argument("--shard", default="trajectories", choices=["trajectories", "tasks"],
                  help="multi-GPU partition: whole trajectories per rank (weak scaling, the headline) or the (window, focal point, "
                       "grid slab) parts of every trajectory (strong scaling of one system, e.g. --workload stress)")
  ap.add_argument("--cold-context", action="store_true", help="reference arm: do not pre-create a torch CUDA context")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
  if args.workload == "stress" and "--steps" not in " ".join(sys.argv):
    args.steps = 10 if args.impl == "native" else 1          # a step is 50 ms here, ~10 s for the reference arm
  if args.impl == "reference":
    if args.steps == 100 and "--steps" not in " ".join(sys.argv):
      args.steps = 6                      # a reference step is ~0.5 s
    run_reference(args)
  else:
    run_native(args)
  sys.stdout.flush()


if __name__ == "__main__":
  main()
