#!/usr/bin/env python
"""bench.py - batched env-steps/s of the GymD2D step path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W                       # this repo's CUDA path
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # N > 1 (contiguous env slices)
    python bench.py --impl reference --steps K --warmup W               # the reference's CPU env.step loop

A "step" is one env.step of the whole batch: ONE fused kernel launch over E envs per GPU (decode ->
per-RB interference -> SINR/SNR -> Shannon capacity -> observation table -> reward -> done).  The workload
is BASELINE.json configs[1]: 4096 default-config (25 RB / 25 CUE / 25 DUE-pair) envs per GPU, positions
resident in HBM, pre-generated uniform random RB/power actions.  Weak scaling: every rank owns
--envs-per-gpu envs.  Prints ONE JSON line on rank 0 (see the keys below and DESIGN.md section 5).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = 'batched_env_steps_per_sec'
UNIT = 'env-steps/s'
RING = 32           # distinct action / output buffer sets cycled through, so per-step I/O never sits in L2


def algorithmic_bytes_per_env_step(N: int, V: int) -> int:
    """read actions 4N + positions 8V; write obs 24N + capacity 4N + reward 4 + done 1."""
    return 32 * N + 8 * V + 5


def hbm_peak():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        try:
            return float(json.loads(f.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------------------
# clocks: sampled through NVML in a background thread while the benchmark runs
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def __init__(self, index: int) -> None:
        self.samples = []           # (t, sm_mhz, reasons_bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _run(self) -> None:
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))))
            except Exception:  # noqa: BLE001
                try:
                    self.samples.append((time.perf_counter(), int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                         int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))))
                except Exception:  # noqa: BLE001
                    break
            time.sleep(0.002)

    def start(self) -> None:
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(1.0)

    def summary(self, windows) -> dict:
        """Median SM clock over the samples that fall inside the timed windows [(t0, t1), ...]."""
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': self.sm_max, 'reasons': [], 'samples': 0}
        inside = [s for s in self.samples if any(t0 <= s[0] <= t1 for t0, t1 in windows)]
        note = 'timed regions'
        if len(inside) < 3:     # the timed region is shorter than the NVML polling period: fall back to every
            inside = [s for s in self.samples if not (s[2] & 0x1)] or self.samples   # sample taken under load
            note = 'whole bench (timed region shorter than the NVML period)'
        clocks = sorted(s[1] for s in inside)
        mask = 0
        for s in inside:
            mask |= s[2]
        reasons = [name for bit, name in self.REASONS.items() if mask & bit and name != 'gpu_idle']
        return {'sm_mhz': clocks[len(clocks) // 2], 'sm_max_mhz': self.sm_max, 'reasons': reasons,
                'samples': len(inside), 'window': note}


# ------------------------------------------------------------------------------------------------------------
# CPU arms: the unmodified reference (baseline/_ref, kind "reference") or the C oracle (kind "port")
# ------------------------------------------------------------------------------------------------------------
_REF_ENV = None


def _ref_worker_init(seed: int) -> None:
    global _REF_ENV
    import random
    sys.path.insert(0, str(ROOT))
    from oracle import ref_runner as R
    random.seed(seed + os.getpid())
    _REF_ENV = R.make_env({})
    _REF_ENV.reset()


def _ref_worker_steps(m: int) -> int:
    """m env.step calls on the reference's default D2DEnv with pre-sampled random actions."""
    env = _REF_ENV
    cue_keys = [k for k in env.actions.keys()]   # (tx, rx) id pairs in canonical order
    keys = [':'.join(k) for k in cue_keys]
    import random
    n_cue = env.action_space['cue'].n
    n_due = env.action_space['due'].n
    acts = [{k: random.randrange(n_cue if k.startswith('cue') else n_due) for k in keys} for _ in range(m)]
    t0 = time.perf_counter()
    for a in acts:
        env.step(a)
    return m if time.perf_counter() - t0 >= 0 else 0


class CpuArm:
    """Times the reference's own CPU implementation of the step path on all host cores."""

    def __init__(self) -> None:
        from oracle import ref_runner as R
        self.cores = os.cpu_count() or 1
        self.kind = 'reference' if R.reference_src() is not None else 'port'
        self.pool = None
        if self.kind == 'reference':
            import multiprocessing as mp
            ctx = mp.get_context('spawn')
            self.pool = ctx.Pool(self.cores, initializer=_ref_worker_init, initargs=(1234,))
            self.pool.map(_ref_worker_steps, [2] * self.cores)       # spin the workers up
        else:
            import numpy as np
            from oracle import d2d_oracle as O
            self.O, self.np = O, np
            self.cfg = O.OracleConfig()
            self.rng = np.random.default_rng(0)

    def run(self, env_steps_per_core: int) -> float:
        """Processes cores * env_steps_per_core env-steps of the default config; returns seconds."""
        t0 = time.perf_counter()
        if self.kind == 'reference':
            self.pool.map(_ref_worker_steps, [env_steps_per_core] * self.cores)
        else:
            E = env_steps_per_core * self.cores
            pos = self.O.random_positions(self.cfg, E, self.rng)
            act = self.O.random_actions(self.cfg, E, self.rng)
            t0 = time.perf_counter()
            self.O.step_batch(self.cfg, pos, act, nthreads=self.cores)
        return time.perf_counter() - t0

    def close(self) -> None:
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def cpu_baseline(budget_s: float = 12.0) -> dict:
    arm = CpuArm()
    per_core = 256 if arm.kind == 'reference' else 4096
    t = arm.run(per_core)
    reps = max(1, int(budget_s / max(t, 1e-3)) - 1)
    total_t, total_n = 0.0, 0
    for _ in range(reps):
        total_t += arm.run(per_core)
        total_n += per_core * arm.cores
    out = {'value': total_n / total_t, 'unit': UNIT, 'cores': arm.cores, 'kind': arm.kind,
           'sample': f'{total_n} env-steps of the default 25/25/25 config ({arm.cores} processes x {reps} x {per_core} '
                     f'env.step calls, one D2DEnv each, random actions) in {total_t:.1f} s'}
    arm.close()
    if arm.kind == 'reference':      # also time the compiled float64 restatement, for context
        try:
            import numpy as np
            from oracle import d2d_oracle as O
            cfg = O.OracleConfig()
            rng = np.random.default_rng(0)
            E = 65536
            pos, act = O.random_positions(cfg, E, rng), O.random_actions(cfg, E, rng)
            O.step_batch(cfg, pos[:1024], act[:1024], nthreads=arm.cores)
            t0 = time.perf_counter()
            O.step_batch(cfg, pos, act, nthreads=arm.cores)
            out['port'] = {'value': E / (time.perf_counter() - t0), 'unit': UNIT, 'cores': arm.cores,
                           'kind': 'port', 'sample': f'{E} env-steps, C oracle with OpenMP'}
        except Exception as exc:  # noqa: BLE001
            out['port'] = {'error': str(exc)}
    return out


def run_reference_arm(args) -> None:
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    arm = CpuArm()
    per_core = 128 if arm.kind == 'reference' else 8192
    for _ in range(args.warmup):
        arm.run(per_core)
    t = 0.0
    for _ in range(args.steps):
        t += arm.run(per_core)
    n = per_core * arm.cores * args.steps
    value = n / t
    sample = (f'each step = {arm.cores} x {per_core} env.step calls of the default 25/25/25 config '
              f'({"unmodified reference via baseline/_ref" if arm.kind == "reference" else "C oracle port, OpenMP"})')
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': 'D2DEnv-v0 defaults (25 RB / 25 CUE / 25 DUE pairs), random actions, CPU env.step loop',
                       'env_steps_per_step': per_core * arm.cores},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': arm.cores, 'kind': arm.kind, 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    arm.close()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------------------
def timed_steps(env, torch, actions, outs, steps, warmup, dist, use_graph=True):
    """W untimed + EXACTLY K timed steps; step i reads actions[i % RING] and writes outs[i % RING].  The K steps are issued as
    replays of ONE CUDA graph of G = min(K, RING) step kernels (+ K mod G eager launches): what a throughput-minded caller
    does, and independent of how busy the host's Python threads are.  The action buffers are pre-generated, so the steps
    carry D2D_STEP_INPUTS_STABLE (include/d2d_b200.h, "Ordering rule").
    Returns (seconds [max over ranks], perf_counter window, launches, launch description)."""
    ring = len(actions)

    def eager(i0, n):
        for i in range(i0, i0 + n):
            env.step(actions[i % ring], out=outs[i % ring], inputs_stable=True)

    graph, G = None, 0
    if use_graph and steps >= 2:
        G = min(steps, ring)
        eager(0, ring)                                   # module load + first-touch before capture
        graph = env.capture_steps([actions[i % ring] for i in range(G)], [outs[i % ring] for i in range(G)], inputs_stable=True)
    eager(0, warmup)
    if graph is not None:
        graph.replay()                                   # untimed: uploads the graph
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = env.launch_count
    t0 = time.perf_counter()
    # keep the device busy for ~0.3 ms while the host enqueues the timed region, so that the window between the two events holds
    # the K steps back to back and not the host's launch latency after an idle device (it varied 5-20 us between ranks and boxes)
    torch.cuda._sleep(600_000)
    start.record()
    done = 0
    if graph is not None:
        for _ in range(steps // G):
            graph.replay()
        done = (steps // G) * G
    eager(done, steps - done)
    stop.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    secs = start.elapsed_time(stop) * 1e-3
    launches = (env.launch_count - l0) + done          # graph replays launch `done` step kernels
    if dist is not None:
        t = torch.tensor([secs], dtype=torch.float64, device=env.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
        dist.barrier()
    how = (f'{steps // G} replay(s) of a CUDA graph of {G} step kernels' + (f' + {steps - done} eager launches' if steps - done else '')
           if graph is not None else f'{steps} eager launches')
    return secs, (t0, t1), launches, how + ('; pre-generated actions (D2D_STEP_INPUTS_STABLE); consecutive steps write different output sets of the ring, '
                                            'so their per-link outputs are stored ahead of griddepcontrol.wait (late wait)')


def host_link_peak(torch, device, nbytes_in, nbytes_out, dist, iters=30):
    """Plain pinned-memory cudaMemcpyAsync of the end-to-end path's per-step sizes, both directions at once, all ranks at
    once: what the box's host link gives this process - the ceiling of any host-buffer API."""
    hin = torch.empty(nbytes_in, dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(nbytes_out, dtype=torch.uint8, pin_memory=True)
    din = torch.empty(nbytes_in, dtype=torch.uint8, device=device)
    dout = torch.empty(nbytes_out, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    for it in range(2):
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return {'d2h_gbs_per_gpu': nbytes_out * iters / dt / 1e9, 'h2d_gbs_per_gpu': nbytes_in * iters / dt / 1e9,
            'how': f'{iters} x (H2D {nbytes_in} B || D2H {nbytes_out} B) pinned cudaMemcpyAsync on two streams, every rank at once'}


def run_cuda_arm(args) -> None:
    import torch
    import torch.distributed as dist

    import gym_d2d_b200 as G
    from gym_d2d_b200.dist import all_reduce_stats, shard_range, summarise

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: gym_d2d_b200 has no CPU path (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    from gym_d2d_b200.dist import EpisodeStatsReducer, bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local)      # before any pinned allocation: host buffers next to this GPU's PCIe root
    pg = None
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        pg = dist
    n_gpus = world
    E = args.envs_per_gpu
    first, _ = shard_range(E * world, rank, world)

    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    def build(num_envs, seed=0, cfg=None, ring=RING):
        env = G.VecD2DEnv(num_envs, dict(cfg or {}), device=torch.device('cuda', local), seed=seed, global_env_offset=first)
        env.reset()
        gen = torch.Generator(device=env.device)
        gen.manual_seed(1000 + rank)
        acts = [env.sample_actions(gen) for _ in range(ring)]
        outs = [env.alloc_outputs() for _ in range(ring)]
        return env, acts, outs

    # ---- headline: configs[1], E envs per GPU --------------------------------------------------------------
    env, acts, outs = build(E)
    N, V = env.num_links, env.num_devices
    B = algorithmic_bytes_per_env_step(N, V)
    comm_stream = torch.cuda.Stream() if world > 1 else None
    secs, win, launches, launch_how = timed_steps(env, torch, acts, outs, args.steps, args.warmup, pg)
    windows.append(win)
    if world > 1:   # episode statistics: one tiny all-reduce, off the step stream (never inside the step)
        stats = env.stats_tensor()
        comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(comm_stream):
            all_reduce_stats(stats)
        torch.cuda.current_stream().wait_stream(comm_stream)
        stats_summary = summarise(stats)
    else:
        stats_summary = summarise(env.stats_tensor())
    value = world * E * args.steps / secs
    peak, peak_src = hbm_peak()
    per_launch_s = secs / args.steps
    achieved = B * E / per_launch_s / 1e9
    geometry = env.step_geometry()
    ring_bytes = RING * (acts[0].numel() * 4 + outs[0].nbytes())

    traffic = None
    prof = ROOT / 'profiles' / 'roofline_r02.json'
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get(f'dram_bytes_per_launch_E{E}')
        except Exception:  # noqa: BLE001
            traffic = None

    # ---- end to end through the C ABI with HOST buffers (d2d_step_host_async): H2D actions + D2H results per step ----
    # The per-step results a host loop needs are the columns that change every step: obs_dyn (sinr, snr) + capacity + reward +
    # done = 12 N + 5 bytes per env-step; the four position columns of the observation table move only on reset and are
    # fetched once (d2d_get_positions).  Outputs land in the library's packed pinned slot buffers: one copy per direction.
    e2e = None
    if not args.skip_e2e:
        DEPTH = 4                                  # steps in flight (D2D_HOST_SLOTS)
        slots = [env.host_slot_buffers(k) for k in range(DEPTH)]
        obs_static = env.obs_static()              # once per reset, outside the per-step loop (E x N x 16 B)
        # the actions go up as int16 (D2D_STEP_ACTIONS_I16: every Discrete space of the reference fits 15 bits; the library widens
        # them on the device): on a host link shared by several GPUs the download gets what the upload leaves
        host_acts = [torch.empty((E, N), dtype=torch.int16, pin_memory=True) for _ in range(4)]
        for h, a in zip(host_acts, acts):
            h.copy_(a.to(torch.int16))
        host_np = [h.numpy() for h in host_acts]
        h2d, d2h = E * N * 2, E * N * 8 + E * N * 4 + E * 4 + E
        e2e_steps = max(200, min(args.steps, 400))
        for i in range(2 * DEPTH):         # warm-up through every pipeline slot
            env.step_host_async(host_np[i % 4], slots[i % DEPTH], i % DEPTH)
            env.step_host_wait(i % DEPTH)
        # parity of the transport: the reassembled table equals the full-table device output of the same step
        env.step_count.zero_()
        env.step(acts[3], out=outs[0])
        torch.cuda.synchronize()
        import numpy as np
        assert np.array_equal(env.assemble_obs(obs_static, slots[DEPTH - 1]['obs_dyn']), outs[0].obs.cpu().numpy()), 'obs_dyn transport mismatch'
        link = host_link_peak(torch, env.device, h2d, d2h, pg)
        torch.cuda.synchronize()
        if pg is not None:
            pg.barrier()
        # DEPTH steps in flight: uploads and downloads overlap the kernels; every step's results reach the host.
        # The window is short and sees the host's other activity, so it is repeated three times and the MEDIAN window is
        # reported (all three are listed in `window_ms`).
        dts = []
        for rep in range(3):
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                if i >= DEPTH:
                    env.step_host_wait(i % DEPTH)
                env.step_host_async(host_np[i % 4], slots[i % DEPTH], i % DEPTH)
            for k in range(DEPTH):
                env.step_host_wait(k)
            t1 = time.perf_counter()
            windows.append((t0, t1))
            dt = t1 - t0
            if pg is not None:
                t = torch.tensor([dt], dtype=torch.float64, device=env.device)
                pg.all_reduce(t, op=pg.ReduceOp.MAX)
                dt = float(t.item())
            dts.append(dt)
        dt = sorted(dts)[1]
        e2e = {'value': world * E * e2e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
               'steps': e2e_steps, 'window_ms': [round(1e3 * x, 3) for x in dts],
               'host_link_gbs_per_gpu': (h2d + d2h) * e2e_steps / dt / 1e9,   # copies in both directions: the PCIe link bounds this number
               'd2h_gbs_per_gpu': d2h * e2e_steps / dt / 1e9, 'host_link_peak': link, 'numa': numa,
               'once_per_reset_bytes': int(obs_static.nbytes),
               'api': 'd2d_step_host_async/_wait via VecD2DEnv.step_host_async (pinned host buffers, four steps in flight; int16 actions '
                      'copied in (D2D_STEP_ACTIONS_I16); obs_dyn (sinr, snr) + capacity + reward + done copied back every step as ONE packed copy; the '
                      'position columns of the observation table are fetched once per reset with d2d_get_positions)'}
    # ---- fused rollouts: d2d_rollout, T counted steps of every env per launch with the actions sampled on the
    # device (Discrete.sample, envs/d2d_env.py:54-60) - no [T][E][N] action input at all; positions read once per launch ----
    fused = None
    if args.fused_steps > 1:
        T = args.fused_steps
        nbuf = max(2, min(8, (192 << 20) // (E * T * (B - 8 * V - 4 * N)) + 1))      # > L2 worth of outputs in flight
        many_outs = [env.alloc_many_outputs(T) for _ in range(nbuf)]
        for k, o_ in enumerate(many_outs):
            env.rollout(T, action_seed=7 + rank, first_step_index=1 + k * T, out=o_)
        torch.cuda.synchronize()
        if pg is not None:
            pg.barrier()
        reps = max(32, min(128, args.steps // T))
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for i in range(reps):
            env.rollout(T, action_seed=7 + rank, first_step_index=1 + i * T, out=many_outs[i % nbuf], inputs_stable=True)
        ev1.record()
        torch.cuda.synchronize()
        windows.append((t0, time.perf_counter()))
        secs_f = ev0.elapsed_time(ev1) * 1e-3
        if pg is not None:
            tt = torch.tensor([secs_f], dtype=torch.float64, device=env.device)
            pg.all_reduce(tt, op=pg.ReduceOp.MAX)
            secs_f = float(tt.item())
        BF = 28 * N + 5 + (8 * V + 1) / T                 # per env-step: outputs only; positions + counter once per launch
        fused = {'workload': f'd2d_rollout: {T} consecutive counted steps of the same {E} envs per launch, actions sampled on the device '
                             f'(no action input), positions read once; {reps} eager launches',
                 'value': world * E * T * reps / secs_f, 'unit': UNIT, 'steps_per_launch': T, 'us_per_step': 1e6 * secs_f / (reps * T),
                 'algorithmic_bytes_per_env_step': BF}
        del many_outs
    env.close()
    del env, acts, outs

    # ---- large batch: BASELINE configs[4]'s per-GPU slice (131072 envs, working set 290 MB > L2) ---------------
    large = None
    episode = None
    if args.large_envs_per_gpu > 0:
        envL, actsL, outsL = build(args.large_envs_per_gpu, seed=1, ring=8)
        stepsL = max(RING, min(args.steps, 4 * RING))
        secsL, winL, _, howL = timed_steps(envL, torch, actsL, outsL, stepsL, max(3, min(args.warmup, 8)), pg)
        windows.append(winL)
        EL = args.large_envs_per_gpu
        achL = B * EL / (secsL / stepsL) / 1e9
        large = {'workload': f'{EL} default-config envs per GPU (BASELINE configs[4] slice; working set > L2)',
                 'value': world * EL * stepsL / secsL, 'unit': UNIT, 'steps': stepsL, 'ms_per_step': 1e3 * secsL / stepsL, 'launch': howL,
                 'roofline': {'bound': 'hbm', 'achieved': achL, 'peak': peak, 'unit': 'GB/s', 'frac': achL / peak,
                              'traffic': None}}
        if prof.exists():
            try:
                large['roofline']['traffic'] = json.loads(prof.read_text()).get(f'dram_bytes_per_launch_E{EL}')
            except Exception:  # noqa: BLE001
                pass
        del actsL, outsL
        # ---- BASELINE configs[4] as specified: the EPISODE loop.  Per episode ONE d2d_episode launch - Simulator.reset (new
        # positions), the uncounted reset step and EPISODE_LENGTH counted steps with on-device sampled actions - and one
        # all-reduce of that episode's reward / capacity statistics over the ranks on a side stream.  Only the counted steps
        # count as env-steps; every slice's outputs (11 per episode) are written to HBM.
        if args.episodes > 0:
            T = 10
            red = EpisodeStatsReducer(envL, None)      # default process group (NCCL) when world > 1
            ep_outs = [envL.alloc_many_outputs(T + 1) for _ in range(2)]
            for i in range(2):
                red.begin_episode(); envL.episode(T, out=ep_outs[i & 1]); red.end_episode()
            red.finish()
            torch.cuda.synchronize()
            if pg is not None:
                pg.barrier()
            nr0 = red.all_reduces
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = envL.launch_count
            t0 = time.perf_counter()
            ev0.record()
            for i in range(args.episodes):
                red.begin_episode()
                envL.episode(T, out=ep_outs[i & 1])
                red.end_episode()
            red.finish()
            ev1.record()
            torch.cuda.synchronize()
            windows.append((t0, time.perf_counter()))
            secsE = ev0.elapsed_time(ev1) * 1e-3
            if pg is not None:
                tt = torch.tensor([secsE], dtype=torch.float64, device=envL.device)
                pg.all_reduce(tt, op=pg.ReduceOp.MAX)
                secsE = float(tt.item())
            last = summarise(red.results[-1])
            BE = (T + 1) * (32 * N - 4 * N + 5) + 8 * V + 1           # 11 slices of outputs (no action / position reads) + the positions written once
            episode = {'workload': f'{args.episodes} episodes of {EL} default-config envs per GPU: d2d_episode (device-side reset + uncounted reset step + '
                                   f'{T} counted steps, on-device sampled actions) in one launch, statistics all-reduced per episode on a side stream',
                       'value': world * EL * T * args.episodes / secsE, 'unit': UNIT, 'episodes': args.episodes,
                       'ms_per_episode': 1e3 * secsE / args.episodes, 'ms_per_counted_step': 1e3 * secsE / (args.episodes * T),
                       'vs_steps_only': (EL * T * args.episodes / secsE) / (EL * stepsL / secsL),
                       'launches_per_episode': (envL.launch_count - l0) / args.episodes,
                       'stats_allreduces': red.all_reduces - nr0, 'allreduce_world': world,
                       'last_episode_stats': {k: last[k] for k in ('env_steps', 'mean_reward', 'mean_capacity_mbps', 'penalties')},
                       'roofline': {'bound': 'hbm', 'algorithmic_bytes_per_env_episode': BE,
                                    'achieved': BE * EL * args.episodes / secsE / 1e9, 'peak': peak, 'unit': 'GB/s',
                                    'frac': BE * EL * args.episodes / secsE / 1e9 / peak, 'traffic': None}}
            if prof.exists():
                try:
                    episode['roofline']['traffic'] = json.loads(prof.read_text()).get(f'dram_bytes_per_launch_episode_E{EL}')
                except Exception:  # noqa: BLE001
                    pass
            del ep_outs
        envL.close()
        del envL

    # ---- dense cell: BASELINE configs[2] (100 RBs / 100 CUEs / 500 DUE pairs, FreeSpacePathLoss, 65536 envs: 1.8 GB per step) ----
    dense = None
    if args.dense_envs_per_gpu > 0:
        ED = args.dense_envs_per_gpu
        envD, actsD, outsD = build(ED, seed=2, cfg=dict(num_rbs=100, num_cues=100, num_due_pairs=500, path_loss_model=G.FreeSpacePathLoss), ring=4)
        BD = algorithmic_bytes_per_env_step(envD.num_links, envD.num_devices)
        stepsD = max(8, min(args.steps, 32))
        secsD, winD, _, howD = timed_steps(envD, torch, actsD, outsD, stepsD, 3, pg)
        windows.append(winD)
        achD = BD * ED / (secsD / stepsD) / 1e9
        gD = envD.step_geometry()
        dense = {'workload': f'BASELINE configs[2]: {ED} envs per GPU of the dense cell (100 RBs, 100 CUEs, 500 DUE pairs: N = 600 links, '
                             'V = 1101 devices), FreeSpacePathLoss; 4 action/output buffer sets of 1.3 GB',
                 'value': world * ED * stepsD / secsD, 'unit': UNIT, 'steps': stepsD, 'ms_per_step': 1e3 * secsD / stepsD,
                 'kernel': 'd2d_step_dense_kernel', 'launch': howD, 'grid': gD['grid'], 'block': gD['block'], 'smem_bytes': gD['smem_bytes'],
                 'roofline': {'bound': 'hbm', 'achieved': achD, 'peak': peak, 'unit': 'GB/s', 'frac': achD / peak,
                              'algorithmic_bytes_per_env_step': BD, 'traffic': None}}
        if prof.exists():
            try:
                dense['roofline']['traffic'] = json.loads(prof.read_text()).get(f'dram_bytes_per_launch_dense_E{ED}')
            except Exception:  # noqa: BLE001
                pass
        envD.close()
        del envD, actsD, outsD

    # ---- the reference's own dict API at E = 1 (BASELINE configs[0] through this repo): gym-style D2DEnv.step(dict) ----
    dict_api = None
    if rank == 0 and args.dict_steps > 0:
        import random as _random
        one = G.D2DEnv({}, device=torch.device('cuda', local))
        one.reset()
        keys = list(one.link_keys)
        n_of = {k: one.action_space['cue' if k.startswith('cue') else 'due'].n for k in keys}
        rnd = _random.Random(0)
        acts1 = [{k: rnd.randrange(n_of[k]) for k in keys} for _ in range(args.dict_steps + 20)]
        for a1 in acts1[:20]:
            one.step(a1)
        t0 = time.perf_counter()
        for a1 in acts1[20:]:
            one.step(a1)
        dt1 = time.perf_counter() - t0
        dict_api = {'value': args.dict_steps / dt1, 'unit': UNIT, 'ms_per_step': 1e3 * dt1 / args.dict_steps, 'steps': args.dict_steps,
                    'workload': "BASELINE configs[0] through this repo: gym_d2d_b200.D2DEnv({}).step(dict of 50 'tx:rx' -> int) at E = 1, "
                                'synchronous d2d_step_host per call (H2D actions, one kernel, D2H results), dict / per-agent obs building '
                                'on one host thread; compare cpu_baseline.value / cpu_baseline.cores (the reference per core)'}
        one.close()

    sampler.stop()
    clocks = sampler.summary(windows)

    base = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        base = cpu_baseline()

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * secs / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'BASELINE configs[1]: {E} batched default-config envs per GPU (25 RB / 25 CUE / 25 DUE pairs, '
                                   'LogDistancePathLoss), device-resident positions, uniform random RB/power actions',
                       'envs_per_gpu': E, 'num_links': N, 'num_devices': V, 'obs': 'compact [E][N][6] float32 table',
                       'parallelism': f'env-sharded x{world}', 'launch': launch_how,
                       'grid': geometry['grid'], 'block': geometry['block'], 'smem_bytes': geometry['smem_bytes'],
                       'l2': f'ring of {RING} action/output buffer sets ({ring_bytes / 1e6:.0f} MB > 126 MB L2) so per-step I/O is '
                             'never L2-resident; the {:.1f} MB position state stays L2-resident by design'.format(E * V * 8 / 1e6),
                       'stats_allreduce': 'once per run on a side stream' if world > 1 else 'none (1 GPU)'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'algorithmic_bytes_per_env_step': B, 'bytes_per_launch': B * E,
                         'peak_source': peak_src,
                         'note': 'duration = timed region / K launches (launch gaps included); E envs = '
                                 f'{E / (geometry["grid"] * geometry["envs_per_block"]):.2f} envs per resident warp slot'},
            'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks, 'episode_stats': stats_summary,
        }
        if fused is not None:
            line['fused_rollout'] = fused
        if large is not None:
            line['large_batch'] = large
        if episode is not None:
            line['episode_loop'] = episode
        if dict_api is not None:
            line['dict_api'] = dict_api
        if dense is not None:
            line['dense_cell'] = dense
        if base is not None:
            line['cpu_baseline'] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2048)
    ap.add_argument('--warmup', type=int, default=64)
    ap.add_argument('--impl', choices=['cuda', 'reference'], default='cuda')
    ap.add_argument('--envs-per-gpu', type=int, default=4096)
    ap.add_argument('--large-envs-per-gpu', type=int, default=131072)
    ap.add_argument('--dense-envs-per-gpu', type=int, default=65536, help='BASELINE configs[2] leg (0 = skip)')
    ap.add_argument('--fused-steps', type=int, default=10, help='T of the d2d_step_many leg (EPISODE_LENGTH); 0/1 skips it')
    ap.add_argument('--episodes', type=int, default=8, help='episodes of the episode_loop leg at --large-envs-per-gpu (0 = skip)')
    ap.add_argument('--dict-steps', type=int, default=300, help='D2DEnv.step calls of the dict_api leg (0 = skip)')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--skip-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        if args.steps == 2048 and args.warmup == 64:       # defaults sized for the GPU arm; keep the CPU arm to ~1 min
            args.steps, args.warmup = 20, 3
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == '__main__':
    main()
