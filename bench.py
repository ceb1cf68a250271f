#!/usr/bin/env python
"""bench.py — UR5PlayAbsRPY1Obj-v0 env-steps/s on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference arm: CPU oracle port on the host cores

A "step" is one env step of every environment of the shard: clip -> IK -> motor targets ->
12 x 300 Hz substeps (collision, dynamics, 50-iteration PGS) -> observation dict -> reward.
Envs are sharded by env index with no data-path collective (weak scaling: envs per GPU fixed);
NCCL only gathers episode statistics and the max-over-ranks time.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENV_ID = 'UR5PlayAbsRPY1Obj-v0'
RELABEL_EVERY, RELABEL_PHASE = 64, 2       # goal relabelling cadence and its phase in the timed region
# DRAM bytes (read + write) of the step pipeline per env step at 65536 envs, from the ncu --set full capture in
# profiles/r2_ncu.md (12 substeps x 0.63 GB + the final setup launch + IK); None for configurations not captured
MEASURED_TRAFFIC_BYTES = {('UR5PlayAbsRPY1Obj-v0', 65536): 7.5e9}
BYTES_PER_ENV_STEP = {'UR5Reach-v0': 416, 'pandaPick-v0': 560, 'UR5PlayAbsRPY1Obj-v0': 1044}
METRIC = 'UR5PlayAbsRPY1Obj-v0 env-steps/s'


# Interaction regions of the playroom scene (scenes.py:46-426; world frame of the compiled model).  The block sub-tasks
# grasp, lift and carry; the drawer sub-task drops the closed fingers into the slot behind the drawer's front plate, pulls
# it out against its stoppers (-y) and pushes it back; the door sub-task presses the closed fingers against the vertical
# bar of the door's ring handle, slides the door +x and back.  Button and dial gestures are traced next to the parts (the
# button sits under the cabinet shelf and the dial under the table edge: a top-down gripper cannot reach them).
_ANCHORS = {'drawer': (-0.131, -0.167, -0.055), 'door': (0.0, 0.322, 0.10), 'button': (-0.20, 0.33, 0.12), 'dial': (0.20, 0.10, 0.07)}
# sub-task scripts: waypoints (dx, dy, dz relative to the anchor, yaw, gripper, dwell steps)
_SCRIPTS = [
    ('block', [(0, 0, 0.15, 0, -1, 0), (0, 0, 0.03, 0, -1, 4)]),                                             # reach block
    ('block', [(0, 0, 0.15, 0, -1, 0), (0, 0, -0.01, 0, -1, 2), (0, 0, -0.01, 0, 1, 10), (0, 0, 0.2, 0, 1, 6),
               (0.05, 0.05, 0.2, 0, 1, 2), (0.05, 0.05, 0.03, 0, -1, 4)]),                                    # grasp, lift, carry, release
    ('drawer', [(0, 0, 0.12, 0, 1, 0), (0, 0, 0, 0, 1, 1), (0, -0.09, 0, 0, 1, 1), (0, 0.03, 0, 0, 1, 1),
                (0, 0, 0.12, 0, 1, 0)]),                                                                       # drawer: pull out, push back
    ('door', [(-0.07, 0, 0.10, 0, 1, 0), (-0.07, 0, 0, 0, 1, 1), (0.09, 0, 0, 0, 1, 2), (0.09, -0.08, 0, 0, 1, 0),
              (0.20, -0.08, 0, 0, 1, 0), (0.20, 0, 0, 0, 1, 0), (0.03, 0, 0, 0, 1, 2), (0.10, -0.08, 0.05, 0, 1, 0)]),   # door: slide +x, slide back
    ('button', [(0, 0, 0.06, 0, 1, 0), (0, 0, 0.0, 0, 1, 6), (0, 0, 0.06, 0, 1, 2)]),                           # button press gesture
    ('dial', [(0, 0, 0.06, 0, -1, 0), (0, 0, 0.0, 0, 1, 6), (0, 0, 0.0, 1.0, 1, 4)]),                           # dial turn gesture
]


def synth_actions(rng, n, steps, env_id, block_xyz=None, ee_xyz=None, jump_frac=0.0):
    """Scripted teleop-shaped synthetic actions: every env runs a random
    sub-task {reach block, grasp + lift, pull drawer, slide door, press button, turn dial} as
    piecewise-linear end-effector waypoints at <= 0.015 m / 0.1 rad per 25 Hz step, the gripper closing
    at contact, then draws the next sub-task.  jump_frac > 0 adds the Random stream's tail (that
    fraction of the steps targets a point of the full +-6 clip box: exercises the +-inc clamp and
    drives arms into the furniture)."""
    if env_id == 'UR5PlayAbsRPY1Obj-v0':
        lo, hi = np.array([-0.30, -0.05, 0.0]), np.array([0.30, 0.50, 0.35])
    else:
        lo, hi = np.array([-0.18, -0.18, -0.05]), np.array([0.18, 0.18, 0.2])
    play = env_id == 'UR5PlayAbsRPY1Obj-v0'
    blk = np.asarray(block_xyz, np.float64) if block_xyz is not None else rng.uniform(lo, hi, (n, 3))
    pos = np.asarray(ee_xyz, np.float64).copy() if ee_xyz is not None else rng.uniform(lo, hi, (n, 3))
    rpy = np.zeros((n, 3))
    grip = -np.ones((n, 1))
    nscripts = len(_SCRIPTS) if play else 2
    task = rng.integers(0, nscripts, n)
    stage = np.zeros(n, np.int64)
    dwell = np.zeros(n, np.int64)
    maxw = max(len(w) for _, w in _SCRIPTS)
    wp = np.zeros((len(_SCRIPTS), maxw, 6))
    nwp = np.zeros(len(_SCRIPTS), np.int64)
    for i, (_, w) in enumerate(_SCRIPTS):
        wp[i, :len(w)] = w
        nwp[i] = len(w)
    anchors = np.zeros((len(_SCRIPTS), n, 3))
    for i, (name, _) in enumerate(_SCRIPTS):
        anchors[i] = blk if name == 'block' else np.asarray(_ANCHORS[name])[None]
    out = np.zeros((steps, n, 7), np.float32)
    idx = np.arange(n)
    for s in range(steps):
        w = wp[task, stage]                                   # [n, 6]
        tgt_p = anchors[task, idx] + w[:, :3]
        tgt_yaw = w[:, 3]
        dp = tgt_p - pos
        dist = np.linalg.norm(dp, axis=1, keepdims=True)
        pos = pos + dp * np.minimum(1.0, 0.015 / np.maximum(dist, 1e-9))
        rpy[:, 2] += np.clip(tgt_yaw - rpy[:, 2], -0.1, 0.1)
        arrived = (dist[:, 0] < 0.015) & (np.abs(tgt_yaw - rpy[:, 2]) < 0.1)
        grip[arrived, 0] = w[arrived, 4]                      # the gripper acts once the waypoint is reached
        dwell = np.where(arrived, dwell + 1, 0)
        adv = arrived & (dwell > w[:, 5])
        stage = np.where(adv, stage + 1, stage)
        dwell = np.where(adv, 0, dwell)
        done = stage >= nwp[task]
        k = int(done.sum())
        if k:
            task[done] = rng.integers(0, nscripts, k)
            stage[done] = 0
        a = np.concatenate([pos, rpy, grip], 1)
        if jump_frac > 0:
            jump = rng.random(n) < jump_frac
            if jump.any():
                a[jump, :6] = rng.uniform(-6, 6, (int(jump.sum()), 6))
        out[s] = a
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in o.strip().split(',')]
                if len(p) >= 6:
                    self.samples.append(p)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.samples[0][1]), 'reasons': reasons,
                'samples': len(self.samples)}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    env_id, seed, wid, budget_s, preroll = args
    try:
        os.sched_setaffinity(0, {wid % os.cpu_count()})
    except Exception:
        pass
    from roboticsplayroompybullet_b200.model import load_model
    from oracle.oracle import Oracle
    m = load_model(env_id)
    o = Oracle(m, seed=seed + wid, env_id=wid)
    d0 = o.reset()
    blk = d0['achieved_goal'][None, :3] if env_id == ENV_ID else None
    ee = d0['obs_quat'][None, :3]
    acts = synth_actions(np.random.default_rng(seed + wid), 1, 4096, env_id, block_xyz=blk, ee_xyz=ee)[:, 0, :]
    # the same scripted stream and the same untimed pre-roll as the GPU arm: the timed sample starts in the steady state
    # of scripted play (sub-tasks in their contact phases), not in the approach from the rest pose
    for i in range(preroll):
        o.step(acts[i])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        o.step(acts[(preroll + n) % len(acts)])
        n += 1
    return n, time.perf_counter() - t0


def cpu_baseline(env_id, seed, budget_s=12.0, preroll=96):
    """Oracle (kind 'port') on every host core, one process per core, bounded to ~budget_s."""
    cores = os.cpu_count() or 1
    ctx = mp.get_context('fork')
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(env_id, seed, w, budget_s, preroll) for w in range(cores)])
    steps = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return {'value': steps / wall, 'unit': 'env-steps/s', 'cores': cores, 'kind': 'port',
            'sample': '%d procs x ~%.0fs of %s steps (reset + the same %d-step scripted pre-roll as the GPU arm, untimed), '
                      '%d env-steps total; PyBullet itself is not installable here or on the GPU box '
                      '(profiles/r2_pybullet_probe.log)' % (cores, budget_s, env_id, preroll, steps)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per = []
    # each "step" = one bounded sample: ~ (budget / steps) seconds of oracle stepping on all cores
    total_budget = min(150.0, max(10.0, 3.0 * (args.steps + args.warmup)))
    b = cpu_baseline(args.env, args.seed, budget_s=total_budget, preroll=args.preroll)
    v = b['value']
    line = {'impl': 'reference', 'metric': METRIC if args.env == ENV_ID else args.env + ' env-steps/s',
            'value': v, 'unit': 'env-steps/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1000.0 / v if v > 0 else None, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': '%s, CPU oracle port (fp64 C restatement of the PyBullet path) on %d host cores'
                       % (args.env, b['cores'])},
            'cpu_baseline': dict(b, value=v),
            'e2e': {'value': v, 'unit': 'env-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0, 'wall_s': time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU fallback (use --impl reference)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from roboticsplayroompybullet_b200.envs import make
    N = args.envs_per_gpu
    if args.scaling == 'strong':                      # BASELINE config 5 as written: a fixed total sharded over the ranks
        assert args.total_envs % world == 0
        N = args.total_envs // world
    K, W = args.steps, args.warmup
    env = make(args.env, num_envs=N, device=local, seed=args.seed, env_offset=rank * N)
    P = args.preroll
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    obs0 = env.reset_device()
    torch.cuda.synchronize()
    play = args.env == ENV_ID
    blk = obs0['achieved_goal'][:, :3].cpu().numpy() if play else None
    ee = obs0['obs_quat'][:, :3].cpu().numpy()
    acts_host = synth_actions(np.random.default_rng(args.seed + rank), N, P + K + W, args.env, block_xyz=blk, ee_xyz=ee,
                              jump_frac=args.jump_frac)
    acts_dev = torch.as_tensor(acts_host).to(dev)
    env.enable_kernel_timing(True)
    # untimed pre-roll: the scripted sub-tasks need ~100 steps to reach their contact phases, so that the
    # timed steps sample the steady state of scripted play rather than the approach from the rest pose
    for s in range(P + W):
        env.step_device(acts_dev[s])
    torch.cuda.synchronize()
    state0 = env.get_state()                  # the e2e leg below starts from the same states as the device-timed leg
    acts_host, acts_dev = acts_host[P:], acts_dev[P:]
    sampler = ClockSampler(local)
    sampler.start()
    # ---- device-resident timed region: exactly K steps, CUDA events on the launching stream,
    #      L2 flushed between steps (outside the per-step event pairs)
    l0 = env.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kern_ms, ik_ms, tier_ms = [], [], []
    succ = torch.zeros((), device=dev)
    rsum = torch.zeros((), device=dev)
    relabel_r = torch.zeros((), device=dev)
    n_relabel, n_reset = 0, 0
    ag_hist = env.dev['achieved_goal'].clone() if play else None
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.start()          # ncu --profile-from-start off: capture starts with the timed steps
    t_wall0 = time.perf_counter()
    for s in range(K):
        flush.fill_(float(s))
        ev[s][0].record()
        obs, r, _, info = env.step_device(acts_dev[W + s])
        if play and s % RELABEL_EVERY == RELABEL_PHASE:
            # goal relabelling (SURVEY.md \u00a78d): the stored achieved goals are rewarded against a LATER
            # achieved goal of the same env, which also becomes the env's desired goal (reset_goal_pos)
            new_goal = obs['achieved_goal'].clone()
            relabel_r += env.compute_reward_device(ag_hist, new_goal).sum()
            env.reset_goal_pos_device(new_goal)
            ag_hist.copy_(new_goal)
            n_relabel += 1
        if (P + W + s + 1) % args.episode_steps == 0:
            env.reset_device(torch.ones(N, dtype=torch.uint8, device=dev))      # episode end (masked reset path)
            n_reset += 1
        ev[s][1].record()
        a, b = env.last_kernel_ms()
        ik_ms.append(a)
        kern_ms.append(b)
        tier_ms.append(env.last_tier_ms())
        succ += info['is_success'].sum()
        rsum += r.sum()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if args.profiler_range:
        torch.cuda.profiler.stop()
    launches = env.launch_count() - l0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    dev_s = sum(step_ms) / 1000.0
    # ---- end-to-end through the public numpy API: host actions -> H2D -> step -> D2H of the whole observation block
    #      into a fresh host array, every step; replayed from the state the device-timed leg started from
    for s in range(2):                        # untimed: first-call allocations of the host path (pinned staging, result arrays)
        env.step(acts_host[W + s])
    env.set_state(state0)                     # same states and actions as the device-timed steps: the same workload, step for step
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        obs, r, done, info = env.step(acts_host[W + s])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # ---- reset cost (environments.py:173-187 on the masked step pipeline): one full-batch reset, device events around the
    #      call (it synchronises once per round), reported beside the step time and amortised over an episode
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    r0.record()
    env.reset_device(torch.ones(N, dtype=torch.uint8, device=dev))
    r1.record()
    torch.cuda.synchronize()
    reset_ms = r0.elapsed_time(r1)
    reset_rounds = env.reset_rounds()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    clocks = sampler.summary()
    # ---- max over ranks, whole-job aggregate
    t = torch.tensor([dev_s, e2e_s, reset_ms], device=dev, dtype=torch.float64)
    stats = torch.stack([succ.double(), rsum.double(), torch.tensor(float(N * K), device=dev, dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)       # the only collective: episode statistics
    dev_s, e2e_s, reset_ms = float(t[0]), float(t[1]), float(t[2])
    total_env_steps = float(stats[2])
    value = total_env_steps / dev_s
    e2e = total_env_steps / e2e_s
    if rank == 0:
        peaks = {}
        pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak_gbs = float(peaks.get('hbm_gbs', 6650.0))
        which = 'measured' if 'hbm_gbs' in peaks else 'fallback'
        kms = float(np.mean(kern_ms))
        alg_bytes = BYTES_PER_ENV_STEP[args.env] * N
        achieved = alg_bytes / (kms / 1000.0) / 1e9
        info_k = env.kernel_info()
        line = {'metric': METRIC if args.env == ENV_ID else args.env + ' env-steps/s', 'value': value,
                'unit': 'env-steps/s', 'n_gpus': world, 'steps': K, 'warmup': W,
                'ms_per_step': 1000.0 * dev_s / K, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': '%s, %d envs per GPU (%d total), sharded by env index; scripted teleop-shaped synthetic '
                                       'sub-task trajectories (reach / grasp+lift+carry / drawer pull+push / door slide / button / dial), goal '
                                       'relabelling every %d steps, %d untimed pre-roll steps; L2 flushed (256 MB fill) '
                                       'between timed steps' % (args.env, N, N * world, RELABEL_EVERY, P),
                           'envs_per_gpu': N, 'substeps_per_step': 12, 'solver_iterations': 50, 'seed': args.seed,
                           'jump_frac': args.jump_frac, 'preroll_steps': P, 'relabel_events': n_relabel,
                           'episode_resets': n_reset},
                'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak_gbs, 'unit': 'GB/s',
                             'frac': achieved / peak_gbs, 'traffic': MEASURED_TRAFFIC_BYTES.get((args.env, N)),
                             'traffic_note': 'bytes per launch set of one env step (dram__bytes_read.sum + dram__bytes_write.sum, profiles/r2_ncu.md)',
                             'kernel': 'step pipeline per env step: 13 x prb_setup_kernel + 12 x (prb_pgs_joint_kernel, prb_pgs_free_kernel, 5 size classes of prb_pgs_arm_kernel)',
                             'setup_kernels_ms': float(np.mean([t[0] for t in tier_ms])),
                             'pgs_kernels_ms': float(np.mean([t[1] for t in tier_ms])),
                             'kernel_ms': kms, 'ik_kernel_ms': float(np.mean(ik_ms)), 'peak_source': which,
                             'algorithmic_bytes_per_env_step': BYTES_PER_ENV_STEP[args.env],
                             'note': 'latency/FP32-issue bound by construction (SURVEY.md §8d); see profiles/',
                             'smem_bytes_per_block': info_k['smem_bytes_per_block'],
                             'regs_per_thread': info_k['regs_per_thread']},
                'e2e': {'value': e2e, 'unit': 'env-steps/s', 'h2d_bytes_per_step': env.h2d_bytes_per_step,
                        'd2h_bytes_per_step': env.d2h_bytes_per_step},
                'gpu_launches': int(launches), 'clocks': clocks,
                'episode_stats': {'success_rate': float(stats[0]) / total_env_steps,
                                  'mean_reward': float(stats[1]) / total_env_steps},
                'capacity_overflow_env_steps': env.overflow_count(),
                'reset': {'full_batch_ms': reset_ms, 'rounds': reset_rounds, 'episode_steps': args.episode_steps,
                          'value_amortised': total_env_steps / K * args.episode_steps /
                          (args.episode_steps * dev_s / K + reset_ms / 1000.0),
                          'note': 'env-steps/s of a whole episode: episode_steps steps + one full-batch reset'},
                'wall_s_timed_region': t_wall}
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline(args.env, args.seed, budget_s=args.cpu_seconds, preroll=args.preroll)
        elif not args.no_cpu_baseline:
            line['cpu_baseline'] = None
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--env', default=ENV_ID)
    ap.add_argument('--envs-per-gpu', type=int, default=65536)
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak: --envs-per-gpu envs on every GPU; strong: --total-envs sharded over the GPUs (BASELINE config 5 as written)')
    ap.add_argument('--total-envs', type=int, default=65536)
    ap.add_argument('--seed', type=int, default=1234)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--preroll', type=int, default=96, help='untimed scripted steps before the warm-up')
    ap.add_argument('--jump-frac', type=float, default=0.0, help='fraction of steps that target the full +-6 clip box (Random-stream tail)')
    ap.add_argument('--profiler-range', action='store_true', help='cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)')
    ap.add_argument('--episode-steps', type=int, default=512, help='masked reset of every env after this many steps')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
