"""Headline benchmark: acoustic frames/s through MFCC compression + energy heat map (+ mean mask and IoU success
counts) on synthetic ACIVW-shaped data - BASELINE.json's metric on configs[1].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames B] [--impl reference] [--single-process]

One "step" is one pass of the hot path over one batch of B resident frames (default 8192 frames = 29 GB of float32
spectra per GPU, far larger than the 126 MB L2, so every step streams from HBM): stages 1 + 2 in one persistent kernel
(aig_mfcc_energy) and the stage-3 IoU sweep of the first half of the batch ("real" stream) against the second half
("reconstructed" stream), 11 thresholds, device-resident counters.  At N > 1 every rank runs the same per-GPU workload
on its own GPU (weak scaling; frames shard trivially) and the only exchange is one NCCL all-reduce of the int64[K+1]
count vector inside the timed region.

Prints ONE JSON line (rank 0):
  value       whole-job frames/s with inputs resident in HBM (CUDA events, max over ranks)
  sustained   the same step repeated for >= 3 s (the board sits at its 1 kW cap after ~0.5 s): frames/s, roofline
              fraction, clocks and throttle reasons of that loop; this is also BASELINE configs[4] (full chain + NCCL
              all-reduce over millions of frames), with the reduced counts checked against a rank-0 recomputation
  roofline    the persistent kernel's algorithmic GB/s (event-timed per launch inside the timed region) against the
              measured HBM peak
  e2e         the same metric through the public host API (pinned NumPy in / out, H2D / D2H inside the timed region),
              next to a bare cudaMemcpyAsync probe of the same bytes run by all ranks at once (bare_link_gbs,
              frac_of_bare_link): what the host <-> device fabric gives at this N
  configs     short runs of the other BASELINE configs (C1 16-frame host call and the reference-shaped add_batch(16);
              C3 5 000 frames heat map 224 x 224 + consensus IoU over 101 thresholds; C4 clip stream sharded by clip;
              the TFRecord loader), each with a parity spot check against the NumPy oracle
  cpu_baseline the oracle port of the reference's NumPy path on this box's host cores (rank 0, N = 1)
`--impl reference` runs only that CPU arm.  `--single-process` drives all N GPUs from one process through
AcousticPathGroup (the reference's callers are single processes) instead of one rank per GPU.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'acoustic frames/sec (36x48x512->MFCC+energy)'
UNIT = 'frames/s'
FRAME_PIXELS, FFT_LEN, MFCC_NUM = 1728, 512, 12
IN_BYTES = FRAME_PIXELS * FFT_LEN * 4                    # 3 538 944 B of spectra per frame
MFCC_BYTES = FRAME_PIXELS * MFCC_NUM * 4                 # 82 944 B MFCC image per frame
ALGO_BYTES_MFCC_KERNEL = IN_BYTES + MFCC_BYTES + FRAME_PIXELS * 4   # 3 628 800 B per frame through the fused MFCC + energy kernel
ALGO_BYTES_PATH = IN_BYTES + MFCC_BYTES + FRAME_PIXELS * 4   # 3 628 800 B per frame, MFCC + energy
OUT_BYTES = MFCC_BYTES + FRAME_PIXELS * 8 + FRAME_PIXELS     # results of the host call: MFCC f32 + energy f64 + mask u8
THRESHOLDS = (0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)
CPU_SAMPLE_FRAMES = 16                                   # BASELINE.json configs[0]
RING_BLOCK = 64                                          # host-generated frames per stream, cycled to fill the ring
LAYOUT_SEED, STREAM_SHIFT = 1000, 3.0                    # synth.structured_power_frames: shared source layout, displacement of stream B
WRITE_ROOF_GBS = 6300.0                                  # st.global / bulk-copy write-only roof measured in round 1 (profiles/r01_write_peak_probe.txt)


def workload_config(frames, n_gpus, single_process=False):
    return {
        'workload': 'configs[1]: MFCC compression + energy heatmap, synthetic ACIVW frames 36x48x512 f32, '
                    '%d resident frames per step per GPU (%.1f GB > L2), + mean mask + IoU sweep (11 thresholds)'
                    % (frames, frames * IN_BYTES / 1e9),
        'frames_per_step_per_gpu': frames,
        'flip180': True,
        'normalize_first': True,
        'l2': 'inputs larger than L2 (no flush needed)',
        'inputs': 'two seeded NumPy streams of %d frames each (synth.structured_power_frames: squared-normal spectra with one to '
                  'three Gaussian sources per frame; stream B shows the same sources displaced by N(0, %.0f px)), generated on '
                  'the host and cycled to fill the ring: stream A = first half, stream B = second half rotated by the rank, '
                  'frame i of A scored against frame i of B' % (RING_BLOCK, STREAM_SHIFT),
        'parallelism': ('one process, %d device threads (AcousticPathGroup), counts summed by one grouped NCCL all-reduce' % n_gpus)
        if single_process else 'frames sharded by rank, dp%d, one NCCL all-reduce of int64[12] per run' % n_gpus,
        'known_deviation': 'an all-floor frame (every mel band under 0.001) has a mathematically zero MFCC image whose min-max '
                           'stretch is rounding noise in the reference and here alike: its mask is not reproduced (DESIGN.md 2); '
                           'the synthetic streams contain no such frame',
    }


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's NumPy path (test infrastructure used as the baseline)
# --------------------------------------------------------------------------------------------------
def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        counts = [p.get('num_threads', 1) for p in threadpool_info() if p.get('user_api') == 'blas']
        if counts:
            return int(max(counts))
    except Exception:
        pass
    return int(os.cpu_count() or 1)


def base_streams(n):
    """The two host-generated streams every rank, the CPU arm and the parity checks share."""
    from acoustic_image_generation_b200 import synth
    return (synth.structured_power_frames(n, 0, LAYOUT_SEED, 0.0),
            synth.structured_power_frames(n, 1, LAYOUT_SEED, STREAM_SHIFT))


def cpu_chain(oracle, power):
    """The reference path on the host for one sample (first half = stream A, second half = stream B):
    get_feats -> flip -> min-max -> find_logen -> mask -> IoU -> counts."""
    mfcc = oracle.mfcc_image(power, flip=True)
    energy, mask = oracle.energy_stage(mfcc, normalize_first=True)
    half = len(mask) // 2
    scores = [oracle.iou_pair(a, b)[2] for a, b in zip(mask[:half], mask[half:])]
    cpu_chain.last = (mfcc, mask)            # for the parity check of the GPU arm against this same sample
    return oracle.success_counts(scores, THRESHOLDS)


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core NumPy's BLAS can."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=int(os.cpu_count() or 1), user_api='blas')
    except Exception:
        pass


def time_cpu(steps, warmup, budget_s=None, streams=None):
    from oracle import acoustic_oracle as oracle
    use_all_host_threads()
    half = CPU_SAMPLE_FRAMES // 2
    a, b = streams if streams is not None else base_streams(half)
    power = np.concatenate([a[:half], b[:half]], 0)    # frames 0..7 of each stream: 16 frames = BASELINE configs[0]
    for _ in range(warmup):
        cpu_chain(oracle, power)
    times = []
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_chain(oracle, power)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_all > budget_s and len(times) >= 3:
            break
    total = sum(times)
    single = None
    try:                                     # the same chain on one BLAS thread, a 3 s sample
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1, user_api='blas'):
            cpu_chain(oracle, power)
            t1, n1 = time.perf_counter(), 0
            while time.perf_counter() - t1 < 3.0:
                cpu_chain(oracle, power)
                n1 += 1
            single = CPU_SAMPLE_FRAMES * n1 / (time.perf_counter() - t1)
    except Exception:
        pass
    return {
        'value': CPU_SAMPLE_FRAMES * len(times) / total,
        'single_thread_value': single,
        'unit': UNIT,
        'cores': blas_threads(),
        'kind': 'port',
        'sample': '%d x %d-frame batches (BASELINE configs[0]; frames 0-7 of both streams) through the NumPy oracle port of '
                  'get_feats/find_logen/IoU, %.1f s, median %.1f ms per batch, host cpu_count=%s'
                  % (len(times), CPU_SAMPLE_FRAMES, total, 1e3 * statistics.median(times), os.cpu_count()),
    }, len(times), total


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base, steps_done, total = time_cpu(args.steps, max(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps_done, 'warmup': max(args.warmup, 1), 'ms_per_step': 1e3 * total / steps_done,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict(workload_config(args.frames, args.gpus),
                       note='CPU arm: each step is a bounded %d-frame sample of the same workload, on ONE host whatever N is '
                            '(the CPU arm does not scale with the GPU count)' % CPU_SAMPLE_FRAMES),
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# clocks sampler (NVML) for the timed regions
# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: 'gpu_idle', 0x2: 'applications_clocks_setting', 0x4: 'sw_power_cap', 0x8: 'hw_slowdown',
               0x10: 'sync_boost', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown', 0x100: 'display_clock_setting'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.power, self.mask, self.max_mhz, self.stop_flag, self.ok = index, [], [], 0, None, False, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:                                    # CUDA ordinal -> NVML device through the UUID
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith('GPU-') else 'GPU-' + uuid)
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.handle, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                break
            time.sleep(0.01)

    def result(self):
        self.stop_flag = True
        if self.ok:
            self.join(timeout=1.0)
        reasons = [name for bit, name in self.REASONS.items() if self.mask & bit and name != 'gpu_idle']
        return {'sm_mhz': statistics.median(self.samples) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'reasons': reasons, 'samples': len(self.samples),
                'power_w_max': max(self.power) if self.power else None}


def bind_host_threads(torch, local, world):
    """Keep this rank's host threads (and, by first touch, its pinned staging buffers) near its GPU: the GPU's NUMA node when
    the host exposes one, else an equal slice of the CPUs this process may run on, so that eight ranks' copy threads do not
    pile onto the same cores.  Returns a description for the bench line."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        prop = torch.cuda.get_device_properties(local)
        bdf = '%04x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        node = -1
        try:
            node = int(open('/sys/bus/pci/devices/%s/numa_node' % bdf).read().strip())
        except Exception:
            pass
        if node >= 0:
            cpus = set()
            for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
            mine = sorted(cpus & set(allowed))
            if mine:
                os.sched_setaffinity(0, mine)
                return {'numa_node': node, 'cpus': len(mine), 'how': 'numa node of the GPU'}
        if world > 1 and len(allowed) >= 2 * world:
            per = len(allowed) // world
            mine = allowed[local * per:(local + 1) * per]
            os.sched_setaffinity(0, mine)
            return {'numa_node': None, 'cpus': len(mine), 'how': 'equal slice of %d visible CPUs (the host exposes no NUMA node for the GPU)' % len(allowed)}
        return {'numa_node': None, 'cpus': len(allowed), 'how': 'unbound'}
    except Exception as exc:                 # never let placement break the bench
        return {'numa_node': None, 'cpus': None, 'how': 'failed: %r' % (exc,)}


# --------------------------------------------------------------------------------------------------
# helpers of the GPU arm
# --------------------------------------------------------------------------------------------------
def expected_counts(torch, mask_a_block, mask_b_block, half, world, steps):
    """What the all-reduced count vector must be, recomputed on this rank from the masks of the two base blocks alone:
    rank r scores frame i of stream A (block frame i % 64) against block frame (i + r) % 64 of stream B, `steps` times."""
    n = mask_a_block.shape[0]
    a = (mask_a_block.reshape(n, -1) != 0)
    b = (mask_b_block.reshape(n, -1) != 0)
    idx = torch.arange(half, device=a.device)
    thr = torch.tensor(THRESHOLDS, dtype=torch.float64, device=a.device)
    pos = torch.zeros(len(THRESHOLDS), dtype=torch.int64, device=a.device)
    for r in range(world):
        ma, mb = a[idx % n], b[(idx + r) % n]
        inter = (ma & mb).sum(1).to(torch.float64)
        union = (ma | mb).sum(1).to(torch.float64)
        iou = inter / union                                           # 0 / 0 = NaN never counts
        pos += (iou[:, None] > thr[None, :]).sum(0)
    return np.concatenate([pos.cpu().numpy() * steps, [half * world * steps]]).astype(np.int64)


def bare_link_probe(torch, dev, h_in, d_in, h_outs, d_outs, steps, barrier):
    """The same bytes as one end-to-end step, as bare pinned cudaMemcpyAsync calls and nothing else: the spectra host ->
    device in 64 MiB pieces on one stream while the result-sized buffers go device -> host on a second one.  All ranks
    run it at the same time (barrier first), so it sees the contention for the host's memory and PCIe fabric that the
    library call sees at this N.  Returns (seconds for `steps` passes, bytes up per pass, bytes down per pass)."""
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    src, dst = h_in.view(-1), d_in.view(-1)
    piece = (64 << 20) // src.element_size()
    for _ in range(1):
        dst[:piece].copy_(src[:piece], non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        with torch.cuda.stream(up):
            for lo in range(0, src.numel(), piece):
                dst[lo:lo + piece].copy_(src[lo:lo + piece], non_blocking=True)
        with torch.cuda.stream(down):
            for h, d in zip(h_outs, d_outs):
                h.copy_(d, non_blocking=True)
        up.synchronize()
        down.synchronize()
    seconds = time.perf_counter() - t0
    return seconds, src.numel() * src.element_size(), sum(h.numel() * h.element_size() for h in h_outs)


def all_max(torch, dist, world, dev, value):
    if world == 1:
        return float(value)
    t = torch.tensor([value], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------------
# sub-records of the other BASELINE configs
# --------------------------------------------------------------------------------------------------
def config_c1(torch, aig, path, stream_a, stream_b, dev):
    """configs[0]: one 16-frame batch, host NumPy in / out (what tf.py_func hands the reference's functions), and the
    reference-shaped evaluation step add_batch(B = 16) (iouenergythreshold.py:213-229)."""
    from acoustic_image_generation_b200 import evaluate
    from oracle import acoustic_oracle as oracle
    power = np.concatenate([stream_a[:8], stream_b[:8]], 0)                 # pageable, 56.6 MB

    def lat(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    t_call = lat(lambda: path.mfcc_energy(power, flip=True, normalize_first=True), 10)
    # the same call from a source that is NOT in the host's caches (a pipeline producing a fresh batch per call): sixteen
    # copies of the batch in rotation, 906 MB
    cold = [power.copy() for _ in range(16)]
    turn = [0]

    def cold_call():
        turn[0] = (turn[0] + 1) % len(cold)
        path.mfcc_energy(cold[turn[0]], flip=True, normalize_first=True)

    t_call_cold = lat(cold_call, 32)
    del cold
    mfcc, energy, mask = path.mfcc_energy(power, flip=True, normalize_first=True)
    want_mfcc = oracle.mfcc_image(power, flip=True)
    want_energy, want_mask = oracle.energy_stage(want_mfcc, normalize_first=True)
    # the reference's evaluation step on the normalised images: real (stream A) vs reconstructed (stream B)
    real, recon = oracle.normalize_acoustic_images(want_mfcc[:8]), oracle.normalize_acoustic_images(want_mfcc[8:])
    real16, recon16 = np.concatenate([real, real], 0), np.concatenate([recon, recon], 0)       # B = 16
    ev = evaluate.AcivwEvaluation(path)
    t_batch_host = lat(lambda: ev.add_batch(real16, recon16), 50)
    d_real, d_recon = torch.from_numpy(real16).to(dev), torch.from_numpy(recon16).to(dev)
    t_batch_dev = lat(lambda: ev.add_batch(d_real, d_recon), 200)
    ev2 = evaluate.AcivwEvaluation(path)
    inter, union = ev2.add_batch(real16, recon16)
    res = ev2.finish()
    rows = [oracle.iou_pair(oracle.mean_mask(oracle.find_logen(a.copy())), oracle.mean_mask(oracle.find_logen(b.copy())))
            for a, b in zip(real16, recon16)]
    want_pos, want_num = oracle.success_counts([r[2] for r in rows], THRESHOLDS)
    frame = real16[0].copy()
    t_find_logen = lat(lambda: aig.find_logen(frame.copy()), 200)
    t0 = time.perf_counter()
    for _ in range(20):
        oracle.find_logen(frame.copy())
    t_find_logen_numpy = (time.perf_counter() - t0) / 20
    return {
        'what': 'configs[0]: 16-frame batch through the host API (pageable NumPy in, NumPy out) and the reference-shaped '
                'evaluation step add_batch(B=16) as ONE kernel launch (aig_acivw_batch, cluster-per-frame form)',
        'mfcc_energy_16_frames_ms': 1e3 * t_call, 'mfcc_energy_16_frames_per_s': 16 / t_call,
        'mfcc_energy_16_frames_note': 'one array sent repeatedly (it stays in the host caches); cold_source = sixteen arrays in rotation',
        'mfcc_energy_16_frames_cold_source_ms': 1e3 * t_call_cold, 'mfcc_energy_16_frames_cold_source_per_s': 16 / t_call_cold,
        'h2d_bytes': int(power.nbytes), 'link_time_ms_at_55GBs': 1e3 * power.nbytes / 55e9,
        'add_batch16_host_numpy_us': 1e6 * t_batch_host, 'add_batch16_device_tensors_us': 1e6 * t_batch_dev,
        'find_logen_dropin_us': 1e6 * t_find_logen, 'find_logen_numpy_oracle_us': 1e6 * t_find_logen_numpy,
        'find_logen_speedup_vs_numpy': t_find_logen_numpy / t_find_logen,
        'parity': {'mfcc_max_abs_err': float(np.abs(mfcc - want_mfcc).max()), 'mfcc_tolerance': 1e-4,
                   'mask_pixels_differing': int((mask != want_mask).sum()),
                   'energy_max_rel_err': float(np.abs(energy - want_energy).max() / np.abs(want_energy).max()),
                   'add_batch_IU_exact': bool(np.array_equal(inter, [r[0] for r in rows]) and np.array_equal(union, [r[1] for r in rows])),
                   'add_batch_counts_exact': bool(np.array_equal(res['pos'], want_pos) and res['num'] == want_num)},
    }


def config_c3(torch, aig, path, dev):
    """configs[2]: 5 000 frames, energy heat map up-sampled to 224 x 224 + consensus-IoU sweep over 101 thresholds on
    FlickrSoundNet-shaped boxes."""
    from acoustic_image_generation_b200 import synth
    from oracle import acoustic_oracle as oracle
    n, shape, block = 5000, (224, 224), 50
    base = synth.smooth_images(block, 20)
    images = torch.from_numpy(base).to(dev).repeat(n // block, 1, 1, 1)
    boxes_h = synth.flickr_boxes(n, 21, height=shape[0], width=shape[1])
    boxes = [torch.from_numpy(np.ascontiguousarray(b)).to(dev) for b in boxes_h]
    thr = torch.linspace(0, 1, 101, dtype=torch.float64, device=dev)
    thr_h = np.linspace(0, 1, 101)
    heat = torch.empty((n,) + shape, dtype=torch.float32, device=dev)
    counts = torch.zeros(102, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        _, mask, _ = path.energy_heatmap(images, False, *shape, want_energy=False, want_mask=True, out=heat)
        return path.ciou_sweep(mask, *boxes, thr, out_hw=shape, pos=counts[:-1], num=counts[-1:])

    for _ in range(3):
        step()
    counts.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    path.set_option('profile', 1)
    e0.record(stream)
    for _ in range(reps):
        i2, u2, _, _ = step()
    e1.record(stream)
    torch.cuda.synchronize()
    prof = path.profile_read()
    path.set_option('profile', 0)
    ms = e0.elapsed_time(e1) / reps
    host = counts.cpu().numpy() // reps
    rates = aig.success_rates(host[:-1], int(host[-1]))
    # parity spot check: 8 frames through the oracle (find_logen -> mean mask -> exact-integer up-sampling -> consensus IoU)
    k = 8
    want_heat, ok_iu = [], True
    for f in range(k):
        energy = oracle.find_logen(base[f].copy())
        want_heat.append(oracle.heatmap(energy, *shape))
        gt = oracle.boxes_to_consensus(*[b[f] for b in boxes_h], shape[0], shape[1])
        w_i2, w_u2, _ = oracle.consensus_iou(gt, oracle.resize_mask(oracle.mean_mask(energy), *shape))
        ok_iu &= int(i2[f]) == w_i2 and int(u2[f]) == w_u2
    heat_err = float(np.abs(heat[:k].cpu().numpy() - np.stack(want_heat, 0)).max())
    heat_ms = prof['energy'][0] / max(prof['energy'][1], 1)
    return {
        'what': 'configs[2]: 5000 frames, find_logen -> heat map 224x224 (one launch, bulk-copy stores) + mean mask -> '
                'consensus-IoU sweep, 101 thresholds, device-resident inputs',
        'frames': n, 'ms_per_pass': ms, 'frames_per_s': n / ms * 1e3,
        'energy_heatmap_kernel_ms': heat_ms, 'energy_heatmap_kernel_frames_per_s': n / heat_ms * 1e3,
        'heat_write_gbs': n * shape[0] * shape[1] * 4 / heat_ms / 1e6,
        'heat_write_frac_of_write_roof_6300': n * shape[0] * shape[1] * 4 / heat_ms / 1e6 / WRITE_ROOF_GBS,
        'bound': 'float64 pipe / issue slots of find_logen (the heat-map stores alone run at 0.8 of the write roof)',
        'auc_101_thresholds': aig.auc(thr_h, rates), 'pos_at_0.5': int(host[50]), 'num': int(host[-1]),
        'parity': {'frames_checked': k, 'heat_max_abs_err': heat_err, 'heat_tolerance': 1e-4, 'I2_U2_exact': bool(ok_iu)},
    }


def config_c4(torch, dist, aig, path, dev, power, half, rank, world, mask_blocks):
    """configs[3]: a stream of 10 s clips at 30 fps (300 frames), MFCC + energy + per-clip and global AUC, sharded by clip."""
    from acoustic_image_generation_b200 import sharding
    frames_per_clip = 300
    clips_per_rank = min(8, half // frames_per_clip)     # the clips are cut from the two halves of the resident ring
    n_clips = clips_per_rank * world
    c0, c1, _, _ = sharding.shard_clips(n_clips, frames_per_clip, rank, world)
    n = (c1 - c0) * frames_per_clip                      # frame pairs of this rank: ring frames [0, n) vs [half, half + n)
    assert n <= half, 'the resident ring is too small for the clip stream'
    thr = torch.tensor(THRESHOLDS, dtype=torch.float64, device=dev)
    mfcc = torch.empty((2 * n, 36, 48, MFCC_NUM), dtype=torch.float32, device=dev)
    energy = torch.empty((2 * n, 36, 48), dtype=torch.float64, device=dev)
    mask = torch.empty((2 * n, 36, 48), dtype=torch.uint8, device=dev)
    clip_pos = torch.zeros((c1 - c0, len(THRESHOLDS)), dtype=torch.int64, device=dev)
    total = torch.zeros(len(THRESHOLDS) + 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        path.mfcc_energy(power[:n], flip=True, normalize_first=True, out=(mfcc[:n], energy[:n], mask[:n]))
        path.mfcc_energy(power[half:half + n], flip=True, normalize_first=True, out=(mfcc[n:], energy[n:], mask[n:]))
        clip_pos.zero_()
        path.iou_sweep_clips(mask[:n], mask[n:], frames_per_clip, thr, pos=clip_pos)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    for _ in range(reps):
        step()
    total[:-1] = clip_pos.sum(0)
    total[-1] = n
    if world > 1:
        path.allreduce_counts(total)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = all_max(torch, dist, world, dev, e0.elapsed_time(e1) / reps)
    clip_host = clip_pos.cpu().numpy()
    clip_auc = [aig.auc(THRESHOLDS, aig.success_rates(p, frames_per_clip)) for p in clip_host]
    tot = total.cpu().numpy()
    want = expected_counts(torch, mask_blocks[0], mask_blocks[1], n, world, 1)
    return {
        'what': 'configs[3]: %d clips x %d frames (10 s at 30 fps) of both streams, MFCC + energy + mean mask, per-clip success '
                'curves in one launch (aig_iou_sweep_clips), sharded by clip over %d GPU(s); global counts all-reduced'
                % (n_clips, frames_per_clip, world),
        'clips': n_clips, 'frames_per_clip': frames_per_clip, 'frame_pairs': n * world,
        'ms_per_pass': ms, 'frames_per_s': 2 * n * world / ms * 1e3,
        'global_auc': aig.auc(THRESHOLDS, aig.success_rates(tot[:-1], int(tot[-1]))),
        'rank0_clip_auc_min': float(min(clip_auc)), 'rank0_clip_auc_max': float(max(clip_auc)),
        'global_counts': [int(v) for v in tot],
        'counts_equal_rank0_recomputation': bool(np.array_equal(tot, want)),
    }


def config_loader(torch, aig, path, dev):
    """The on-disk side ("next" row N3 / N1-N3 composed): GZIP TFRecord files -> device tensors through AcousticBatches,
    against the files' bare zlib inflate time (the ceiling a reader of this format has on one thread per file)."""
    import gzip
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    from acoustic_image_generation_b200 import loader, synth, tfrecord
    files, records, frames = 8, 6, 12
    with tempfile.TemporaryDirectory() as tmp:
        paths = []
        for f in range(files):
            blobs = []
            for r in range(records):
                images = synth.smooth_images(frames, 100 * f + r) * np.float32(7.0) - np.float32(3.0)
                audio = synth.audio_rows(frames, 100 * f + r, np.int32)
                blobs.append(tfrecord.encode_sequence_example(
                    {'classes': r % 10, 'location': f % 3, 'audio_image/height': 36, 'audio_image/width': 48, 'audio_image/depth': 12,
                     'audio_data/mics': 1, 'audio_data/samples': 1024},
                    {'audio/image': [x.tobytes() for x in images], 'audio/data': [x.tobytes() for x in audio]}))
            paths.append(tfrecord.write_sequence_examples(os.path.join(tmp, 'Data_%03d.tfrecord' % f), blobs))
        n_frames = files * records * frames
        raw = [open(p, 'rb').read() for p in paths]
        inflated = sum(len(gzip.decompress(b)) for b in raw)

        def inflate_all(workers):
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=workers) as pool:
                list(pool.map(gzip.decompress, raw))
            return time.perf_counter() - t0

        inflate_all(4)
        t_inflate4 = min(inflate_all(4) for _ in range(3))
        t_inflate1 = min(inflate_all(1) for _ in range(2))

        def load_all():
            n = 0
            for batch in loader.AcousticBatches(paths, batch_frames=48, path=path, workers=4, low_pass=True):
                n += batch['acoustic'].shape[0]
            torch.cuda.synchronize()
            return n

        load_all()
        t0 = time.perf_counter()
        got = load_all()
        t_load = time.perf_counter() - t0
    assert got == n_frames
    return {
        'what': 'GZIP TFRecord files (%d files x %d records x %d frames of [36,48,12] f32 + 1024 int32 audio samples) -> '
                'device tensors through loader.AcousticBatches (4 reader threads, upload through the pinned staging ring, '
                'normalise / audio MFCC / low-pass / filtered-audio MFCC / tile kernels)' % (files, records, frames),
        'frames': n_frames, 'compressed_mb': sum(len(b) for b in raw) / 1e6, 'inflated_mb': inflated / 1e6,
        'loader_frames_per_s': n_frames / t_load,
        'zlib_inflate_ceiling_frames_per_s_4_threads': n_frames / t_inflate4,
        'zlib_inflate_ceiling_frames_per_s_1_thread': n_frames / t_inflate1,
        'loader_frac_of_inflate_ceiling_4_threads': t_inflate4 / t_load,
    }


# --------------------------------------------------------------------------------------------------
# GPU arm, one process per GPU
# --------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import acoustic_image_generation_b200 as aig

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    # stdout carries exactly ONE JSON line: anything native libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    placement = bind_host_threads(torch, local, world)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    path = aig.AcousticPath(local, stream=stream.cuda_stream)
    if world > 1:
        path.init_comm()                            # libaig's own NCCL communicator (aig_comm_init)
    frames = args.frames
    half = frames // 2
    # Synthetic spectra generated ON THE HOST by the seeded NumPy generator the tests and the CPU arm use, then cycled on
    # the device: stream A fills the first half of the ring, stream B (same sources, displaced) the second half, rotated
    # by the rank so that every rank scores different pairs; RING_BLOCK frames each (226 MB > the 126 MB L2).  Every
    # value is reproducible on the CPU and every rank's counts can be recomputed from the two blocks' masks.
    stream_a, stream_b = base_streams(RING_BLOCK)
    block_a, block_b = torch.from_numpy(stream_a).to(dev), torch.from_numpy(stream_b).to(dev)
    power = torch.empty((frames, 36, 48, FFT_LEN), device=dev, dtype=torch.float32)
    for f0 in range(0, half, RING_BLOCK):
        n = min(RING_BLOCK, half - f0)
        power[f0:f0 + n].copy_(block_a[:n])
    idx_b = (torch.arange(frames - half, device=dev) + rank) % RING_BLOCK
    for f0 in range(0, frames - half, 256):
        sel = idx_b[f0:f0 + 256]
        power[half + f0:half + f0 + len(sel)].copy_(block_b[sel])
    mfcc = torch.empty((frames, 36, 48, MFCC_NUM), device=dev, dtype=torch.float32)
    energy = torch.empty((frames, 36, 48), device=dev, dtype=torch.float64)
    mask = torch.empty((frames, 36, 48), device=dev, dtype=torch.uint8)
    thr = torch.tensor(THRESHOLDS, device=dev, dtype=torch.float64)
    counts = torch.zeros(len(THRESHOLDS) + 1, device=dev, dtype=torch.int64)   # pos[0..K-1], num

    def step():
        path.mfcc_energy(power, flip=True, normalize_first=True, out=(mfcc, energy, mask))
        path.iou_sweep(mask[:half], mask[half:2 * half], thr, pos=counts[:-1], num=counts[-1:])

    def timed_loop(steps):
        """`steps` steps + the all-reduce between two events; returns (ms, clocks, launches, profile, host counts)."""
        counts.zero_()
        path.set_option('profile', 1)
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = path.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        if world > 1:
            path.allreduce_counts(counts)           # the path's only exchange: int64[K+1] success counts, NCCL via libaig,
                                                    # enqueued on the same stream as the sweeps that filled the vector
        e1.record(stream)
        barrier()
        ms = all_max(torch, dist, world, dev, e0.elapsed_time(e1))
        clocks = sampler.result()
        launches = path.launch_count - launches0
        prof = path.profile_read()
        path.set_option('profile', 0)
        return ms, clocks, launches, prof, counts.cpu().numpy().copy()

    for _ in range(args.warmup):
        step()
    if world > 1:
        path.allreduce_counts(counts)               # warm-up: NCCL sets up its channels on the first collective
    barrier()
    elapsed_ms, clocks, launches, prof, host_counts = timed_loop(args.steps)
    total_frames = world * frames * args.steps
    value = total_frames / (elapsed_ms / 1e3)
    # the masks of the two base blocks, from this rank's own timed pass: stream A block = ring frames 0..63; stream B block
    # frame j sits at ring position half + ((j - rank) mod 64)
    inv = (torch.arange(RING_BLOCK, device=dev) - rank) % RING_BLOCK
    mask_blocks = (mask[:RING_BLOCK].clone(), mask[half:half + RING_BLOCK][inv].clone())
    counts_ok = bool(np.array_equal(host_counts, expected_counts(torch, mask_blocks[0], mask_blocks[1], half, world, args.steps)))
    rates = aig.success_rates(host_counts[:-1], max(int(host_counts[-1]), 1))
    auc = aig.auc(THRESHOLDS, rates)

    # ---- sustained: the same step for >= args.sustain_seconds (BASELINE configs[4]: millions of frames + the all-reduce) ----
    sustained = None
    if args.sustain_seconds > 0:
        n_sust = max(args.steps, int(np.ceil(args.sustain_seconds * 1e3 / (elapsed_ms / args.steps))))
        s_ms, s_clocks, s_launches, s_prof, s_counts = timed_loop(n_sust)
        s_mfcc_ms, s_mfcc_launches = s_prof['mfcc']
        s_achieved = ALGO_BYTES_MFCC_KERNEL * frames / (s_mfcc_ms / max(s_mfcc_launches, 1) / 1e3) / 1e9
        sustained = {
            'what': 'the same step back to back for >= %.0f s (burst != sustained on this board: it reaches its power cap on memory '
                    'traffic alone); also BASELINE configs[4]: MFCC + energy + IoU counts over %.1f M frames with the NCCL '
                    'all-reduce of the count vector at the end' % (args.sustain_seconds, world * frames * n_sust / 1e6),
            'steps': n_sust, 'seconds': s_ms / 1e3, 'frames': world * frames * n_sust,
            'value': world * frames * n_sust / (s_ms / 1e3), 'unit': UNIT, 'ms_per_step': s_ms / n_sust,
            'kernel_achieved_gbs': s_achieved, 'clocks': s_clocks, 'gpu_launches': int(s_launches),
            'counts': [int(v) for v in s_counts],
            'counts_equal_rank0_recomputation': bool(np.array_equal(
                s_counts, expected_counts(torch, mask_blocks[0], mask_blocks[1], half, world, n_sust))),
        }

    # ---- end to end through the host API: pinned NumPy in / out, copies inside the timed region ----
    e2e_frames = min(args.e2e_frames, frames)
    pin = lambda shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True)
    h_power = pin((e2e_frames, 36, 48, FFT_LEN), torch.float32)
    h_power.copy_(power[:e2e_frames])
    h_out = (pin((e2e_frames, 36, 48, MFCC_NUM), torch.float32), pin((e2e_frames, 36, 48), torch.float64),
             pin((e2e_frames, 36, 48), torch.uint8))
    np_power, np_out = h_power.numpy(), tuple(t.numpy() for t in h_out)
    for _ in range(2):
        path.mfcc_energy(np_power, flip=True, normalize_first=True, out=np_out)
    e2e_times, bare_times = [], []
    for _ in range(args.e2e_rounds):                 # library call and bare copies alternate, all ranks in step; every round counts
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            path.mfcc_energy(np_power, flip=True, normalize_first=True, out=np_out)   # returns after the D2H copies
        torch.cuda.synchronize()
        e2e_s = all_max(torch, dist, world, dev, time.perf_counter() - t0)
        e2e_times.append(e2e_s)
        bare_s, up_bytes, down_bytes = bare_link_probe(torch, dev, h_power, power[:e2e_frames], h_out,
                                                       (mfcc[:e2e_frames], energy[:e2e_frames], mask[:e2e_frames]),
                                                       args.e2e_steps, barrier)
        bare_s = all_max(torch, dist, world, dev, bare_s)
        bare_times.append(bare_s)
    # the reported figure is the aggregate over ALL rounds (total frames / total time), not the best round
    e2e_calls = args.e2e_steps * len(e2e_times)
    e2e_value = world * e2e_frames * e2e_calls / sum(e2e_times)
    link_gbs = world * e2e_calls * (up_bytes + down_bytes) / sum(e2e_times) / 1e9
    bare_gbs = world * e2e_calls * (up_bytes + down_bytes) / sum(bare_times) / 1e9
    e2e_round_values = [world * e2e_frames * args.e2e_steps / t for t in e2e_times]
    # the end-to-end result must be the resident result (same frames)
    assert np.array_equal(np_out[2], mask[:e2e_frames].cpu().numpy()), 'e2e mask differs from the resident run'
    # the same call the way the reference makes it: ordinary (pageable) NumPy arrays in, freshly allocated arrays out
    pageable_value = None
    if world == 1:
        pg_power = np.array(np_power)                                  # pageable copy
        pg_steps = max(args.e2e_steps, 10)
        for _ in range(4):                                             # the first calls also grow the heap for the outputs
            path.mfcc_energy(pg_power, flip=True, normalize_first=True)
        t0 = time.perf_counter()
        for _ in range(pg_steps):
            pg_out = path.mfcc_energy(pg_power, flip=True, normalize_first=True)
        pageable_value = e2e_frames * pg_steps / (time.perf_counter() - t0)
        assert np.array_equal(pg_out[2], np_out[2]), 'pageable e2e mask differs from the pinned run'
        del pg_power, pg_out
    del h_power, h_out

    configs = {}
    if not args.no_configs and half >= 300:          # collective at N > 1: every rank takes the same branch
        configs['C4_clip_stream'] = config_c4(torch, dist, aig, path, dev, power, half, rank, world, mask_blocks)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world == 1 and not args.no_configs:
        configs['C1_batch16'] = config_c1(torch, aig, path, stream_a, stream_b, dev)
        configs['C3_flickr_5k'] = config_c3(torch, aig, path, dev)
        configs['loader_tfrecord'] = config_loader(torch, aig, path, dev)

    # ---- roofline of the dominant kernel (fused MFCC + energy), event-timed per launch in the timed region ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    mfcc_ms, mfcc_launches = prof['mfcc']
    energy_ms, energy_launches = prof['energy']
    frames_per_launch = frames * args.steps / max(mfcc_launches, 1)
    achieved = (ALGO_BYTES_MFCC_KERNEL * frames_per_launch) / (mfcc_ms / max(mfcc_launches, 1) / 1e3) / 1e9
    traffic, traffic_source = None, None
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        traffic = t['fused_kernel_bytes_per_frame'] * frames_per_launch
        traffic_source = ('NOT measured in this run: dram__bytes_read + dram__bytes_write per frame of one `ncu --set full` capture '
                          'of this kernel (profiles/ncu_traffic.json: %s), scaled to the frames of one launch' % t.get('source', 'see file'))
    except Exception:
        pass
    roofline = {
        'bound': 'hbm', 'kernel': 'mfcc_energy_fused_kernel (fused MFCC + energy persistent kernel)', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
        'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_source,
        'peak_source': 'MEASURED_PEAKS.json hbm_gbs (measured copy)' if peaks else 'fallback 6650 GB/s',
        'frac_of_nominal_8TBs': achieved / 8000.0,
        # context: a pure-read stream tops out at 7.40 TB/s on this pool and any stream that also writes 2.3 % of its bytes
        # (this kernel's MFCC output) at 6.9 TB/s - tools/read_peak.cu, profiles/r01_read_peak_probe*.txt
        'frac_of_read_plus_2pct_write_probe_6900': achieved / 6900.0,
        'algorithmic_bytes_per_frame': ALGO_BYTES_MFCC_KERNEL,
        'frames_per_launch': frames_per_launch, 'launches': mfcc_launches,
        'avg_launch_ms': mfcc_ms / max(mfcc_launches, 1),
        'kernel_share_of_step': mfcc_ms / elapsed_ms,
        'energy_kernel_ms_per_launch': energy_ms / max(energy_launches, 1),
        'path_gbs_whole_step': ALGO_BYTES_PATH * frames * args.steps / (elapsed_ms / 1e3) / 1e9,
    }
    if sustained is not None:
        sustained['frac'] = sustained['kernel_achieved_gbs'] / peak
        sustained['frac_of_burst_roofline_frac'] = sustained['frac'] / roofline['frac']
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(frames, world),
        'roofline': roofline,
        'sustained': sustained,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': e2e_frames * IN_BYTES,
                'd2h_bytes_per_step': e2e_frames * OUT_BYTES,
                'frames_per_step': e2e_frames, 'steps': e2e_calls, 'rounds': len(e2e_times),
                'statistic': 'all rounds: frames of every timed call / their summed time (max over ranks per round)',
                'round_values': e2e_round_values,
                'api': 'AcousticPath.mfcc_energy(pinned numpy) -> aig_mfcc_energy, synchronous',
                'link_gbs': link_gbs, 'bare_link_gbs': bare_gbs, 'frac_of_bare_link': link_gbs / bare_gbs,
                'bare_link_probe': 'the same bytes as bare pinned cudaMemcpyAsync (64 MiB H2D pieces, result-sized D2H on a second '
                                   'stream), all %d rank(s) at once; link_gbs and bare_link_gbs are whole-job H2D + D2H GB/s' % world,
                'host_placement_rank0': placement,
                'pageable_numpy_value': pageable_value},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'result': {'auc': auc, 'num': int(host_counts[-1]), 'pos': [int(v) for v in host_counts[:-1]],
                   'counts_equal_rank0_recomputation': counts_ok},
        'configs': configs,
    }
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = time_cpu(10 ** 6, 1, budget_s=args.cpu_seconds, streams=(stream_a, stream_b))[0]
        # the CPU sample is frames 0..7 of both streams - ring frames 0..7 and half..half+7 of this rank (same generator, same
        # seeds): the timed GPU pass must have produced the reference's results for them
        ref_mfcc, ref_mask = cpu_chain.last
        k = CPU_SAMPLE_FRAMES // 2
        gpu_mfcc = np.concatenate([mfcc[:k].cpu().numpy(), mfcc[half:half + k].cpu().numpy()], 0)
        gpu_mask = np.concatenate([mask[:k].cpu().numpy(), mask[half:half + k].cpu().numpy()], 0)
        differing = np.argwhere(gpu_mask != ref_mask)
        line['parity_check'] = {
            'frames': 2 * k, 'against': 'NumPy oracle on the same host-generated frames (CPU baseline sample)',
            'mfcc_max_abs_err': float(np.abs(gpu_mfcc - ref_mfcc).max()), 'mfcc_tolerance': 1e-4,
            'mask_pixels_differing': int(len(differing)), 'mask_pixels': int(gpu_mask.size),
            'boundary_pixels': [[int(v) for v in idx] for idx in differing[:20]]}
        assert line['parity_check']['mfcc_max_abs_err'] <= 1e-4, line['parity_check']
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# GPU arm, ONE process driving N GPUs (the shape of the reference's callers)
# --------------------------------------------------------------------------------------------------
def run_single_process(args):
    import torch
    import acoustic_image_generation_b200 as aig
    from acoustic_image_generation_b200.group import AcousticPathGroup

    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm')
    n_dev = min(args.gpus, torch.cuda.device_count())
    group = AcousticPathGroup(list(range(n_dev)), thresholds=THRESHOLDS)
    frames = args.frames                             # per device
    half = frames // 2
    stream_a, stream_b = base_streams(RING_BLOCK)
    rings, outs = [], []
    for d in range(n_dev):
        dev = torch.device('cuda', d)
        a, b = torch.from_numpy(stream_a).to(dev), torch.from_numpy(stream_b).to(dev)
        power = torch.empty((frames, 36, 48, FFT_LEN), device=dev, dtype=torch.float32)
        for f0 in range(0, half, RING_BLOCK):
            power[f0:f0 + min(RING_BLOCK, half - f0)].copy_(a[:min(RING_BLOCK, half - f0)])
        idx_b = (torch.arange(frames - half, device=dev) + d) % RING_BLOCK
        for f0 in range(0, frames - half, 256):
            sel = idx_b[f0:f0 + 256]
            power[half + f0:half + f0 + len(sel)].copy_(b[sel])
        rings.append(power)
        outs.append((torch.empty((frames, 36, 48, MFCC_NUM), device=dev, dtype=torch.float32),
                     torch.empty((frames, 36, 48), device=dev, dtype=torch.float64),
                     torch.empty((frames, 36, 48), device=dev, dtype=torch.uint8)))
        del a, b

    def device_steps(i, steps):
        """`steps` steps on device i from its own host thread; returns the device-timed milliseconds."""
        torch.cuda.set_device(i)
        p, power, (mfcc, energy, mask) = group.paths[i], rings[i], outs[i]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            p.mfcc_energy(power, flip=True, normalize_first=True, out=(mfcc, energy, mask))
            p.iou_sweep(mask[:half], mask[half:2 * half], group._thr[i], pos=group.counts[i][:-1], num=group.counts[i][-1:])
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    def all_devices(steps):
        futures = [group._pool.submit(device_steps, i, steps) for i in range(n_dev)]
        return max(f.result() for f in futures)

    all_devices(args.warmup)
    group.reset_counts()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = group.launch_count
    elapsed_ms = all_devices(args.steps)
    clocks = sampler.result()
    launches = group.launch_count - launches0
    totals = group.reduce_counts()                   # one grouped NCCL all-reduce (or a host sum without NCCL)
    value = n_dev * frames * args.steps / (elapsed_ms / 1e3)
    masks0 = (outs[0][2][:RING_BLOCK].clone(), outs[0][2][half:half + RING_BLOCK].clone())
    counts_ok = bool(np.array_equal(totals, expected_counts(torch, masks0[0], masks0[1], half, n_dev, args.steps)))

    # end to end: ONE host call on pinned NumPy arrays, sharded over the devices by the group
    e2e_frames = min(args.e2e_frames, frames) * n_dev
    pin = lambda shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True)
    h_power = pin((e2e_frames, 36, 48, FFT_LEN), torch.float32)
    for lo in range(0, e2e_frames, RING_BLOCK):
        n = min(RING_BLOCK, e2e_frames - lo)
        h_power[lo:lo + n].copy_(torch.from_numpy(stream_a[:n]))
    h_out = (pin((e2e_frames, 36, 48, MFCC_NUM), torch.float32), pin((e2e_frames, 36, 48), torch.float64),
             pin((e2e_frames, 36, 48), torch.uint8))
    np_power, np_out = h_power.numpy(), tuple(t.numpy() for t in h_out)
    for _ in range(2):
        group.mfcc_energy(np_power, flip=True, normalize_first=True, out=np_out)
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        group.mfcc_energy(np_power, flip=True, normalize_first=True, out=np_out)
    e2e_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        group.mfcc_energy(np_power, flip=True, normalize_first=True, out=np_out, chunk_frames=None)
    e2e_equal_s = time.perf_counter() - t0
    single = group.paths[0].mfcc_energy(np_power[:RING_BLOCK], flip=True, normalize_first=True)
    same = all(np.array_equal(x, y[:RING_BLOCK]) for x, y in zip(single, np_out))
    rates = aig.success_rates(totals[:-1], max(int(totals[-1]), 1))
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': n_dev, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(frames, n_dev, single_process=True),
        'mode': 'single-process: AcousticPathGroup, one host thread and one libaig handle per device',
        'counts_merged_by': 'aig_group_allreduce_counts (ncclCommInitAll)' if group.nccl else 'host sum',
        'e2e': {'value': e2e_frames * args.e2e_steps / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': e2e_frames * IN_BYTES,
                'd2h_bytes_per_step': e2e_frames * OUT_BYTES, 'frames_per_step': e2e_frames, 'steps': args.e2e_steps,
                'api': 'AcousticPathGroup.mfcc_energy(pinned numpy): one host call, frames spread over %d devices in 64-frame pieces '
                       'taken dynamically' % n_dev,
                'value_with_equal_shards': e2e_frames * args.e2e_steps / e2e_equal_s,
                'sharded_result_equals_one_gpu_result': bool(same)},
        'gpu_launches': int(launches), 'clocks': clocks,
        'result': {'auc': aig.auc(THRESHOLDS, rates), 'num': int(totals[-1]), 'pos': [int(v) for v in totals[:-1]],
                   'counts_equal_rank0_recomputation': counts_ok},
    }
    group.close()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--frames', type=int, default=8192,
                    help='resident frames per step per GPU')
    ap.add_argument('--e2e-frames', type=int, default=256, help='frames per end-to-end step (pinned host batch)')
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--e2e-rounds', type=int, default=3, help='library call / bare-copy probe alternations (all of them counted)')
    ap.add_argument('--sustain-seconds', type=float, default=3.2, help='length of the sustained loop (0 = skip)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='CPU baseline time budget')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the C1 / C3 / C4 / loader sub-records')
    ap.add_argument('--single-process', action='store_true', help='one process drives all --gpus devices (AcousticPathGroup)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3                                 # timing rule: at least three warm-up steps
    if args.impl == 'reference':
        run_reference(args)
    elif args.single_process:
        run_single_process(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
