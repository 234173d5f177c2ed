#!/usr/bin/env python
"""bench.py - throughput of the fused DSWx-HLS classification pass on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
synthetic HLS S30 tiles, 3660 x 3660, full product - 6 int16 bands + Fmask +
float32 DEM (50-px margin) + LAND + ocean mask -> WTR, BWTR, CONF, DIAG and the
coverage counters.  One *step* = one launch of the fused kernel over a batch of
``--tiles`` distinct device-resident tiles per GPU (default 16 = 5.1 GB of
traffic per step, far larger than the 126 MB L2, so nothing is cache-hot
between steps).  Tiles are sharded by tile over the ranks with no data-path
communication (weak scaling: the per-GPU batch is fixed).

One JSON line on stdout (rank 0).  ``value`` = Mpixel/s with inputs resident in
HBM; ``e2e`` = the same metric through the host-buffer API
(proteus_b200.classify_tile: pinned numpy arrays in, numpy layers out, H2D and
D2H inside the timed region); ``roofline`` = algorithmic bytes / kernel time
against the measured HBM copy peak; ``cpu_baseline`` = the numpy restatement of
the reference chain (oracle/, kind "port") timed on one host core on one tile.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 3660
PIXELS_PER_TILE = TILE * TILE
BYTES_IN_PER_PX = 6 * 2 + 1 + 4 + 1 + 1      # bands + fmask + DEM + LAND + ocean  
BYTES_OUT_PER_PX = 1 + 1 + 1 + 2             # WTR + BWTR + CONF + DIAG
ALGO_BYTES_PER_PX = BYTES_IN_PER_PX + BYTES_OUT_PER_PX      # 24
METRIC = 'dswx_hls_classification_throughput'
UNIT = 'Mpixel/s'
FALLBACK_HBM_GBS = 6650.0                    # /opt/skills/guides/B200_PROFILING.md


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', choices=('ours', 'reference'), default='ours')
    ap.add_argument('--tiles', type=int, default=16, help='tiles per GPU per step')
    ap.add_argument('--workload', choices=('tiles', 'timeseries'), default='tiles',
                    help="tiles: configs[1] batch sharded by tile (default); timeseries: the acquisitions of --tiles "
                         "share DEM/LAND/ocean and the e2e leg keeps them on the device.  The other BASELINE configs "
                         "(mosaic, 64-tile batch, 365-acquisition series, worst case, config 0) ride on the same line "
                         "as secondary records unless --no-extras")
    ap.add_argument('--no-extras', action='store_true', help='headline workload only')
    ap.add_argument('--series', type=int, default=365, help='acquisitions of the time-series record')
    ap.add_argument('--sustained-s', type=float, default=2.0,
                    help='seconds of back-to-back launches for roofline.sustained (0 = skip)')
    ap.add_argument('--mosaic-size', type=int, default=21960)
    ap.add_argument('--size', type=int, default=TILE, help='tile edge in pixels (default 3660)')
    ap.add_argument('--e2e-steps', type=int, default=0, help='host-path steps (default: min(steps, 20))')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cpu-rows', type=int, default=0,
                    help='rows of the tile the CPU baseline processes (default: whole tile)')
    return ap.parse_args()


# ---------------------------------------------------------------------------
# clocks (pynvml; sampled DURING the timed regions)
# ---------------------------------------------------------------------------
_REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown',
            0x40: 'hw_thermal_slowdown', 0x80: 'hw_power_brake_slowdown',
            0x2: 'applications_clocks_setting', 0x10: 'sync_boost'}


class ClockSampler:
    """SM clock + throttle reasons of this rank's GPU, sampled by a thread WHILE a timed region runs; one record per
    named segment (start(name) ... pause())."""

    def __init__(self, device_index, period_s=float(os.environ.get('PB200_BENCH_NVML_PERIOD', '0.02'))):
        self.segments, self.max_mhz = {}, None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._current = None
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            except Exception:
                pass
            self.h = None
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(h)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        self.h = h
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.period = period_s
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:                       # pragma: no cover
            self.nv = None
            self.error = repr(e)

    def _sample(self, seg):
        nv = self.nv
        try:
            seg['samples'].append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            try:
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in _REASONS.items():
                if mask & bit:
                    seg['reasons'].add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            seg = self._current
            if self._active.is_set() and seg is not None:
                self._sample(seg)
            time.sleep(self.period)

    def start(self, name='tiles'):
        self._current = self.segments.setdefault(name, {'samples': [], 'reasons': set()})
        if self.nv:
            self._sample(self._current)               # at least one sample even for a region shorter than the period
        self._active.set()

    def pause(self):
        seg = self._current
        if self.nv and seg is not None:
            self._sample(seg)                          # ... and one at its end, still under load
        self._active.clear()

    def stop(self):
        self._stop.set()

    def summary(self, name='tiles'):
        seg = self.segments.get(name)
        if not self.nv or not seg or not seg['samples']:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(seg['samples']), 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(seg['reasons']), 'samples': len(seg['samples'])}


# ---------------------------------------------------------------------------
# host memory placement for the e2e leg: pinned buffers on the GPU's own NUMA node
# ---------------------------------------------------------------------------
_NUMA_NOTE = 'not bound'


def _bind_to_gpu_numa_node(device_index):
    """Before any pinned allocation: prefer host memory of the NUMA node the GPU hangs off (set_mempolicy
    MPOL_PREFERRED) and, when the process may run there, its CPUs.  At N = 8 every rank otherwise allocates its
    256 MB + 67 MB of pinned buffers on the node it happens to start on and half the H2D traffic crosses the socket
    interconnect.  Best effort; the outcome is recorded in e2e.numa."""
    global _NUMA_NOTE
    try:
        import ctypes
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = None
        for i in range(pynvml.nvmlDeviceGetCount()):
            hh = pynvml.nvmlDeviceGetHandleByIndex(i)
            u = pynvml.nvmlDeviceGetUUID(hh)
            u = u.decode() if isinstance(u, bytes) else u
            if uuid in u:
                h = hh
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(':')[0]) == 8:
            bus = bus[4:]
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            _NUMA_NOTE = 'GPU reports no NUMA node'
            return
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = set()
            for part in f.read().strip().split(','):
                a, _, b = part.partition('-')
                cpus.update(range(int(a), int(b or a) + 1))
        note = [f'GPU {device_index} on NUMA node {node}']
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL(None, use_errno=True)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(8 * ctypes.sizeof(mask)))   # set_mempolicy(MPOL_PREFERRED)
        note.append('memory policy: preferred' if rc == 0 else f'set_mempolicy failed (errno {ctypes.get_errno()})')
        allowed = os.sched_getaffinity(0)
        local = allowed & cpus
        if local:
            os.sched_setaffinity(0, local)
            note.append(f'{len(local)} local CPUs')
        else:
            note.append('no local CPU in the cpuset')
        _NUMA_NOTE = '; '.join(note)
    except Exception as e:
        _NUMA_NOTE = f'not bound ({e!r})'


# ---------------------------------------------------------------------------
# CPU baseline: the numpy restatement of the reference chain (oracle/, "port")
# ---------------------------------------------------------------------------
def _oracle_chain_on_rows(tile, r0, r1):
    from oracle import dswx_oracle as O              # allowed here: cpu_baseline / reference arm only
    m = tile['dem_margin']
    return O.reference_chain(
        [b[r0:r1] for b in tile['bands']], tile['fmask'][r0:r1],
        tile['dem'][r0:r1 + 2 * m] if tile['dem'] is not None else None,
        tile['land'][r0:r1] if tile['land'] is not None else None,
        tile['ocean'][r0:r1] if tile['ocean'] is not None else None,
        tile['sun_azimuth'], tile['sun_elevation'])


def cpu_baseline_one_core(tile, rows):
    """Time the port on one core over the first ``rows`` rows of the tile."""
    h, w = tile['fmask'].shape
    rows = min(rows or h, h)
    t0 = time.perf_counter()
    out = _oracle_chain_on_rows(tile, 0, rows)
    dt = time.perf_counter() - t0
    mpx = rows * w / 1e6
    return out, {'value': mpx / dt, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                 'seconds': dt,
                 'sample': f'rows 0..{rows} of one synthetic S30 tile ({rows}x{w} px, full product), '
                           'numpy restatement of dswx_hls.py:5088-5369 (oracle/dswx_oracle.py), 1 process'}


_G_TILE = None


def _pool_job(args):
    r0, r1 = args
    out = _oracle_chain_on_rows(_G_TILE, r0, r1)
    return int(out['counters'][0]), int(out['WTR'].astype('int64').sum())


def _load_synth_without_the_library():
    """proteus_b200/synth.py by path: importing the package would map libproteus_b200.so into this process, and the
    reference arm must not hold any product code (only numpy + oracle/)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('pb200_synth_standalone',
                                                  os.path.join(ROOT, 'proteus_b200', 'synth.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# Per core, the numpy port (oracle/dswx_oracle.py: look-up tables instead of one pass per class, explicit differences
# instead of np.gradient) is faster than the reference's own functions called in generate_dswx_layers' order; measured
# in the build container (8 vCPU) on the same synthetic tile by scripts/port_vs_reference.py ->
# profiles/r2_port_vs_reference.json.  The reference arm therefore flatters the CPU side by about this factor.
PORT_VS_REFERENCE_PER_CORE = None


def _port_vs_reference():
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_port_vs_reference.json')) as f:
            d = json.load(f)
        return {'ratio': d['port_vs_reference_per_core'], 'source': 'profiles/r2_port_vs_reference.json '
                '(scripts/port_vs_reference.py: live unmodified reference vs the port, same tile, one core, '
                f"build container, {d.get('rows')} rows)"}
    except Exception:
        return {'ratio': 4.0, 'source': 'SURVEY.md section 6 probe: reference 1.65 Mpixel/s/core vs port 6.6'}


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm (numpy port; the
    reference itself is not installable here and does not exist on the GPU
    box) on all host cores; one step = ONE WHOLE tile of the GPU arm's
    config, split into row strips over the processes."""
    global _G_TILE
    import multiprocessing as mp
    import numpy as np
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    synth = _load_synth_without_the_library()
    assert 'proteus_b200' not in sys.modules, 'the reference arm must not import the product package'
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    size = args.size
    rows = args.cpu_rows or size                     # default: the whole tile (same config as the GPU arm)
    tile = synth.make_tile(0, size, size)
    _G_TILE = tile
    n_procs = max(1, min(cores, rows // 8))
    # 4 strips per process: the fill wedge and the water share differ between strips, smaller strips balance the pool
    n_jobs = max(1, min(rows // 8, 4 * n_procs))
    bounds = np.linspace(0, rows, n_jobs + 1).astype(int)
    jobs = [(int(bounds[i]), int(bounds[i + 1])) for i in range(n_jobs) if bounds[i + 1] > bounds[i]]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ctx = mp.get_context('fork')
    with ctx.Pool(n_procs) as pool:
        for _ in range(warm):
            pool.map(_pool_job, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(_pool_job, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    mpx = rows * size / 1e6
    value = mpx * steps / dt
    pvr = _port_vs_reference()
    sample = (f'{steps} steps x {"the whole" if rows == size else f"rows 0..{rows} of one"} synthetic S30 {size}x{size} tile '
              f'(full product), {len(jobs)} row strips over {n_procs} processes; numpy port of dswx_hls.py:5088-5369')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warm, 'ms_per_step': 1e3 * dt / steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int16', 'data': 'synthetic',
        'tiles_per_s': value * 1e6 / (size * size),
        'config': _config_block(size, args.tiles, args.gpus, args.workload),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': n_procs, 'kind': 'port', 'sample': sample,
                         'port_vs_reference_per_core': pvr['ratio'], 'port_vs_reference_source': pvr['source']},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'reference is pure Python + GDAL; GDAL/yamale are absent so it cannot be pip-installed; '
                'its per-pixel numpy chain is timed through the validated port (oracle/), which is '
                f"~{pvr['ratio']:.1f}x faster per core than the reference's own functions: the ratio against this arm is conservative",
    }
    print(json.dumps(line), flush=True)
    return 0


def _config_block(size, n_tiles, world, workload='tiles'):
    """The `config` object: identical in both arms (the driver compares them)."""
    px = size * size
    return {
        'workload': ('configs[3] time series (shared DEM/LAND/ocean): ' if workload == 'timeseries' else '') +
                    f'configs[1]: synthetic HLS S30 tile {size}x{size} full product '
                    f'(6 int16 bands + Fmask + DEM + LAND + ocean -> WTR/BWTR/CONF/DIAG + counters), '
                    f'{n_tiles} distinct device-resident tiles per GPU per step, sharded by tile',
        'tile': [size, size], 'tiles_per_gpu_per_step': n_tiles,
        'layers_out': ['WTR', 'BWTR', 'CONF', 'DIAG'], 'bytes_per_pixel': ALGO_BYTES_PER_PX,
        'l2_policy': f'inputs per step {n_tiles * px * BYTES_IN_PER_PX / 1e9:.2f} GB >> 126 MB L2; no flush needed',
        'parallelism': f'tile-sharded x{world}, no data-path collective'}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def _peak():
    peak, src = FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peak = float(json.load(f)['hbm_gbs'])
            src = 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    except Exception:
        pass
    return peak, src


class _Env:
    """rank / world / barrier / max-over-ranks helpers shared by the workloads."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.dev = f'cuda:{self.local_rank}'

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device='cuda')
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def free(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def _time_launches(env, sampler, segment, run, steps, warmup=3):
    """W warm-up + K timed calls of ``run`` on the current stream: CUDA events on the launching stream, barrier +
    synchronize on both sides, clocks sampled during the timed region.  Returns (ms on this rank, max over ranks)."""
    torch = env.torch
    stream = torch.cuda.current_stream()
    for _ in range(max(warmup, 3)):
        run()
    env.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start(segment)
    ev0.record(stream)
    for _ in range(steps):
        run()
    ev1.record(stream)
    env.barrier()
    sampler.pause()
    ms = ev0.elapsed_time(ev1)
    return ms, env.max_over_ranks(ms)


def run_mosaic(env, args, sampler, steps):
    """BASELINE configs[4]: one mosaic-size^2 raster row-stripped over the ranks, one DEM halo row per neighbour
    exchanged with NCCL send/recv every step (through the C ABI's pb200_halo_exchange_dem)."""
    import proteus_b200 as pb
    from proteus_b200 import mosaic, synth
    torch = env.torch
    size, m, rank, world, dev = args.mosaic_size, 50, env.rank, env.world, env.dev
    r0, r1 = mosaic.strip_bounds(size, world)[rank]
    n = r1 - r0
    t = synth.make_device_batch(1, n, size, device=dev, seed=2000 + rank, n_distinct=1, full_product=False)[0]
    g = torch.Generator(device=dev)
    g.manual_seed(3000 + rank)
    d0, d1 = mosaic.dem_rows_for_strip(r0, r1, size, m)
    coarse = torch.randn((1, 1, (d1 - d0) // 25 + 2, (size + 2 * m) // 25 + 2), device=dev, generator=g)
    dem_local = (torch.nn.functional.interpolate(coarse, size=(d1 - d0, size + 2 * m), mode='bilinear',
                                                 align_corners=True)[0, 0] * 400.0 + 800.0).contiguous()
    land = torch.full((n, size), 255, dtype=torch.uint8, device=dev)
    land[:, ::7] = 21
    land[::5] = 200
    ocean = torch.ones((n, size), dtype=torch.uint8, device=dev)
    ocean[:, : size // 10] = 0
    ctx = pb.get_context(env.local_rank)
    exchange = 'torch'
    if world > 1:
        try:
            if not getattr(ctx, '_pb200_comm', False):
                mosaic.init_library_comm(ctx, rank, world)
                ctx._pb200_comm = True
            exchange = 'library'
        except Exception as e:                       # NCCL not loadable through dlopen: torch.distributed P2P
            sys.stderr.write(f'[bench] library communicator unavailable ({e}); using torch.distributed P2P\n')
    out = {}
    for overlap in ((True, False) if world > 1 else (False,)):
        strip = mosaic.MosaicStrip(t['bands'], t['fmask'], dem_local, land, ocean, r0, r1, size,
                                   sun_azimuth=150.0, sun_elevation=45.0, params=pb.make_params(),
                                   outputs=pb.GRADED_LAYERS, rank=rank, world=world, ctx=ctx,
                                   exchange=exchange, overlap=overlap)
        seg = f'mosaic_overlap{int(overlap)}'
        ms, ms_max = _time_launches(env, sampler, seg, strip.run, steps, args.warmup)
        ms_per_step = ms_max / steps
        achieved = n * size * ALGO_BYTES_PER_PX / (ms / steps / 1e3) / 1e9
        out['overlap' if overlap else 'exchange_then_one_launch'] = {
            'value': size * size / 1e6 / (ms_per_step / 1e3), 'unit': UNIT, 'ms_per_step': ms_per_step,
            'launches_per_step': 2 if overlap else 1,
            'roofline_frac_this_rank': achieved / _peak()[0], 'clocks': sampler.summary(seg)}
        del strip
        env.free()
    best = max(out, key=lambda k: out[k]['value'])
    res = dict(out[best])
    res.update({
        'workload': f'configs[4]: one synthetic {size}x{size} mosaic (full product) row-stripped over {world} GPU(s); '
                    'every step exchanges one DEM halo row per neighbour with NCCL send/recv '
                    f'({"pb200_halo_exchange_dem" if exchange == "library" else "torch.distributed P2P"}) and '
                    'classifies the strip',
        'scaling': 'strong', 'steps': steps, 'schedule': best, 'schedules': out,
        'rows_per_rank': n, 'halo_bytes_per_neighbour': (size + 2 * m) * 4, 'exchange': exchange,
        'tiles_per_s': res['value'] * 1e6 / PIXELS_PER_TILE})
    del t, dem_local, land, ocean
    env.free()
    return res


def run_batch64(env, args, sampler, steps):
    """BASELINE configs[2] as worded: a FIXED batch of 64 tiles sharded by tile over the ranks (strong scaling)."""
    import proteus_b200 as pb
    from proteus_b200 import mosaic, synth
    size, total = args.size, 64
    mine = mosaic.shard_tiles(total, env.rank, env.world)
    tiles = synth.make_device_batch(len(mine), size, size, device=env.dev, seed=1000 + env.rank, n_distinct=min(4, len(mine)))
    plan = pb.Plan(tiles, pb.make_params(collapse_wtr_classes=True), pb.GRADED_LAYERS)
    stream = env.torch.cuda.current_stream()
    ms, ms_max = _time_launches(env, sampler, 'batch64', lambda: plan.run(stream), steps, args.warmup)
    ms_per_step = ms_max / steps
    res = {'workload': f'configs[2]: fixed batch of {total} synthetic S30 tiles {size}x{size} (full product), tile i on '
                       f'rank i mod {env.world}, no communication; one step = the whole batch',
           'scaling': 'strong', 'tiles_total': total, 'tiles_per_gpu': len(mine), 'steps': steps,
           'value': total * size * size / 1e6 / (ms_per_step / 1e3), 'unit': UNIT, 'ms_per_step': ms_per_step,
           'tiles_per_s': total / (ms_per_step / 1e3),
           'roofline_frac_this_rank': len(mine) * size * size * ALGO_BYTES_PER_PX / (ms / steps / 1e3) / 1e9 / _peak()[0],
           'clocks': sampler.summary('batch64')}
    plan.close()
    del plan, tiles
    env.free()
    return res


def run_timeseries(env, args, sampler):
    """BASELINE configs[3] as worded: 365 acquisitions of one MGRS tile (shared DEM / LAND / ocean), all device
    resident, sharded by acquisition over the ranks; roofline sweep over the number of acquisitions per launch."""
    import proteus_b200 as pb
    from proteus_b200 import mosaic, synth
    torch = env.torch
    size, total = args.size, args.series
    mine = mosaic.shard_tiles(total, env.rank, env.world)
    tiles = synth.make_device_batch(len(mine), size, size, device=env.dev, seed=1000 + env.rank, n_distinct=min(4, len(mine)),
                                    shared_ancillary=True)
    params = pb.make_params(collapse_wtr_classes=True)
    # outputs of the whole series stay resident too: every plan of the sweep writes into these
    outs = [{n: torch.empty((size, size), device=env.dev, dtype=torch.int16 if n == 'DIAG' else torch.uint8)
             for n in pb.GRADED_LAYERS} for _ in tiles]
    counters = torch.zeros((len(tiles), 12), dtype=torch.int64, device=env.dev)
    stream = torch.cuda.current_stream()
    peak = _peak()[0]
    points, sweep = [], []
    b = 1
    while b < len(tiles):
        points.append(b)
        b *= 2
    points.append(len(tiles))
    px = size * size
    for b in points:
        plans = [pb.Plan(tiles[i:i + b], params, pb.GRADED_LAYERS, outputs_into=outs[i:i + b],
                         counters_into=counters[i:i + b]) for i in range(0, len(tiles) - b + 1, b)][:max(1, min(8, len(tiles) // b))]
        # enough launches for >= 0.25 s per point (>= 10 clock samples), cycling over distinct acquisitions
        per = b * 75e-6
        reps = max(3, int(0.25 / (per * len(plans))))

        def run_all():
            for pl in plans:
                pl.run(stream)
        seg = f'series_b{b}'
        ms, ms_max = _time_launches(env, sampler, seg, run_all, reps, 3)
        n_px = b * len(plans) * reps * px
        gpx = env.sum_over_ranks(n_px) / 1e9 / (ms_max / 1e3)
        sweep.append({'acquisitions_per_launch': b, 'launches': reps * len(plans), 'Gpixel_per_s': gpx,
                      'roofline_frac_this_rank': n_px * ALGO_BYTES_PER_PX / (ms / 1e3) / 1e9 / peak,
                      'clocks': sampler.summary(seg)})
        for pl in plans:
            pl.close()
        del plans
    whole = sweep[-1]
    res = {'workload': f'configs[3]: time series of {total} acquisitions of one synthetic MGRS tile {size}x{size} (bands + '
                       f'Fmask per acquisition, DEM / LAND / ocean shared), all inputs and outputs device resident '
                       f'({len(tiles) * px * (13 + 5) / 1e9:.1f} GB on this GPU), acquisitions sharded over {env.world} GPU(s)',
           'scaling': 'strong', 'acquisitions_total': total, 'acquisitions_per_gpu': len(tiles),
           'value': whole['Gpixel_per_s'] * 1e3, 'unit': UNIT, 'tiles_per_s': whole['Gpixel_per_s'] * 1e9 / px,
           'sweep': sweep}
    del tiles, outs, counters
    env.free()
    return res


def run_small_batch(env, args, sampler, steps, *, name, adversarial=False, full_product=True, layers=None, bytes_per_px=ALGO_BYTES_PER_PX,
                    wedge=0.08):
    import proteus_b200 as pb
    from proteus_b200 import synth
    size, n_tiles = args.size, args.tiles
    tiles = synth.make_device_batch(n_tiles, size, size, device=env.dev, seed=4000 + env.rank, n_distinct=min(4, n_tiles),
                                    adversarial=adversarial, full_product=full_product, wedge=wedge)
    fill_share = float(sum(float((t['fmask'] == 255).float().mean()) for t in tiles[:4]) / min(4, len(tiles)))
    plan = pb.Plan(tiles, pb.make_params(collapse_wtr_classes=True), layers or pb.GRADED_LAYERS)
    stream = env.torch.cuda.current_stream()
    ms, ms_max = _time_launches(env, sampler, name, lambda: plan.run(stream), steps, args.warmup)
    ms_per_step = ms_max / steps
    res = {'value': env.world * n_tiles * size * size / 1e6 / (ms_per_step / 1e3), 'unit': UNIT, 'ms_per_step': ms_per_step,
           'steps': steps, 'bytes_per_pixel': bytes_per_px,
           'roofline_frac_this_rank': n_tiles * size * size * bytes_per_px / (ms / steps / 1e3) / 1e9 / _peak()[0],
           'clocks': sampler.summary(name), 'fill_share': round(fill_share, 4)}
    plan.close()
    del plan, tiles
    env.free()
    return res


def copy_ceiling_ms(env, size, reuse):
    """What the box allows for the e2e copy pattern alone: the H2D bytes of a tile in the host pipeline's strips plus
    the concurrent D2H of the four graded layers, tiles back to back, no kernel.  Returns (mean, best) of 5 runs of 6
    tiles, per tile, each the max over ranks: with 4 or 8 ranks the host side of the box saturates and the runs scatter
    (N = 8: best 11.6, mean 16.3 ms, profiles/r2_e2e_probe_n8.json) - the mean is what a stream of tiles sees."""
    import numpy as np
    import proteus_b200 as pb
    torch = env.torch
    edges = [0, 1024, 2048, 3072, 3392, size] if size == TILE else [0, size]

    def pinned(shape, dt):
        return torch.from_numpy(pb.pinned_empty(shape, dt))
    ins = [pinned((size, size), np.int16) for _ in range(6)] + [pinned((size, size), np.uint8) for _ in range(1 if reuse else 3)]
    dem = None if reuse else pinned((size + 100, size + 100), np.float32)
    outs = [pinned((size, size), np.int16)] + [pinned((size, size), np.uint8) for _ in range(3)]
    for x in ins + outs + ([dem] if dem is not None else []):
        x.zero_()
    dins = [torch.empty_like(x, device=env.dev) for x in ins]
    ddem = torch.empty_like(dem, device=env.dev) if dem is not None else None
    douts = [torch.zeros_like(x, device=env.dev) for x in outs]
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def once(n_tiles=6):
        # n_tiles back to back, like a stream of tiles with two in flight: the H2D engine never waits for the D2H of the
        # previous tile; per-tile time = total / n_tiles
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_tiles):
            evs = []
            with torch.cuda.stream(s_in):
                for a, b in zip(edges[:-1], edges[1:]):
                    for h, d in zip(ins, dins):
                        d[a:b].copy_(h[a:b], non_blocking=True)
                    if dem is not None:
                        ddem[a + 49:b + 51].copy_(dem[a + 49:b + 51], non_blocking=True)
                    e = torch.cuda.Event()
                    e.record(s_in)
                    evs.append(e)
            with torch.cuda.stream(s_out):
                for (a, b), e in zip(zip(edges[:-1], edges[1:]), evs):
                    s_out.wait_event(e)
                    for h, d in zip(outs, douts):
                        h[a:b].copy_(d[a:b], non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / n_tiles
    for _ in range(2):
        once()
    env.barrier()
    runs = [once() for _ in range(5)]
    env.barrier()
    return env.max_over_ranks(sum(runs) / len(runs)), env.max_over_ranks(min(runs))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device - the product has no CPU path to time')
    env = _Env()
    rank, world, local_rank = env.rank, env.world, env.local_rank
    torch.cuda.set_device(local_rank)
    _bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    import proteus_b200 as pb
    from proteus_b200 import synth

    size, n_tiles = args.size, args.tiles
    px_per_tile = size * size
    sampler = ClockSampler(local_rank)
    peak, peak_src = _peak()

    # ---- device-resident batch (weak scaling: n_tiles per rank) -------------
    tiles = synth.make_device_batch(n_tiles, size, size, device=f'cuda:{local_rank}',
                                    seed=1000 + rank, n_distinct=min(4, n_tiles),
                                    shared_ancillary=(args.workload == 'timeseries'))
    params = pb.make_params(collapse_wtr_classes=True)
    plan = pb.Plan(tiles, params, pb.GRADED_LAYERS)
    kernel_name = plan.kernel_name
    stream = torch.cuda.current_stream()

    # cudaProfilerStart/Stop bracket the timed regions: `ncu --profile-from-start off` then lists exactly the
    # kernels launched inside them (profiles/); no effect without a profiler attached
    torch.cuda.cudart().cudaProfilerStart()
    ms_total, ms_total_max = _time_launches(env, sampler, 'tiles', lambda: plan.run(stream), args.steps, args.warmup)
    torch.cuda.cudart().cudaProfilerStop()
    ms_per_step = ms_total_max / args.steps
    mpx_per_step_all = world * n_tiles * px_per_tile / 1e6
    value = mpx_per_step_all / (ms_per_step / 1e3)

    # kernel-only roofline on this rank (the step IS one kernel launch)
    kernel_ms = ms_total / args.steps
    algo_bytes = n_tiles * px_per_tile * ALGO_BYTES_PER_PX
    achieved = algo_bytes / (kernel_ms / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            tj = json.load(f)
            traffic = tj['dram_bytes_per_pixel'] * n_tiles * px_per_tile
    except Exception:
        pass

    # ---- sustained regime: >= args.sustained_s seconds of back-to-back launches ----------------------------------
    sustained = None
    if args.sustained_s > 0:
        n_sus = max(args.steps, int(args.sustained_s * 1e3 / kernel_ms) + 1)
        ms_sus, _ = _time_launches(env, sampler, 'sustained', lambda: plan.run(stream), n_sus, 1)
        ach = algo_bytes / (ms_sus / n_sus / 1e3) / 1e9
        sustained = {'launches': n_sus, 'seconds': ms_sus / 1e3, 'kernel_ms': ms_sus / n_sus, 'achieved': ach,
                     'frac': ach / peak, 'Gpixel_per_s': n_tiles * px_per_tile / (ms_sus / n_sus) / 1e6,
                     'clocks': sampler.summary('sustained'),
                     'note': 'same plan, launches back to back for this long: the board reaches its power cap '
                             '(sw_power_cap) and the SM clock settles below the burst clock of the headline run'}

    # ---- end to end through the host-buffer API ------------------------------
    e2e = None
    parity = None
    cpu_baseline = None
    host_tile = None
    if not args.no_e2e or not args.no_cpu_baseline:
        host_tile = synth.make_tile(rank, size, size)
    if not args.no_e2e:
        pin = dict(bands=[pb.pinned_copy(b) for b in host_tile['bands']],
                   fmask=pb.pinned_copy(host_tile['fmask']), dem=pb.pinned_copy(host_tile['dem']),
                   land=pb.pinned_copy(host_tile['land']), ocean=pb.pinned_copy(host_tile['ocean']))
        outbuf = {n: pb.pinned_empty((size, size), np.uint16 if n == 'DIAG' else np.uint8)
                  for n in pb.GRADED_LAYERS}
        outbuf['counters'] = pb.pinned_empty((12,), np.uint64)

        # time series: the acquisitions of one MGRS tile share DEM / LAND / ocean, which stay on the device
        reuse = args.workload == 'timeseries'

        def host_step(reuse_ancillary=reuse):
            return pb.classify_tile(pin['bands'], pin['fmask'], pin['dem'], pin['land'], pin['ocean'],
                                    host_tile['sun_azimuth'], host_tile['sun_elevation'],
                                    params=params, outputs=pb.GRADED_LAYERS, out=outbuf,
                                    reuse_ancillary=reuse_ancillary)
        e2e_steps = args.e2e_steps or min(args.steps, 20)
        host_res = host_step(False)                     # uploads the ancillary rasters
        for _ in range(3):
            host_res = host_step()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_res = host_step()
        torch.cuda.synchronize()
        dt_single = env.max_over_ranks(time.perf_counter() - t0)

        # the streaming call: two tiles in flight (TilePipeline: enqueue tile k + 1, then wait for tile k), each with
        # its own output buffers; H2D of every tile's inputs and D2H of its layers inside the timed region as before
        outbuf2 = {n: pb.pinned_empty((size, size), np.uint16 if n == 'DIAG' else np.uint8) for n in pb.GRADED_LAYERS}
        outbuf2['counters'] = pb.pinned_empty((12,), np.uint64)
        outs = (outbuf, outbuf2)
        pipe = pb.TilePipeline()

        def submit(i, reuse_ancillary=reuse):
            return pipe.submit(pin['bands'], pin['fmask'], pin['dem'], pin['land'], pin['ocean'],
                               host_tile['sun_azimuth'], host_tile['sun_elevation'], params=params,
                               outputs=pb.GRADED_LAYERS, out=outs[i & 1], reuse_ancillary=reuse_ancillary)
        submit(0, False)
        submit(1, False)                                # each slot uploads its own ancillary rasters once
        for i in range(2, 6):
            submit(i)
        pipe.flush()
        env.barrier()
        sampler.start('e2e')
        torch.cuda.cudart().cudaProfilerStart()
        t0 = time.perf_counter()
        n_done = 0
        for i in range(e2e_steps):
            n_done += submit(i) is not None
        tail = pipe.flush()
        n_done += len(tail)
        host_res = tail[-1]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        torch.cuda.cudart().cudaProfilerStop()
        sampler.pause()
        assert n_done == e2e_steps
        dt = env.max_over_ranks(dt)
        dem_rows_copied = size + 2
        h2d = px_per_tile * (12 + 1) + (0 if reuse else px_per_tile * 2 + dem_rows_copied * (size + 100) * 4)
        d2h = px_per_tile * BYTES_OUT_PER_PX + 12 * 8
        e2e = {'value': world * e2e_steps * px_per_tile / 1e6 / dt, 'unit': UNIT,
               'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'steps': e2e_steps, 'ms_per_tile': 1e3 * dt / e2e_steps,
               'api': 'proteus_b200.TilePipeline.submit (classify_tile on two alternating host pipelines, '
                      'pb200_classify_host_ex + pb200_host_wait): pinned numpy in, numpy out, 1 tile per step per GPU, '
                      'two tiles in flight' + ('; reuse_ancillary=True: DEM / LAND / ocean of the tile stay on '
                                               'the device between acquisitions' if reuse else ''),
               'single_call': {'value': world * e2e_steps * px_per_tile / 1e6 / dt_single, 'ms_per_tile': 1e3 * dt_single / e2e_steps,
                               'api': 'proteus_b200.classify_tile, one synchronous call per tile'},
               'numa': _NUMA_NOTE, 'clocks': sampler.summary('e2e')}
        del pin, outbuf, outbuf2, pipe
        try:
            e2e['ceiling_ms'], e2e['ceiling_best_ms'] = copy_ceiling_ms(env, size, reuse)
            e2e['ceiling_note'] = ('copy-only time per tile of the same H2D + concurrent D2H byte pattern, 6 tiles back to back '
                                   f'on all {world} rank(s) at once, no kernel (max over ranks; mean and best of 5 runs); '
                                   'scripts/e2e_probe_n.py separates the directions')
            e2e['frac_of_copy_ceiling'] = e2e['ceiling_ms'] / e2e['ms_per_tile']
        except Exception as e:                           # the probe must never cost the bench line
            e2e['ceiling_ms'] = None
            e2e['ceiling_note'] = f'probe failed: {e!r}'

    # ---- CPU baseline (rank 0, N = 1 only) + parity in the same run ----------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_out, cpu_baseline = cpu_baseline_one_core(host_tile, args.cpu_rows)
        pvr = _port_vs_reference()
        cpu_baseline['port_vs_reference_per_core'] = pvr['ratio']
        cpu_baseline['port_vs_reference_source'] = pvr['source']
        if not args.no_e2e:
            rows = cpu_out['WTR'].shape[0]
            keymap = {'WTR': 'WTR_COLLAPSED', 'BWTR': 'BWTR', 'CONF': 'CONF', 'DIAG': 'DIAG'}
            mism = {n: int((host_res[n][:rows] != cpu_out[k]).sum()) for n, k in keymap.items()}
            parity = {'checked_pixels': int(rows * size), 'mismatches': mism,
                      'bit_exact': all(v == 0 for v in mism.values())}

    # ---- the other BASELINE configs as secondary records on the same line ----------------------------------------
    plan.close()
    del plan, tiles
    env.free()
    extras = {}
    if not args.no_extras:
        ex_steps = max(3, min(args.steps, 20))
        for key, fn in (
                ('mosaic', lambda: run_mosaic(env, args, sampler, ex_steps)),
                ('batch64_strong', lambda: run_batch64(env, args, sampler, ex_steps)),
                ('timeseries365', lambda: run_timeseries(env, args, sampler)),
                ('adversarial_worst_case', lambda: dict(
                    run_small_batch(env, args, sampler, ex_steps, name='adversarial', adversarial=True),
                    workload=f'{n_tiles} tiles per GPU of full-range int16 noise in every band (about 40 % of the pixels '
                             'wrap an int16 sum and take the scalar patch path), uniform random Fmask / LAND bytes: '
                             'the data-dependent worst case of the fused kernel')),
                ('swath_edge_tiles', lambda: dict(
                    run_small_batch(env, args, sampler, ex_steps, name='swath_edge', wedge=0.8),
                    workload=f'{n_tiles} tiles per GPU at the edge of a swath: a diagonal no-data wedge over about a third '
                             'of every tile; Mpixel/s '
                             'counts every pixel of the tile, fill included')),
                ('config0_l30', lambda: dict(
                    run_small_batch(env, args, sampler, ex_steps, name='config0', full_product=False,
                                    layers=('DIAG', 'WTR'), bytes_per_px=16),
                    workload=f'configs[0]: {n_tiles} L30-style tiles per GPU without DEM / LAND / ocean, DIAG + WTR only '
                             '(13 B in + 3 B out per pixel)'))):
            try:
                extras[key] = fn()
            except Exception as e:                       # a secondary record must never cost the headline
                extras[key] = {'error': repr(e)}
                env.free()

    clocks = sampler.summary('tiles')
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int16', 'data': 'synthetic',
            'tiles_per_s': value * 1e6 / px_per_tile,
            'config': _config_block(size, n_tiles, world, args.workload),
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': algo_bytes, 'kernel_ms': kernel_ms,
                         'frac_of_nominal_8TBs': achieved / 8000.0,
                         'kernel': kernel_name, 'sustained': sustained},
            'e2e': e2e, 'cpu_baseline': cpu_baseline, 'parity': parity,
            'gpu_launches': args.steps, 'clocks': clocks,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    sampler.stop()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
