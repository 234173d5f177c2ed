"""Multi-GPU partitioning of the classification path.

Two schemes:

* **by tile** - MGRS tiles / acquisitions are independent: tile ``i`` goes to
  rank ``i % world`` and nothing is exchanged (``shard_tiles``).
* **row strips of one oversized raster** - every function of the path is
  point-wise except the terrain-shadow stencil (``np.gradient`` inside
  ``_compute_opera_shadow_layer``, dswx_hls.py:4255) whose radius is one DEM
  row.  Each rank owns a contiguous strip of rows and needs ONE DEM row from
  each neighbour: ``exchange_dem_halo`` posts the two sends / receives as one
  ``torch.distributed`` P2P batch (NCCL send/recv over NVLink on GPUs; gloo in
  the CPU tests).  ``MosaicStrip`` classifies the interior rows while the
  exchange is in flight and the two boundary rows afterwards.

The reference never splits a raster; parity is defined against the reference
run on the whole raster.
"""
from __future__ import annotations

import numpy as np

from .params import DEM_MARGIN_IN_PIXELS


EDGE_ROWS = 4        # rows at either end of a strip that wait for the halo exchange (overlap mode)


def shard_tiles(n_tiles, rank, world):
    """Indices of the tiles rank ``rank`` processes (round robin)."""
    return list(range(int(rank), int(n_tiles), int(world)))


def strip_bounds(height, world, align=32):
    """Row ranges [r0, r1) of ``world`` contiguous strips.  Strip starts are
    multiples of ``align`` rows (keeps every plane's strip pointer 16-byte
    aligned for any width that is a multiple of 4)."""
    height, world = int(height), int(world)
    blocks = -(-height // align)
    per, extra = divmod(blocks, world)
    bounds, r = [], 0
    for k in range(world):
        nb = per + (1 if k < extra else 0)
        r1 = min(height, r + nb * align)
        bounds.append((r, r1))
        r = r1
    return bounds


def dem_rows_for_strip(r0, r1, height, margin=DEM_MARGIN_IN_PIXELS):
    """DEM rows (in the margin-included DEM array) a rank holds locally: its
    own pixel rows, plus the whole top margin on the first strip and the whole
    bottom margin on the last - i.e. a partition of the DEM rows, no overlap."""
    d0 = 0 if r0 == 0 else margin + r0
    d1 = height + 2 * margin if r1 == height else margin + r1
    return d0, d1


def exchange_dem_halo(dem_ext, n_rows, rank, world, group=None):
    """Fill ``dem_ext[0]`` and ``dem_ext[n_rows + 1]`` with the neighbouring
    ranks' boundary DEM rows.

    ``dem_ext`` is a (n_rows + 2, pitch) float32 tensor whose rows
    ``1 .. n_rows`` already hold this rank's DEM rows (pixel rows of the
    strip).  Rank 0 keeps its own ``dem_ext[0]`` (taken from the DEM margin by
    the caller) and the last rank its own ``dem_ext[n_rows + 1]``.
    Returns the list of outstanding P2P requests (empty when world == 1)."""
    import torch.distributed as dist
    if world == 1:
        return []
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, dem_ext[1], rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, dem_ext[0], rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, dem_ext[n_rows], rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, dem_ext[n_rows + 1], rank + 1, group))
    return dist.batch_isend_irecv(ops) if ops else []


def init_library_comm(ctx, rank, world, group=None):
    """Give ``ctx`` (a ``proteus_b200.Context``) the NCCL communicator of the C ABI (``pb200_comm_init``): rank 0
    draws the unique id, ``torch.distributed`` (any backend - here it is only the host channel for 128 bytes) hands it
    round, every rank joins.  Collective over ``group``."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib
    lib = ctx._lib
    idb = (C.c_uint8 * _lib.COMM_ID_BYTES)()
    if rank == 0:
        _lib.check(lib.pb200_comm_unique_id(idb))
    if world > 1:
        box = [bytes(idb) if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idb = (C.c_uint8 * _lib.COMM_ID_BYTES)(*box[0])
    _lib.check(lib.pb200_comm_init(ctx.handle, idb, int(rank), int(world)))
    return ctx


class MosaicStrip:
    """One rank's row strip of an oversized raster, device resident.

    ``exchange``: 'torch' - ``torch.distributed`` P2P batch (NCCL on GPUs, gloo in the CPU tests); 'library' - the C
    ABI's own exchange (``pb200_halo_exchange_dem`` on the communicator of ``init_library_comm``), what a non-Python
    host uses.  ``overlap``: True - exchange on a side stream while the interior rows are classified, the two
    boundary bands (4 rows each) afterwards (two launches); False - exchange first, then ONE launch over the whole strip.

    Parameters
    ----------
    bands, fmask, land, ocean : torch CUDA tensors of the strip, [n_rows, W]
    dem_local : float32 CUDA tensor [d1 - d0, W + 2*margin] with the DEM rows
        of ``dem_rows_for_strip`` (margin columns included)
    r0, r1, height : strip rows and full raster height
    """

    def __init__(self, bands, fmask, dem_local, land, ocean, r0, r1, height, *,
                 sun_azimuth, sun_elevation, params=None, outputs=None,
                 margin=DEM_MARGIN_IN_PIXELS, rank=0, world=1, group=None, ctx=None,
                 exchange='torch', overlap=True):
        import torch
        from .engine import GRADED_LAYERS, Plan, get_context
        from .params import make_params
        self.rank, self.world, self.group = int(rank), int(world), group
        if exchange not in ('torch', 'library'):
            raise ValueError("exchange: 'torch' or 'library'")
        self.exchange, self.overlap = exchange, bool(overlap)
        self.r0, self.r1, self.height = int(r0), int(r1), int(height)
        n = self.r1 - self.r0
        self.n_rows = n
        w = int(fmask.shape[1])
        pitch = int(dem_local.shape[1])
        assert pitch == w + 2 * margin, 'dem_local must carry the column margins'
        d0, d1 = dem_rows_for_strip(self.r0, self.r1, self.height, margin)
        assert int(dem_local.shape[0]) == d1 - d0, (tuple(dem_local.shape), d0, d1)
        # extended DEM: row 0 and row n+1 are the halo rows
        self.dem_ext = torch.empty((n + 2, pitch), dtype=torch.float32, device=fmask.device)
        first = margin + self.r0 - d0              # local index of the strip's first pixel row
        self.dem_ext[1:n + 1].copy_(dem_local[first:first + n])
        if self.r0 == 0:
            self.dem_ext[0].copy_(dem_local[first - 1])            # from the DEM margin
        if self.r1 == self.height:
            self.dem_ext[n + 1].copy_(dem_local[first + n])
        self.params = params if params is not None else make_params()
        self.layers = tuple(outputs or GRADED_LAYERS)
        ctx = ctx or get_context()
        self.ctx = ctx
        dev = fmask.device
        self.outputs = {name: torch.empty((n, w), device=dev,
                                          dtype=torch.int16 if name == 'DIAG' else torch.uint8)
                        for name in self.layers}
        self.counters = torch.zeros((1, 12), dtype=torch.int64, device=dev)

        def sub(rows):
            a, b = rows
            t = dict(bands=[x[a:b] for x in bands], fmask=fmask[a:b],
                     land=land[a:b] if land is not None else None,
                     ocean=ocean[a:b] if ocean is not None else None,
                     dem=self.dem_ext, dem_off=(1 + a, margin),
                     sun_azimuth=sun_azimuth, sun_elevation=sun_elevation)
            outs = {k: v[a:b] for k, v in self.outputs.items()}
            return t, outs

        self._interior = self._edges = self._whole = None
        if self.overlap:
            # the rows whose stencil reaches a halo row are the first and the last one; they are classified as 4-row bands
            # so that every piece keeps a height that is a multiple of 4 and 16-byte aligned planes (the TMA-fed kernel)
            e = EDGE_ROWS if n >= 3 * EDGE_ROWS else 1
            edge_rows = [(0, min(e, n))] + ([(max(e, n - e), n)] if n > e else [])
            if n > 2 * e:
                t, o = sub((e, n - e))
                self._interior = Plan([t], self.params, self.layers, ctx=ctx, outputs_into=[o],
                                      counters_into=self.counters)
            tiles, outs = zip(*[sub(r) for r in edge_rows])
            self._edges = Plan(list(tiles), self.params, self.layers, ctx=ctx, outputs_into=list(outs),
                               counters_into=self.counters.expand(len(tiles), 12))
        else:
            t, o = sub((0, n))
            self._whole = Plan([t], self.params, self.layers, ctx=ctx, outputs_into=[o], counters_into=self.counters)
        self._side = torch.cuda.Stream(device=dev) if fmask.is_cuda else None

    def _exchange(self, halo, stream):
        """Fill the two halo rows of ``dem_ext`` on ``stream`` (the current stream of the caller's context)."""
        import ctypes as C
        from . import _lib
        if halo is not None:
            if halo[0] is not None:
                self.dem_ext[0].copy_(halo[0])
            if halo[1] is not None:
                self.dem_ext[self.n_rows + 1].copy_(halo[1])
        elif self.exchange == 'library':
            _lib.check(self.ctx._lib.pb200_halo_exchange_dem(
                self.ctx.handle, self.dem_ext.data_ptr(), self.n_rows, int(self.dem_ext.shape[1]),
                C.c_void_p(stream.cuda_stream)))
        else:
            for r in exchange_dem_halo(self.dem_ext, self.n_rows, self.rank, self.world, self.group):
                r.wait()

    def run(self, halo=None):
        """One classification of the strip, asynchronous on the current stream: the halo exchange (overlapped with
        the interior rows when ``overlap``), then the rows that need it.

        ``halo`` = (row_above, row_below) tensors (either may be None) replaces
        the exchange - used to emulate several ranks inside one process."""
        import torch
        cur = torch.cuda.current_stream()
        if not self.overlap:
            self._exchange(halo, cur)
            self._whole.run(cur)
            return
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            self._exchange(halo, self._side)
        if self._interior is not None:
            self._interior.run(cur)
        cur.wait_stream(self._side)
        self._edges.run(cur)

    def allreduce_counters(self):
        """Raster-wide coverage counters on every rank (D:5104-5111 over the whole raster): one 3 x uint64 sum."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        if self.world == 1:
            return self.counters
        if self.exchange == 'library':
            _lib.check(self.ctx._lib.pb200_comm_allreduce_u64(
                self.ctx.handle, self.counters.data_ptr(), 3, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:
            dist.all_reduce(self.counters, op=dist.ReduceOp.SUM, group=self.group)
        return self.counters

    def zero_counters(self):
        self.counters.zero_()

    def results(self):
        import torch
        torch.cuda.synchronize()
        res = {k: (v.cpu().numpy().view(np.uint16) if k == 'DIAG' else v.cpu().numpy())
               for k, v in self.outputs.items()}
        res['counters'] = self.counters[0].cpu().numpy().astype(np.uint64)
        return res


def otsu_threshold_of_strips(local_counts, is_normalized=True, group=None):
    """Otsu threshold (dswx_hls.py:1638-1686) of a raster held as row strips by the ranks of ``group``: the 256
    exact per-value counts of every strip (``dswx_hls._otsu_counts`` on the GPU, int64 tensor) are summed with ONE
    all-reduce - the only exchange the 'otsu' shadow algorithm needs - and every rank derives
    the same threshold from the sum."""
    import torch
    import torch.distributed as dist
    from .dswx_hls import otsu_threshold_from_counts
    counts = local_counts if isinstance(local_counts, torch.Tensor) else torch.as_tensor(np.asarray(local_counts))
    counts = counts.to(torch.int64).clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return otsu_threshold_from_counts(counts.cpu().numpy(), is_normalized)
