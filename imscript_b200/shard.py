"""Row-band sharding of one large plane across ranks (BASELINE config C4).

One process per GPU.  Rank g owns output rows [g*h/N, (g+1)*h/N) of the plane
and keeps them resident together with `up`/`down` halo rows.  Every step the
halo rows are refreshed from the vertical neighbours with point-to-point
sends (torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU
tests) -- the only exchange the path has, no reduction, no gather -- and the
band then goes through morsi_cuda_apply_band_device().  Bands at the image
edge get no neighbour data: rows outside the image are absent
(src/morsi.c:30-35).  The fused two-stage kernels recompute the intermediate
halo locally, so the halo is stages x reach input rows (option i).

torch is plumbing here (device memory for the NCCL buffers, the process
group); the kernels are libmorsi_cuda's and run on torch's current stream so
that they are ordered with the transfers.
"""
import ctypes

from . import binding as B


class BandPlan:
    """Pure bookkeeping: which rows a rank owns, holds, sends and receives."""

    def __init__(self, h, rank, world, up, down):
        self.h, self.rank, self.world, self.up, self.down = h, rank, world, up, down
        self.b0 = h * rank // world              # first owned row
        self.b1 = h * (rank + 1) // world        # one past the last owned row
        self.i0 = max(0, self.b0 - up)           # first held row
        self.i1 = min(h, self.b1 + down)         # one past the last held row
        self.rows_held = self.i1 - self.i0
        self.rows_own = self.b1 - self.b0
        self.own_offset = self.b0 - self.i0      # held-row index of the first owned row

    def transfers(self):
        """[(kind, peer, first_held_row, n_rows)]: what exchange() posts, in order.
        The band of a neighbour may be shorter than the halo (many ranks, small
        images): only the rows the neighbour actually owns are exchanged here."""
        t = []
        o, n = self.own_offset, self.rows_own
        if self.rank > 0:
            peer_rows = self.b0 - self.h * (self.rank - 1) // self.world
            t.append(("send", self.rank - 1, o, min(self.down, n)))                 # my top rows -> its bottom halo
            t.append(("recv", self.rank - 1, o - min(self.up, peer_rows), min(self.up, peer_rows)))
        if self.rank < self.world - 1:
            peer_rows = self.h * (self.rank + 2) // self.world - self.b1
            t.append(("send", self.rank + 1, o + n - min(self.up, n), min(self.up, n)))
            t.append(("recv", self.rank + 1, o + n, min(self.down, peer_rows)))
        return [x for x in t if x[3] > 0]

    def halo_complete(self):
        """True when one exchange with the direct neighbours fills the whole halo."""
        ok = True
        if self.rank > 0:
            ok &= self.b0 - self.h * (self.rank - 1) // self.world >= min(self.up, self.b0)
        if self.rank < self.world - 1:
            ok &= self.h * (self.rank + 2) // self.world - self.b1 >= min(self.down, self.h - self.b1)
        return ok


def exchange(x, plan, dist):
    """Refresh the halo rows of the held band `x` (a (rows_held, w) tensor)."""
    if plan.world == 1:
        return
    ops = []
    for kind, peer, r0, n in plan.transfers():
        fn = dist.isend if kind == "send" else dist.irecv
        ops.append(dist.P2POp(fn, x[r0:r0 + n], peer))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class BandJob:
    def __init__(self, L, op, e, w, h, rank, world, dist, torch, seed, dist_kind=0):
        self.L, self.op, self.w, self.h = L, op, w, h
        self.dist, self.torch = dist, torch
        self.e = e
        self.e_p = e.ctypes.data_as(B._i32p)
        up, down = B.halo_rows(op, e)
        self.plan = p = BandPlan(h, rank, world, up, down)
        if not p.halo_complete():
            raise B.MorsiError(1, f"bands of {h // world} rows are shorter than the {up}-row halo")
        if torch is not None:
            self.x = torch.empty((p.rows_held, w), dtype=torch.float32, device="cuda")
            self.y = torch.empty((p.rows_own, w), dtype=torch.float32, device="cuda")
            self.x_ptr, self.y_ptr = self.x.data_ptr(), self.y.data_ptr()
            self.stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        else:
            self._bx = B.DeviceBuffer(p.rows_held * w * 4)
            self._by = B.DeviceBuffer(p.rows_own * w * 4)
            self.x_ptr, self.y_ptr = self._bx.ptr, self._by.ptr
            self.stream = None
        # own rows only: the halo rows arrive through the exchange
        own = self.x_ptr + p.own_offset * w * 4
        B.check(L.morsi_cuda_synth(own, w, p.rows_own, p.b0, 0, seed, dist_kind, self.stream))
        B.check(L.morsi_cuda_sync(self.stream))

    def exchange(self):
        if self.plan.world > 1:
            exchange(self.x, self.plan, self.dist)

    def compute(self):
        p = self.plan
        B.check(self.L.morsi_cuda_apply_band_device(self.op, self.e_p, self.x_ptr, p.i0, p.rows_held,
                                                    self.y_ptr, p.b0, p.rows_own, self.w, self.h, self.stream))

    def step(self):
        self.exchange()
        self.compute()

    # ---- end to end: the band lives in (pinned) HOST memory -----------------
    def host_buffers(self):
        """Pinned host copies of the owned input rows and of the output band."""
        p, L = self.plan, self.L
        self.h_x, self.h_y = B._vp(), B._vp()
        nbytes = p.rows_own * self.w * 4
        B.check(L.morsi_cuda_host_alloc(ctypes.byref(self.h_x), nbytes))
        B.check(L.morsi_cuda_host_alloc(ctypes.byref(self.h_y), nbytes))
        own = self.x_ptr + p.own_offset * self.w * 4
        B.check(L.morsi_cuda_memcpy_d2h(self.h_x, own, nbytes, self.stream))
        B.check(L.morsi_cuda_sync(self.stream))
        return nbytes

    def e2e_step(self):
        """host -> device copy of the owned rows, halo exchange between the
        ranks, the kernels, device -> host copy of the result, all on one stream"""
        p, L = self.plan, self.L
        nbytes = p.rows_own * self.w * 4
        own = self.x_ptr + p.own_offset * self.w * 4
        B.check(L.morsi_cuda_memcpy_h2d(own, self.h_x, nbytes, self.stream))
        self.exchange()
        self.compute()
        B.check(L.morsi_cuda_memcpy_d2h(self.h_y, self.y_ptr, nbytes, self.stream))
        B.check(L.morsi_cuda_sync(self.stream))

    def free_host_buffers(self):
        self.L.morsi_cuda_host_free(self.h_x)
        self.L.morsi_cuda_host_free(self.h_y)
