"""Row-band sharding of one large plane across ranks (BASELINE config C4).

One process per GPU.  Rank g owns output rows [g*h/N, (g+1)*h/N) of the plane
and keeps them resident together with `halo` rows above and below.  Every
step the halo rows are refreshed from the vertical neighbours with
point-to-point sends (torch.distributed / NCCL over NVLink: the only exchange
the path has -- no reduction, no gather), then the band goes through
morsi_cuda_apply_band_device().  Bands at the image edge get no neighbour
data: rows outside the image are absent (src/morsi.c:30-35).

torch is plumbing here (device memory for the NCCL buffers, the process
group); the kernels are libmorsi_cuda's and run on torch's current stream so
that they are ordered with the NCCL transfers.
"""
import ctypes

from . import binding as B


class BandJob:
    def __init__(self, L, op, e, w, h, rank, world, dist, torch, seed, dist_kind=0):
        self.L, self.op, self.w, self.h = L, op, w, h
        self.rank, self.world, self.dist, self.torch = rank, world, dist, torch
        self.e = e
        self.e_p = e.ctypes.data_as(B._i32p)
        up, down = B.halo_rows(op, e)
        self.up, self.down = up, down
        self.b0 = h * rank // world
        self.b1 = h * (rank + 1) // world
        self.i0 = max(0, self.b0 - up)
        self.i1 = min(h, self.b1 + down)
        rows_in = self.i1 - self.i0
        if torch is not None:
            self.x = torch.empty((rows_in, w), dtype=torch.float32, device="cuda")
            self.y = torch.empty((self.b1 - self.b0, w), dtype=torch.float32, device="cuda")
            self.x_ptr, self.y_ptr = self.x.data_ptr(), self.y.data_ptr()
            self.stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        else:
            self._bx = B.DeviceBuffer(rows_in * w * 4)
            self._by = B.DeviceBuffer((self.b1 - self.b0) * w * 4)
            self.x_ptr, self.y_ptr = self._bx.ptr, self._by.ptr
            self.stream = None
        # own rows only: the halo rows arrive through the exchange
        own = self.x_ptr + (self.b0 - self.i0) * w * 4
        B.check(L.morsi_cuda_synth(own, w, self.b1 - self.b0, self.b0, 0, seed, dist_kind, self.stream))
        B.check(L.morsi_cuda_sync(self.stream))

    def exchange(self):
        """Refresh the halo rows from the neighbours (the path's one exchange)."""
        if self.world == 1:
            return
        t, d = self.torch, self.dist
        ops = []
        o = self.b0 - self.i0                       # halo rows held above the band
        n_own = self.b1 - self.b0
        if self.rank > 0:                           # upper neighbour
            ops.append(d.P2POp(d.isend, self.x[o:o + self.down], self.rank - 1))   # my top rows -> its bottom halo
            ops.append(d.P2POp(d.irecv, self.x[0:o], self.rank - 1))
        if self.rank < self.world - 1:              # lower neighbour
            ops.append(d.P2POp(d.isend, self.x[o + n_own - self.up:o + n_own], self.rank + 1))
            ops.append(d.P2POp(d.irecv, self.x[o + n_own:], self.rank + 1))
        for r in d.batch_isend_irecv(ops):
            r.wait()

    def compute(self):
        B.check(self.L.morsi_cuda_apply_band_device(self.op, self.e_p, self.x_ptr, self.i0, self.i1 - self.i0,
                                                    self.y_ptr, self.b0, self.b1 - self.b0,
                                                    self.w, self.h, self.stream))

    def step(self):
        self.exchange()
        self.compute()
