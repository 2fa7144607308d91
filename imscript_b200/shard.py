"""Row-band sharding of one large plane across ranks (BASELINE config C4): the
Python face of the morsi_shard_* C API (imscript_b200/csrc/shard.cu).

One process per GPU.  Rank g owns output rows [g*h/N, (g+1)*h/N) of the plane
and keeps them resident together with the halo rows of its vertical neighbours.
The DATA PATH is entirely inside libmorsi_cuda: every step the rank's boundary
rows are stored straight into the neighbours' halo rows over NVLink (CUDA IPC
mappings, device-side flags), the interior rows are computed meanwhile and the
edge strips once the halo has landed.  What Python does is plumbing: gather the
128-byte shard handles of all ranks once (torch.distributed all_gather_object,
MPI, a file ... anything) and call the C entry points.

`BandPlan` restates the C bookkeeping (which rows a rank owns, holds, sends and
receives) for the CPU tests (tests/test_shard_gloo.py), where the same exchange
is played over gloo.  Bands at the image edge get no neighbour data: rows
outside the image are absent (src/morsi.c:30-35).  The fused two-stage kernels
recompute the intermediate halo locally, so the halo is stages x reach input
rows (option i).
"""
import ctypes

from . import binding as B


class BandPlan:
    """Pure bookkeeping: which rows a rank owns, holds, sends and receives
    (the same arithmetic as morsi_shard_create / shard_push in shard.cu)."""

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
        """[(kind, peer, first_held_row, n_rows)]: what an exchange moves, in order.
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
    """The halo exchange over torch.distributed point-to-point calls, for the CPU
    (gloo) tests of the bookkeeping; the GPU path does this inside libmorsi_cuda."""
    if plan.world == 1:
        return
    ops = []
    for kind, peer, r0, n in plan.transfers():
        fn = dist.isend if kind == "send" else dist.irecv
        ops.append(dist.P2POp(fn, x[r0:r0 + n], peer))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def gather_handles(handle, dist=None, world=1):
    """All ranks' shard handles, in rank order, as one bytes object (plumbing:
    any transport will do; here torch.distributed's object all-gather)."""
    handle = bytes(handle)
    assert len(handle) == B.SHARD_HANDLE_BYTES
    if dist is None or world == 1:
        return handle
    out = [None] * world
    dist.all_gather_object(out, handle)
    assert all(isinstance(o, bytes) and len(o) == B.SHARD_HANDLE_BYTES for o in out)
    return b"".join(out)


class ShardJob:
    """One rank of a row-band-sharded plane, driven through the C API.
    Buffers 0 and 1 are alternating inputs (a stream of images is uploaded into
    one while the other is processed), buffer 2 receives the result."""

    def __init__(self, L, op, e, w, h, rank, world, device, dist=None, seed=4, dist_kind=0, nbuf=3):
        self.L, self.op, self.w, self.h, self.rank, self.world = L, op, w, h, rank, world
        self.e = e
        self.e_p = e.ctypes.data_as(B._i32p)
        up, down = B.halo_rows(op, e)
        self.plan = BandPlan(h, rank, world, up, down)
        self.s = B._vp()
        B.check(L.morsi_shard_create(ctypes.byref(self.s), device, rank, world, w, h, max(up, down), nbuf))
        hd = ctypes.create_string_buffer(B.SHARD_HANDLE_BYTES)
        B.check(L.morsi_shard_handle(self.s, hd))
        table = gather_handles(hd.raw, dist, world)
        B.check(L.morsi_shard_connect(self.s, table))
        self.stream = B._vp(L.morsi_shard_stream(self.s))
        self.k = 0
        p = self.plan
        self.n_inputs = 2 if nbuf >= 3 else 1
        self.out_buf = nbuf - 1
        # own rows only: the halo rows arrive through the exchange
        for b in range(self.n_inputs):
            B.check(L.morsi_cuda_synth(self.buffer_ptr(b) + p.own_offset * w * 4, w, p.rows_own, p.b0, 0, seed,
                                       dist_kind, self.stream))
        B.check(L.morsi_shard_sync(self.s))
        self.h_x = self.h_y = None

    def buffer_ptr(self, b):
        return self.L.morsi_shard_buffer(self.s, b)

    def out_ptr(self):
        """device pointer of the first OWNED output row"""
        return self.buffer_ptr(self.out_buf) + self.plan.own_offset * self.w * 4

    def step(self):
        B.check(self.L.morsi_shard_apply(self.s, self.op, self.e_p, self.k % self.n_inputs, self.out_buf))
        self.k += 1

    def sync(self):
        B.check(self.L.morsi_shard_sync(self.s))

    def halo_bytes(self):
        return int(self.L.morsi_shard_halo_bytes(self.s))

    # ---- end to end: the band lives in (pinned) HOST memory -----------------
    def host_buffers(self):
        """Pinned host copies of the owned input rows and of the output band."""
        p, L = self.plan, self.L
        self.h_x, self.h_y = B._vp(), B._vp()
        nbytes = p.rows_own * self.w * 4
        B.check(L.morsi_cuda_host_alloc(ctypes.byref(self.h_x), nbytes))
        B.check(L.morsi_cuda_host_alloc(ctypes.byref(self.h_y), nbytes))
        B.check(L.morsi_cuda_memcpy_d2h(self.h_x, self.buffer_ptr(0) + p.own_offset * self.w * 4, nbytes, self.stream))
        self.sync()
        return nbytes

    def e2e_step(self):
        """morsi_shard_apply_host: boundary rows up and pushed first, then the band
        streams through upload / kernels / download; synchronous"""
        B.check(self.L.morsi_shard_apply_host(self.s, self.op, self.e_p, self.h_x, self.h_y))

    def free_host_buffers(self):
        if self.h_x is not None:
            self.L.morsi_cuda_host_free(self.h_x)
            self.L.morsi_cuda_host_free(self.h_y)
            self.h_x = self.h_y = None

    def destroy(self):
        self.free_host_buffers()
        if self.s:
            self.L.morsi_shard_destroy(self.s)
            self.s = None
