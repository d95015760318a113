"""Row-sharded iterated SpMV (power iteration) over the GPUs of one box: one process per GPU.

New with respect to the reference, which is single-device; specified by
BASELINE.json: rows are split by nnz-balanced merge-path boundaries (the same device search
that cuts tiles), every rank keeps its rows of A and a full replica of x, and after each
SpMV every rank needs everybody's slice of the new x.

Three exchanges ("auto" tries them in this order):

  "mc"    x replicas live in torch symmetric memory bound to one NVLink multicast object
          (NVLS): the SpMV kernel stores each y value once with multimem.st and the NVSwitch
          replicates it into every GPU's replica, so a rank's egress is its own slice, not
          (P-1) copies of it -- which matters because nnz-balanced row blocks have very unequal
          row counts (on R-MAT scale 27 one of 8 ranks owns 40 % of the rows).
  "p2p"   the SpMV kernel itself stores each y value into the local replica AND into the
          peer-mapped replicas of the other GPUs (cudaIpc handles over NVLink / NVSwitch), so
          the all-gather is fused into the kernel's epilogue and overlaps its own gathers; the
          only collective left is the 8-byte all-reduce of the squared norm, which doubles as
          the step barrier.
  "nccl"  the kernel stores locally, then an NCCL all-gather of the (uneven) slices.

x is double-buffered: step t reads buffer t % 2 and writes buffer (t + 1) % 2; a rank cannot
start step t + 1 before the norm all-reduce of step t, which every rank enters after its own
step-t kernel has finished reading and writing, so one collective per step orders everything.

The power iteration is x <- A x / ||A x||; the 1/||.|| is applied by the next SpMV through its
device-side alpha, so the host never synchronises inside the loop.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import torch

from . import _lib, generate, spmv as spmv_mod

MAILBOX_BYTES = 256   # SPMVB200_MAILBOX_BYTES in include/spmv_b200.h


class _RawDeviceArray:
    """A cudaMalloc'ed buffer exposed to torch without copying (__cuda_array_interface__)."""

    def __init__(self, n: int, dtype: torch.dtype):
        self.n = n
        self.dtype = dtype
        self.itemsize = torch.empty(0, dtype=dtype).element_size()
        p = C.c_void_p()
        _lib.check(_lib.lib().spmvb200_device_malloc(max(n, 1) * self.itemsize, C.byref(p)),
                   "spmvb200_device_malloc")
        self.ptr = p.value
        typestr = {torch.float32: "<f4", torch.float64: "<f8"}[dtype]
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self) -> torch.Tensor:
        return torch.as_tensor(self, device=f"cuda:{torch.cuda.current_device()}")

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(_lib.lib().spmvb200_ipc_export(C.c_void_p(self.ptr), buf), "spmvb200_ipc_export")
        return buf.raw

    def free(self):
        if self.ptr:
            _lib.lib().spmvb200_device_free(C.c_void_p(self.ptr))
            self.ptr = 0


def _ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    _lib.check(_lib.lib().spmvb200_ipc_open(handle, C.byref(p)), "spmvb200_ipc_open")
    return p.value


@dataclass
class Shard:
    rank: int
    world: int
    row_begin: int
    row_end: int
    nnz_begin: int
    nnz_end: int
    row_bounds: list
    csr: generate.Csr  # local rows, offsets rebased to 0, n_cols = global


def shard_rows(global_csr: generate.Csr, rank: int, world: int, value_seed=None, row_bounds=None,
               weight=(1, 1)) -> Shard:
    """Cut rank's rows out of a global CSR held on this device.  Boundaries come from the
    device merge-path search (bit-exact against oracle.cpu.row_split in the tests), with a row
    weighing `weight` = (num, den) nonzeros, or are given (row_bounds, world + 1 values)."""
    if row_bounds is not None:
        rb = [int(v) for v in row_bounds]
        if len(rb) != world + 1 or rb[0] != 0 or rb[-1] != global_csr.n_rows or \
                any(a > b for a, b in zip(rb, rb[1:])):
            raise ValueError("row_bounds must be world + 1 non-decreasing rows from 0 to n_rows")
    elif world > 1:
        rb = spmv_mod.row_split(global_csr.Ap, world, nnz=global_csr.nnz, weight=weight)
    else:
        rb = [0, global_csr.n_rows]
    r0, r1 = rb[rank], rb[rank + 1]
    k0, k1 = int(global_csr.Ap[r0].item()), int(global_csr.Ap[r1].item())
    if world == 1:
        local = global_csr
    else:
        Ap = (global_csr.Ap[r0:r1 + 1] - k0).contiguous()
        Aj = global_csr.Aj[k0:k1].clone()
        Ax = global_csr.Ax[k0:k1].clone()
        local = generate.Csr(r1 - r0, global_csr.n_cols, k1 - k0, Ap, Aj, Ax,
                             f"{global_csr.name}[rank {rank}/{world}]")
    return Shard(rank, world, r0, r1, k0, k1, rb, local)


def rebalanced_bounds(global_csr: generate.Csr, row_bounds, times, weight=(1, 1)):
    """Row boundaries for a re-split from measured per-shard times: the cost f(r) = den*Ap[r] + num*r
    of every current boundary, the equal-time quantiles of the piecewise-linear time-over-cost
    curve (spmv.rebalance_targets), and the device search for the rows at those costs."""
    wn, wd = int(weight[0]), int(weight[1])
    idx = torch.tensor([int(b) for b in row_bounds], dtype=torch.int64, device=global_csr.Ap.device)
    ap = [int(v) for v in global_csr.Ap[idx].tolist()]
    cost = [wd * a + wn * int(b) for a, b in zip(ap, row_bounds)]
    targets = spmv_mod.rebalance_targets(cost, times)
    rows = spmv_mod.rows_at_cost(global_csr.Ap, targets, (wn, wd))
    return [0] + rows + [global_csr.n_rows]


class PowerIteration:
    """x <- A x / ||A x||, A row-sharded over `world` ranks (world = 1: the whole matrix)."""

    def __init__(self, shard: Shard, n_rows_global: int, kind: str = "auto", exchange: str = "auto",
                 group=None):
        """Every step goes through libspmvb200 and needs CUDA.  The four device operations of a
        step are methods (_setup_buffers, _local_spmv, _local_sumsq, _alpha_from_sumsq) so that
        tests/test_dist_cpu.py can drive the exchange protocol -- double buffering, uneven
        all-gather, norm all-reduce -- over gloo with a subclass of its own that replaces them by
        the CPU oracle; this module contains no host arithmetic."""
        import torch.distributed as dist
        self.dist = dist
        self.shard = shard
        self.world, self.rank = shard.world, shard.rank
        self.kind = kind
        self.n = n_rows_global
        self.group = group
        m = shard.csr
        self.dtype = m.Ax.dtype
        self.vbits = m.Ax.element_size() * 8
        assert m.n_cols == n_rows_global, "power iteration needs a square matrix"
        self.exchange = exchange if shard.world > 1 else "none"
        self.exchange_note = ""
        self._raw, self._symm = [], []
        self.peer_ptrs = [[], []]   # per buffer: peers' base addresses mapped into this process
        self.mc_ptrs = [0, 0]       # per buffer: NVLink multicast address (exchange "mc")
        # Every replica of x is allocated with a small tail: buffer 0's holds this rank's mailbox
        # of the norm exchange (csrc/power.cu), so whatever maps the replicas into the peers --
        # cudaIpc, symmetric memory, multicast -- maps the mailboxes too.
        item = torch.empty(0, dtype=self.dtype).element_size()
        self._tail_off = (self.n * item + 255) // 256 * 256
        self._alloc_elems = (self._tail_off + MAILBOX_BYTES + item - 1) // item
        self._xchg_step = 0         # never reset: the mailbox slots are addressed by it
        self._calls_on_shard = 0    # SpMVs issued on the current shard (reset() keeps the matrix)
        dev = self._setup_buffers()
        self._local_events = None   # set by shard_local_ms: (start, stop) events per step
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.alpha = torch.ones(1, dtype=self.dtype, device=dev)
        self.step_no = 0
        # the norm exchange kernel replaces sum of squares + all-reduce + 1/sqrt wherever the
        # mailboxes are mapped into the peers (always on one GPU); "nccl" keeps the collective
        self.fused_norm = dev == "cuda" and self.exchange in ("none", "p2p", "mc")
        if self.fused_norm:
            self._init_mailbox()
        self.reset()

    # ------------------------------------------------------------------ setup
    def _setup_buffers(self) -> str:
        """Allocate the two replicas of x and map them into the peers; returns the device."""
        shard = self.shard
        if not torch.cuda.is_available():
            raise RuntimeError("PowerIteration needs a CUDA device; there is no CPU path")
        dev = "cuda"
        # "auto" takes the multicast exchange at every world size and falls back to peer stores where
        # there is no NVLS.  Measured on R-MAT scale 27 with the multicast-specialised tile kernel:
        # 5.68 / 2.93 / 1.62 ms per step at 2 / 4 / 8 GPUs against 5.87 / 3.30 / 2.6 with peer stores
        # (rows are split by nonzeros, so one rank owns most of the ROWS and sends them world-1 times).
        if self.exchange in ("mc", "auto"):
            try:
                self._alloc_symmetric()
                self.exchange = "mc"
            except Exception as e:  # no symmetric memory / no NVLS here
                self.exchange_note = f"multicast unavailable ({type(e).__name__}: {e}); "
                self.exchange = "p2p"
                self._symm, self.xbuf, self._full = [], [], []
        if self.exchange != "mc":
            self._raw = [_RawDeviceArray(self._alloc_elems, self.dtype) for _ in range(2)]
            self._full = [r.tensor() for r in self._raw]
            self.xbuf = [t[:self.n] for t in self._full]
        if self.exchange == "p2p":
            try:
                self._map_peers()
            except Exception as e:  # no IPC / no peer access: fall back to NCCL, and say so
                self.exchange = "nccl"
                self.exchange_note += f"p2p unavailable ({type(e).__name__}: {e}); nccl all-gather"
        return dev

    def _sync(self):
        torch.cuda.synchronize()

    def _local_spmv(self, x, y, peers):
        """y = alpha * A_local x, fanned out to the peers' replicas."""
        m = self.shard.csr
        # the shard never changes between steps: from its second SpMV on, what earlier calls
        # derived from it (tile coordinates, the hot-x plan) is reused
        spmv_mod.spmv_ex(self.kind, m.Ap, m.Aj, m.Ax, x, y, n_cols=self.n, alpha_dev=self.alpha,
                         y_peers=peers, multicast=self.exchange == "mc",
                         static_pattern=self._calls_on_shard > 0)
        self._calls_on_shard += 1

    def _local_sumsq(self, y):
        st = _lib.lib().spmvb200_sum_squares(self.vbits, y.numel(), y.data_ptr(), self.sumsq.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "spmvb200_sum_squares")

    def _alpha_from_sumsq(self):
        st = _lib.lib().spmvb200_inv_sqrt(self.vbits, self.sumsq.data_ptr(), self.alpha.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "spmvb200_inv_sqrt")

    def _alloc_symmetric(self):
        """x replicas in torch symmetric memory: every rank allocates the same buffers, the
        rendezvous maps them into one NVLink multicast object (NVLS), and a single
        multimem.st from the SpMV kernel lands in all replicas."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else self.dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        self._full = [symm_mem.empty(self._alloc_elems, dtype=self.dtype, device=dev) for _ in range(2)]
        self.xbuf = [t[:self.n] for t in self._full]
        for t in self._full:
            try:
                h = symm_mem.rendezvous(t, group.group_name)
            except TypeError:
                h = symm_mem.rendezvous(t, group=group)
            self._symm.append(h)
        self.mc_ptrs = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self._symm]
        ok = torch.tensor([int(all(self.mc_ptrs))], device="cuda")
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            raise RuntimeError("no multicast pointer")

    def _init_mailbox(self):
        """Mailbox of the norm exchange in the tail of replica 0: all slots empty (-1.0), and the
        addresses under which every rank's mailbox is reached from here."""
        base = self._full[0].data_ptr()
        self._mailbox = base + self._tail_off
        tail = self._full[0].view(torch.uint8)[self._tail_off:self._tail_off + MAILBOX_BYTES]
        tail.view(torch.float64).fill_(-1.0)
        self.xchg_error = torch.zeros(1, dtype=torch.int32, device="cuda")
        self._mailbox_mc = 0
        ptrs = [0] * self.world
        ptrs[self.rank] = self._mailbox
        if self.exchange == "mc":
            self._mailbox_mc = self.mc_ptrs[0] + self._tail_off
        elif self.exchange == "p2p":
            others = [q for q in range(self.world) if q != self.rank]
            for q, p in zip(others, self.peer_ptrs[0]):
                ptrs[q] = p + self._tail_off
        self._mailbox_of_rank = (C.c_void_p * self.world)(*ptrs)
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.group)   # every mailbox is empty before anybody publishes

    def _map_peers(self):
        handles = [r.ipc_handle() for r in self._raw]
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, handles, group=self.group)
        for q in range(self.world):
            if q == self.rank:
                continue
            for b in range(2):
                self.peer_ptrs[b].append(_ipc_open(gathered[q][b]))
        # every rank must have every mapping before anybody stores through one
        self.dist.barrier(group=self.group)

    def reset(self):
        """x0 = 1/sqrt(n), alpha = 1."""
        self.xbuf[0].fill_(1.0 / (self.n ** 0.5))
        self.xbuf[1].zero_()
        self.alpha.fill_(1.0)
        self.step_no = 0
        self._sync()
        if self.world > 1:
            self.dist.barrier(group=self.group)

    # ------------------------------------------------------------------ one step
    def step(self):
        s, m = self.shard, self.shard.csr
        cur, nxt = self.step_no & 1, (self.step_no + 1) & 1
        x = self.xbuf[cur]
        y = self.xbuf[nxt][s.row_begin:s.row_end]
        item = self.xbuf[nxt].element_size()
        if self.exchange == "mc":
            peers = [self.mc_ptrs[nxt] + s.row_begin * item]
        elif self.exchange == "p2p":
            peers = [p + s.row_begin * item for p in self.peer_ptrs[nxt]]
        else:
            peers = []
        if self._local_events is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._local_events.append(ev)
            ev[0].record()
        self._local_spmv(x, y, peers)
        if self.step_no == 0 and self.exchange in ("p2p", "mc"):
            # The kernels send only rows that have nonzeros to the peers; an empty row's
            # entry must therefore already be 0 in every replica.  Buffer 1 starts zeroed;
            # buffer 0 held x0 and is cleared here, after this step's kernel has read it
            # and before the step barrier lets any peer store into it.
            self.xbuf[0].zero_()
        if self.fused_norm and self._local_events is None:
            # sum of squares, exchange of the per-rank sums, alpha and the step barrier: one kernel
            st = _lib.lib().spmvb200_norm_exchange(
                self.vbits, y.numel(), y.data_ptr(), self.rank, self.world, self._xchg_step,
                self._mailbox, self._mailbox_of_rank, self._mailbox_mc or None, self.sumsq.data_ptr(),
                self.alpha.data_ptr(), self.xchg_error.data_ptr(), torch.cuda.current_stream().cuda_stream)
            _lib.check(st, "spmvb200_norm_exchange")
            self._xchg_step += 1
            self.step_no += 1
            return
        self._local_sumsq(y)
        if self._local_events is not None:
            self._local_events[-1][1].record()
        if self.world > 1:
            if self.exchange == "nccl":
                views = [self.xbuf[nxt][s.row_bounds[q]:s.row_bounds[q + 1]] for q in range(self.world)]
                even = len({v.numel() for v in views}) == 1
                if even or self.dist.get_backend(self.group) == "nccl":
                    self.dist.all_gather(views, y, group=self.group)   # NCCL takes uneven slices
                else:
                    for q in range(self.world):                        # gloo: one broadcast per slice
                        self.dist.broadcast(views[q], src=q, group=self.group)
            # the norm all-reduce is also the step barrier for the peer stores
            self.dist.all_reduce(self.sumsq, group=self.group)
        self._alpha_from_sumsq()
        self.step_no += 1

    # ------------------------------------------------------------------ re-balancing
    def set_shard(self, shard: Shard):
        """Continue with another row split of the same matrix (same ranks, same x replicas);
        the iteration restarts from x0."""
        if shard.world != self.world or shard.rank != self.rank or shard.csr.n_cols != self.n:
            raise ValueError("set_shard: the new shard must belong to the same matrix and rank")
        self.shard = shard
        self._calls_on_shard = 0
        self.reset()

    def shard_local_ms(self, steps: int = 5):
        """Mean duration of this rank's local part of a step -- everything between the step
        barriers that does not wait for a peer: partition, tile kernel, carry fix-up, sum of
        squares -- over `steps` steps (after two warm-up steps), for every rank (the same list on
        all ranks).  CUDA events on the stream the step runs on."""
        # two untimed steps first: the first SpMV on a new shard searches the partition and the
        # second builds the hot-x / table plan (tens of ms on a large shard), neither of which a
        # later step pays
        for _ in range(2):
            self.step()
        torch.cuda.synchronize()
        self._local_events = []
        try:
            for _ in range(steps):
                self.step()
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in self._local_events) / max(len(self._local_events), 1)
        finally:
            self._local_events = None
        t = torch.zeros(self.world, dtype=torch.float64, device="cuda")
        t[self.rank] = ms
        if self.world > 1:
            self.dist.all_reduce(t, group=self.group)
        return [float(v) for v in t.tolist()]

    def rebalance(self, global_csr: generate.Csr, steps: int = 5, weight=(1, 1)):
        """Measure every rank's local step time on the current split, move the row boundaries to
        the equal-time quantiles, re-shard and restart.  Returns (times, new row bounds)."""
        times = self.shard_local_ms(steps)
        rb = rebalanced_bounds(global_csr, self.shard.row_bounds, times, weight)
        self.set_shard(shard_rows(global_csr, self.rank, self.world, row_bounds=rb))
        return times, rb

    def current_x(self) -> torch.Tensor:
        """The latest iterate, not yet scaled by alpha (= 1 / its norm)."""
        return self.xbuf[self.step_no & 1]

    def check_exchange(self):
        """Raise if a rank failed to arrive at a norm exchange (the kernel gives up after ~4 s)."""
        if getattr(self, "fused_norm", False):
            e = int(self.xchg_error.item())
            if e:
                raise RuntimeError(f"norm exchange: rank {e - 1} did not arrive within the time limit")

    def eigen_estimate(self) -> float:
        """||A x_k|| with ||x_k|| = 1: converges to |lambda_max|."""
        self.check_exchange()
        return float(self.sumsq.item()) ** 0.5

    def close(self):
        self._sync()
        self.check_exchange()
        if self.world > 1:
            self.dist.barrier(group=self.group)
        for b in range(2):
            for p in self.peer_ptrs[b]:
                _lib.lib().spmvb200_ipc_close(C.c_void_p(p))
        self.peer_ptrs = [[], []]
        self.xbuf = []
        self._full = []
        self._symm = []
        for r in self._raw:
            r.free()


def even_bounds(n: int, world: int):
    """Host-I/O slices: row q*n//world .. (q+1)*n//world belongs to rank q's host buffers."""
    return [q * n // world for q in range(world + 1)]


def overlap_sizes(src_bounds, src_rank, dst_bounds):
    """How many rows of src_bounds' block `src_rank` fall into each block of dst_bounds (both
    partitions of [0, n) into contiguous ascending blocks): the send split of an all-to-all that
    moves a vector from one row partition to another."""
    a0, a1 = src_bounds[src_rank], src_bounds[src_rank + 1]
    return [max(0, min(a1, dst_bounds[q + 1]) - max(a0, dst_bounds[q])) for q in range(len(dst_bounds) - 1)]


class ShardedHostSpMV:
    """y = A x with x and y in HOST memory, A (square) row-sharded over the ranks.

    Two row partitions are in play.  The COMPUTE partition is the shard's (nnz-balanced, so the
    row counts are very unequal: on R-MAT scale 27 one of 8 ranks owns 42 % of the rows).  The
    HOST-I/O partition is even: rank q's host buffers hold rows q*n/W .. (q+1)*n/W of x and of y, so
    every rank moves n/W values each way over its own PCIe link whatever it computes (with the I/O
    tied to the compute rows, the rank owning 42 % of them bounded the call: 7.3 ms per step at 8
    GPUs against a 2.2 ms device step).  Between the two, NVLink: the even slices of x are
    all-gathered into a full-length device x; after the SpMV the computed rows go to the ranks
    whose host slices they belong to with one all-to-all (uneven splits, each row sent once).
    `slots` calls are in flight at once (own stream and device buffers each), like
    matrix.CsrMatrix.spmv_many on one GPU.  All ranks must submit in the same order (the
    collectives pair up by issue order)."""

    def __init__(self, shard: Shard, n_global: int, kind: str = "auto", slots: int = 3, group=None):
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise RuntimeError("ShardedHostSpMV needs a CUDA device; there is no CPU path")
        if shard.csr.n_cols != n_global:
            raise ValueError("the shard must carry the global column count")
        self.dist, self.group = dist, group
        self.shard, self.n, self.kind = shard, int(n_global), kind
        self.world, self.rank = shard.world, shard.rank
        self.io_bounds = even_bounds(self.n, self.world)
        self.io_begin, self.io_end = self.io_bounds[self.rank], self.io_bounds[self.rank + 1]
        # rows this rank computes, split by the host slice they go to; rows of this rank's host
        # slice, split by the rank that computes them (ascending in both, so pieces arrive in order)
        self.send_split = overlap_sizes(shard.row_bounds, self.rank, self.io_bounds)
        self.recv_split = overlap_sizes(self.io_bounds, self.rank, shard.row_bounds)
        self.even = all(self.io_bounds[q + 1] - self.io_bounds[q] == self.io_end - self.io_begin
                        for q in range(self.world))
        dt = shard.csr.Ax.dtype
        self._slots = []
        for _ in range(int(slots)):
            x = torch.empty(self.n, dtype=dt, device="cuda")
            self._slots.append({
                "stream": torch.cuda.Stream(), "x": x,
                "y": torch.empty(shard.csr.n_rows, dtype=dt, device="cuda"),
                "y_io": torch.empty(self.io_end - self.io_begin, dtype=dt, device="cuda"),
                "views": [x[self.io_bounds[q]:self.io_bounds[q + 1]] for q in range(self.world)]})
        self._calls = 0
        torch.cuda.synchronize()

    @property
    def n_slots(self) -> int:
        return len(self._slots)

    def submit(self, slot: int, x_local: torch.Tensor, y_local: torch.Tensor) -> None:
        """x_local / y_local: this rank's host slices, rows io_begin .. io_end, CPU tensors (pinned
        for the copies to overlap); untouched until wait(slot)."""
        m, sl = self.shard.csr, self._slots[slot]
        rows = self.io_end - self.io_begin
        if x_local.numel() != rows or y_local.numel() != rows or x_local.dtype != m.Ax.dtype \
                or y_local.dtype != m.Ax.dtype or x_local.is_cuda or y_local.is_cuda:
            raise ValueError("x_local / y_local must be CPU tensors of rows io_begin..io_end and the matrix dtype")
        with torch.cuda.stream(sl["stream"]):
            mine = sl["views"][self.rank]
            mine.copy_(x_local, non_blocking=True)
            if self.world > 1:
                if self.even:
                    self.dist.all_gather_into_tensor(sl["x"], mine, group=self.group)
                else:
                    self.dist.all_gather(sl["views"], mine, group=self.group)
            # the shard is resident and never changes: every call after the first vouches for it
            # (tile coordinates are cached per stream and checked against a tag, so a slot's
            # first flagged call still searches)
            spmv_mod.spmv_ex(self.kind, m.Ap, m.Aj, m.Ax, sl["x"], sl["y"], n_cols=self.n,
                             stream=sl["stream"], static_pattern=self._calls > 0)
            if self.world > 1:
                self.dist.all_to_all_single(sl["y_io"], sl["y"], output_split_sizes=self.recv_split,
                                            input_split_sizes=self.send_split, group=self.group)
                y_local.copy_(sl["y_io"], non_blocking=True)
            else:
                y_local.copy_(sl["y"], non_blocking=True)
        self._calls += 1

    def wait(self, slot: int) -> None:
        self._slots[slot]["stream"].synchronize()

    def spmv_many(self, xs, ys) -> None:
        """ys[i] = (A @ x_i)[io_begin:io_end] for a sequence of right-hand sides given by host slices."""
        k, n = self.n_slots, 0
        for i, (x, y) in enumerate(zip(xs, ys)):
            if i >= k:
                self.wait(i % k)
            self.submit(i % k, x, y)
            n = i + 1
        for s in range(min(n, k)):
            self.wait(s)

    def close(self):
        torch.cuda.synchronize()
        self._slots = []


def init_distributed():
    """(rank, world, local_rank) from torchrun's environment; NCCL over NVLink when world > 1."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local_rank
