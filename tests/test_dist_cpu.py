"""CPU suite, world_size 2 over gloo: the exchange protocol of the row-sharded power iteration
(spmv_samples_b200/dist.py) -- nnz-balanced row split, double-buffered x, uneven all-gather,
norm all-reduce as the step barrier, lagged device-style alpha -- gives the same iterates as
the single-process iteration.  The four device operations of a step are replaced by the CPU
oracle in a subclass that lives HERE (HostProtocolIteration): the product class has no host
arithmetic and no hook for any; the CUDA kernels are covered by the -m gpu suite."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cpu, generators as g
from spmv_samples_b200 import generate
from spmv_samples_b200.dist import PowerIteration, Shard


class HostProtocolIteration(PowerIteration):
    """PowerIteration with its device operations replaced by the CPU oracle (cpu_navie.hpp:3-17)
    and CPU tensors, so that the exchange protocol runs over gloo.  Test infrastructure."""

    def _setup_buffers(self):
        self.xbuf = [torch.zeros(self.n, dtype=self.dtype) for _ in range(2)]
        self._full = list(self.xbuf)
        if self.exchange != "none":
            self.exchange = "nccl"      # the collective path; there is nothing to map into peers
        return "cpu"

    def _sync(self):
        pass

    def _local_spmv(self, x, y, peers):
        csr = self.shard.csr
        out = cpu.spmv(csr.Ap.numpy(), csr.Aj.numpy(), csr.Ax.numpy(), x.numpy())
        y.copy_(torch.from_numpy(out * np.float32(self.alpha.item())))

    def _local_sumsq(self, y):
        self.sumsq[0] = (y.double() ** 2).sum()

    def _alpha_from_sumsq(self):
        s2 = float(self.sumsq[0])
        self.alpha[0] = 1.0 / (s2 ** 0.5) if s2 > 0 else 1.0


def make_shard(Ap, Aj, Ax, rank, world):
    n = Ap.shape[0] - 1
    rb = cpu.row_split(Ap, world).tolist() if world > 1 else [0, n]
    r0, r1 = rb[rank], rb[rank + 1]
    k0, k1 = int(Ap[r0]), int(Ap[r1])
    local = generate.Csr(r1 - r0, n, k1 - k0, torch.from_numpy((Ap[r0:r1 + 1] - k0).copy()),
                         torch.from_numpy(Aj[k0:k1].copy()), torch.from_numpy(Ax[k0:k1].copy()), "t")
    return Shard(rank, world, r0, r1, k0, k1, rb, local)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, steps, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Ap, Aj, Ax = g.rmat(10, 16, 7)
    shard = make_shard(Ap, Aj, Ax, rank, world)
    it = HostProtocolIteration(shard, Ap.shape[0] - 1, exchange="nccl")
    assert it.exchange == "nccl"
    for _ in range(steps):
        it.step()
    np.save(os.path.join(out_dir, f"x_{rank}.npy"), it.current_x().numpy())
    np.save(os.path.join(out_dir, f"s_{rank}.npy"), it.sumsq.numpy())
    it.close()
    dist.destroy_process_group()


def test_two_rank_power_iteration_matches_single_process(tmp_path):
    steps = 6
    mp.spawn(_worker, args=(2, _free_port(), steps, str(tmp_path)), nprocs=2, join=True)
    x0 = np.load(tmp_path / "x_0.npy")
    x1 = np.load(tmp_path / "x_1.npy")
    assert np.array_equal(x0, x1)                      # both replicas hold the same iterate
    assert np.array_equal(np.load(tmp_path / "s_0.npy"), np.load(tmp_path / "s_1.npy"))
    # single process, same arithmetic: y = alpha * (A x), alpha = 1/||y_prev||
    Ap, Aj, Ax = g.rmat(10, 16, 7)
    n = Ap.shape[0] - 1
    x = np.full(n, 1.0 / np.sqrt(n), dtype=np.float32)
    alpha = np.float32(1.0)
    for _ in range(steps):
        y = cpu.spmv(Ap, Aj, Ax, x) * alpha
        # the two ranks add their partial sums of squares in rank order
        rb = cpu.row_split(Ap, 2)
        ss = sum(float((y[rb[q]:rb[q + 1]].astype(np.float64) ** 2).sum()) for q in range(2))
        alpha = np.float32(1.0 / np.sqrt(ss))
        x = y
    assert np.array_equal(x0, x)
    assert abs(float(np.load(tmp_path / "s_0.npy")[0]) - ss) <= 1e-12 * ss


def test_row_split_is_nnz_balanced_and_covers_all_rows():
    Ap, _, _ = g.rmat(12, 16, 3)
    n, nnz = Ap.shape[0] - 1, int(Ap[-1])
    for parts in (2, 4, 8):
        rb = cpu.row_split(Ap, parts)
        assert rb[0] == 0 and rb[-1] == n and np.all(np.diff(rb) >= 0)
        work = np.array([(rb[q + 1] - rb[q]) + (Ap[rb[q + 1]] - Ap[rb[q]]) for q in range(parts)])
        ideal = (n + nnz) / parts
        # each shard is within one (longest) row of the ideal share of the merge path
        assert np.all(np.abs(work - ideal) <= np.diff(Ap).max() + 1)


def test_power_iteration_refuses_to_run_without_cuda_or_hook():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    Ap, Aj, Ax = g.rmat(6, 4, 1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        PowerIteration(make_shard(Ap, Aj, Ax, 0, 1), Ap.shape[0] - 1)


# ------------------------------------------------------------------ weighted / re-balanced split
def test_weighted_split_restatement_is_the_merge_path_at_weight_one():
    from oracle import cpu, generators as g
    from spmv_samples_b200 import spmv
    for Ap in (g.rmat(11, 16, 3, offset_dtype=np.int64)[0], g.ragged(2000, 2000, 7.0, 2)[0],
               g.lap2d(20)[0]):
        n_rows, nnz = Ap.shape[0] - 1, int(Ap[-1])
        for parts in (2, 3, 5, 8):
            t = spmv.split_targets(n_rows + nnz, parts)
            assert [0] + cpu.rows_at_cost(Ap, t, (1, 1)).tolist() + [n_rows] == cpu.row_split(Ap, parts).tolist()
        # weight 0: pure nonzero balance, every shard within one row of nnz / parts
        t = spmv.split_targets(nnz, 4)
        rb = [0] + cpu.rows_at_cost(Ap, t, (0, 1)).tolist() + [n_rows]
        longest = int(np.diff(Ap).max())
        for q in range(4):
            assert abs(int(Ap[rb[q + 1]] - Ap[rb[q]]) - nnz / 4) <= longest + 1


def test_rebalance_targets_moves_boundaries_towards_equal_time():
    from spmv_samples_b200 import spmv
    cost = [0, 1000, 2000, 3000, 4000]
    # equal times: nothing moves
    assert spmv.rebalance_targets(cost, [1.0, 1.0, 1.0, 1.0]) == [1000, 2000, 3000]
    # shard 0 took twice as long as the others: it must shrink, the targets stay ordered
    t = spmv.rebalance_targets(cost, [2.0, 1.0, 1.0, 1.0])
    assert t[0] < 1000 and t == sorted(t) and all(0 <= v <= 4000 for v in t)
    assert abs(t[0] - 625) <= 1          # 5/4 time units per shard -> 1.25 / 2 of shard 0's cost
    # a model where time really is piecewise uniform in cost is balanced in one round
    dens = [2.0, 1.0, 1.0, 1.0]
    def shard_time(lo, hi):
        return sum(dens[g] * max(0, min(hi, cost[g + 1]) - max(lo, cost[g])) / 1000.0 for g in range(4))
    b = [0] + t + [4000]
    times = [shard_time(b[q], b[q + 1]) for q in range(4)]
    assert max(times) - min(times) <= 0.01 * max(times)
    # degenerate inputs
    assert spmv.rebalance_targets([0, 10], [3.0]) == []
    assert spmv.rebalance_targets(cost, [0.0, 0.0, 0.0, 0.0]) == [1000, 2000, 3000]


def test_host_buffer_front_ends_refuse_to_run_without_cuda():
    """No CPU fallback in the product: the sharded host-buffer call and the power iteration both
    raise when there is no CUDA device (the GPU suite covers the real path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    from spmv_samples_b200 import generate
    from spmv_samples_b200.dist import PowerIteration, Shard, ShardedHostSpMV
    Ap = torch.tensor([0, 1, 2], dtype=torch.int32)
    csr = generate.Csr(2, 2, 2, Ap, torch.tensor([0, 1], dtype=torch.int32), torch.ones(2), "tiny")
    shard = Shard(0, 1, 0, 2, 0, 2, [0, 2], csr)
    with pytest.raises(RuntimeError, match="CUDA"):
        ShardedHostSpMV(shard, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        PowerIteration(shard, 2)


# ------------------------------------------------------------------ host-I/O redistribution
def test_overlap_sizes_partition_both_ways():
    """ShardedHostSpMV moves y from the compute partition (uneven, nnz-balanced) to the even
    host-I/O partition with one all-to-all: the send splits of all ranks must tile every
    destination slice exactly, in row order."""
    from spmv_samples_b200.dist import even_bounds, overlap_sizes
    for n, rb in ((100, [0, 3, 50, 51, 100]), (17, [0, 0, 17, 17, 17]), (8, [0, 2, 4, 6, 8]), (5, [0, 5, 5, 5, 5])):
        w = len(rb) - 1
        io = even_bounds(n, w)
        assert io[0] == 0 and io[-1] == n and all(b - a in (n // w, n // w + 1) for a, b in zip(io, io[1:]))
        send = [overlap_sizes(rb, p, io) for p in range(w)]
        recv = [overlap_sizes(io, q, rb) for q in range(w)]
        for p in range(w):
            assert sum(send[p]) == rb[p + 1] - rb[p]
            for q in range(w):
                assert send[p][q] == recv[q][p]
        for q in range(w):
            assert sum(recv[q]) == io[q + 1] - io[q]


def _a2a_worker(rank, world, port, out_dir):
    from spmv_samples_b200.dist import even_bounds, overlap_sizes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, rb = 1001, [0, 13, 700, 1001][:world] + [1001]
    rb = [0, 13, 1001] if world == 2 else [0, 13, 700, 1001]
    io = even_bounds(n, world)
    y_full = torch.arange(n, dtype=torch.float32) * 0.5
    mine = y_full[rb[rank]:rb[rank + 1]].clone()                 # the rows this rank "computed"
    out = torch.full((io[rank + 1] - io[rank],), float("nan"))
    dist.all_to_all_single(out, mine, output_split_sizes=overlap_sizes(io, rank, rb),
                           input_split_sizes=overlap_sizes(rb, rank, io))
    assert torch.equal(out, y_full[io[rank]:io[rank + 1]])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_all_to_all_moves_rows_between_partitions(world, tmp_path):
    mp.spawn(_a2a_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
