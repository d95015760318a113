"""GPU suite: power-iteration helpers, the single-GPU PowerIteration against the oracle, the
host-buffer matrix object over device arrays, the dominant-kernel timer, and (when the box has
at least two GPUs) the two-rank row-sharded iteration with both exchanges."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from oracle import cpu, generators as g

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_cuda(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_sum_squares_and_inv_sqrt(dt):
    from spmv_samples_b200 import _lib
    L = _lib.lib()
    v = g.gen_x(3, 1_000_003, dt)
    d = dev(v)
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    alpha = torch.zeros(1, dtype=d.dtype, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    bits = 32 if dt == np.float32 else 64
    _lib.check(L.spmvb200_sum_squares(bits, d.numel(), d.data_ptr(), ss.data_ptr(), s), "sumsq")
    _lib.check(L.spmvb200_inv_sqrt(bits, ss.data_ptr(), alpha.data_ptr(), s), "inv_sqrt")
    torch.cuda.synchronize()
    expect = float((v.astype(np.float64) ** 2).sum())
    assert abs(float(ss.item()) - expect) <= 1e-12 * expect
    assert abs(float(alpha.item()) - expect ** -0.5) <= 1e-6 * expect ** -0.5
    # deterministic: same bits every time
    ss2 = torch.zeros(1, dtype=torch.float64, device="cuda")
    _lib.check(L.spmvb200_sum_squares(bits, d.numel(), d.data_ptr(), ss2.data_ptr(), s), "sumsq")
    torch.cuda.synchronize()
    assert float(ss2.item()) == float(ss.item())


@pytest.mark.parametrize("kind", ["merge", "vector", "auto"])
def test_single_gpu_power_iteration_tracks_the_oracle(kind):
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.dist import PowerIteration, shard_rows
    m = gen.rmat(12, 16, 7, offset=torch.int64)
    Ap, Aj, Ax = g.rmat(12, 16, 7, offset_dtype=np.int64)
    n = m.n_rows
    it = PowerIteration(shard_rows(m, 0, 1), n, kind=kind)
    x = np.full(n, 1.0 / np.sqrt(n), dtype=np.float64)
    alpha = 1.0
    for _ in range(8):
        it.step()
        y = cpu.spmv_fp64(Ap, Aj, Ax, x.astype(np.float32)) * alpha
        alpha = 1.0 / np.sqrt((y ** 2).sum())
        x = y
    torch.cuda.synchronize()
    got = it.current_x().cpu().numpy().astype(np.float64)
    # fp32 iterates drift from the fp64 recurrence by rounding only
    assert np.linalg.norm(got - x) <= 1e-4 * np.linalg.norm(x)
    assert abs(it.eigen_estimate() - np.sqrt((x ** 2).sum())) <= 1e-4 * np.sqrt((x ** 2).sum())
    it.close()


def test_matrix_object_over_device_arrays():
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.matrix import CsrMatrix
    m = gen.uniform_rows(5000, 5000, 8, 4)
    Ap, Aj, Ax = g.uniform_rows(5000, 5000, 8, 4)
    x = g.gen_x(9, 5000)
    mat = CsrMatrix.from_device(m)
    y = mat.spmv(x, kind="auto")
    y64 = cpu.spmv_fp64(Ap, Aj, Ax, x)
    assert np.all(np.abs(y - y64) <= 1e-5 * cpu.abs_scale(Ap, Aj, Ax, x))
    mat.close()
    assert int(m.Ap[-1]) == m.nnz      # borrowed arrays are still alive


def test_dominant_kernel_timer():
    from spmv_samples_b200 import _lib, generate as gen, spmv
    m = gen.uniform_rows(200000, 200000, 16, 4)
    x = gen.gen_x(m.n_cols, 1)
    y = torch.empty(m.n_rows, device="cuda")
    ms, cnt = C.c_double(), C.c_int64()
    spmv.set_option("time_main_kernel", 1)
    try:
        _lib.lib().spmvb200_main_kernel_time(C.byref(ms), C.byref(cnt))
        for kind in ("merge", "vector", "light"):
            for _ in range(4):
                spmv.SpMV(kind, m.n_rows, m.n_cols, m.nnz, m.Ap, m.Aj, m.Ax, x, y)
        _lib.lib().spmvb200_main_kernel_time(C.byref(ms), C.byref(cnt))
    finally:
        spmv.set_option("time_main_kernel", 0)
    assert cnt.value == 12 and 0.0 < ms.value < 1000.0
    _lib.lib().spmvb200_main_kernel_time(C.byref(ms), C.byref(cnt))
    assert cnt.value == 0


# ------------------------------------------------------------------ two ranks, two GPUs
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, exchange, out_dir, rebalance=False):
    import torch.distributed as dist
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.dist import PowerIteration, shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = gen.rmat(19 if rebalance else 14, 16, 7, offset=torch.int64)
    shard = shard_rows(m, rank, world)
    it = PowerIteration(shard, m.n_rows, kind="auto", exchange=exchange)
    if rebalance:
        # a deliberately lopsided split first, then two rounds of re-splitting from measured times
        it.set_shard(shard_rows(m, rank, world, row_bounds=[0, m.n_rows // 64, m.n_rows]))
        for _ in range(2):
            it.step()
            it.rebalance(m, steps=3)
    for _ in range(6):
        it.step()
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"x_{exchange}_{rank}.npy"), it.current_x().cpu().numpy())
    open(os.path.join(out_dir, f"ex_{exchange}_{rank}.txt"), "w").write(
        it.exchange + " " + ",".join(map(str, it.shard.row_bounds)))
    from spmv_samples_b200 import spmv
    open(os.path.join(out_dir, f"tbl_{exchange}_{rank}.txt"), "w").write(
        str(spmv.hot_x_info(it.shard.csr.Aj)["table_columns"]))
    it.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["mc", "p2p", "nccl"])
def test_two_gpu_sharded_iteration_matches_single_gpu(tmp_path, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.dist import PowerIteration, shard_rows
    mp.spawn(_rank_main, args=(2, _free_port(), exchange, str(tmp_path)), nprocs=2, join=True)
    x0 = np.load(tmp_path / f"x_{exchange}_0.npy")
    x1 = np.load(tmp_path / f"x_{exchange}_1.npy")
    assert np.array_equal(x0, x1)
    used, bounds = open(tmp_path / f"ex_{exchange}_0.txt").read().split(" ")
    Ap, _, _ = g.rmat(14, 16, 7, offset_dtype=np.int64)
    assert bounds == ",".join(map(str, cpu.row_split(Ap, 2).tolist()))     # bit-exact split
    m = gen.rmat(14, 16, 7, offset=torch.int64)
    it = PowerIteration(shard_rows(m, 0, 1), m.n_rows, kind="auto")
    for _ in range(6):
        it.step()
    ref = it.current_x().cpu().numpy()
    it.close()
    assert np.linalg.norm(x0.astype(np.float64) - ref) <= 1e-5 * np.linalg.norm(ref)


@pytest.mark.parametrize("exchange", ["mc", "p2p"])
def test_two_gpu_table_kernel_feeds_the_exchange(tmp_path, exchange, monkeypatch):
    """The persistent table form of the merge tile kernel (forced through SPMVB200_OPTS in the rank
    processes: these shards are far too small for it by default) with the fused exchange -- peer
    stores or one multimem.st per row: the replicas agree bit for bit and the iteration is the
    single-GPU one."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.dist import PowerIteration, shard_rows
    monkeypatch.setenv("SPMVB200_OPTS", "hot_x_table=1")
    mp.spawn(_rank_main, args=(2, _free_port(), exchange, str(tmp_path)), nprocs=2, join=True)
    monkeypatch.delenv("SPMVB200_OPTS")
    x0 = np.load(tmp_path / f"x_{exchange}_0.npy")
    x1 = np.load(tmp_path / f"x_{exchange}_1.npy")
    assert np.array_equal(x0, x1)
    assert all(int(open(tmp_path / f"tbl_{exchange}_{r}.txt").read()) > 0 for r in range(2))
    m = gen.rmat(14, 16, 7, offset=torch.int64)
    it = PowerIteration(shard_rows(m, 0, 1), m.n_rows, kind="auto")
    for _ in range(6):
        it.step()
    ref = it.current_x().cpu().numpy()
    it.close()
    assert np.linalg.norm(x0.astype(np.float64) - ref) <= 1e-5 * np.linalg.norm(ref)


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_rebalanced_split_matches_single_gpu(tmp_path, exchange):
    """Start from a lopsided split, re-split twice from measured per-rank kernel times: the
    iteration must still be the single-GPU one, and the split must have moved towards balance."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.dist import PowerIteration, shard_rows
    mp.spawn(_rank_main, args=(2, _free_port(), exchange, str(tmp_path), True), nprocs=2, join=True)
    x0 = np.load(tmp_path / f"x_{exchange}_0.npy")
    x1 = np.load(tmp_path / f"x_{exchange}_1.npy")
    assert np.array_equal(x0, x1)
    b0 = open(tmp_path / f"ex_{exchange}_0.txt").read().split(" ")[1]
    b1 = open(tmp_path / f"ex_{exchange}_1.txt").read().split(" ")[1]
    assert b0 == b1                                     # every rank derived the same split
    m = gen.rmat(19, 16, 7, offset=torch.int64)         # big enough for kernel times to differ
    mid = int(b0.split(",")[1])
    assert m.n_rows // 64 < mid < m.n_rows              # the boundary moved off the lopsided start
    it = PowerIteration(shard_rows(m, 0, 1), m.n_rows, kind="auto")
    for _ in range(6):
        it.step()
    ref = it.current_x().cpu().numpy()
    it.close()
    assert np.linalg.norm(x0.astype(np.float64) - ref) <= 1e-5 * np.linalg.norm(ref)


def _host_rank_main(rank, world, port, out_dir):
    import torch.distributed as dist
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.dist import ShardedHostSpMV, shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m = gen.rmat(14, 16, 7, offset=torch.int64)
    shard = shard_rows(m, rank, world)
    hs = ShardedHostSpMV(shard, m.n_rows, kind="auto", slots=3)
    r0, r1 = hs.io_begin, hs.io_end        # even host slices, whatever rows the rank computes
    assert (r0, r1) == (rank * m.n_rows // world, (rank + 1) * m.n_rows // world)
    xs = [torch.from_numpy(g.gen_x(200 + i, m.n_cols)[r0:r1].copy()).pin_memory() for i in range(7)]
    ys = [torch.full((r1 - r0,), float("nan")).pin_memory() for _ in range(7)]
    hs.spmv_many(xs, ys)
    hs.close()
    np.save(os.path.join(out_dir, f"y_{rank}.npy"), np.stack([y.numpy() for y in ys]))
    np.save(os.path.join(out_dir, f"rows_{rank}.npy"), np.array([r0, r1]))
    dist.destroy_process_group()


def test_two_gpu_sharded_host_buffer_spmv(tmp_path):
    """x and y sharded by rows in host memory: upload of the local slice, all-gather on the
    device, SpMV, download of the local slice, three calls in flight."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_host_rank_main, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    Ap, Aj, Ax = g.rmat(14, 16, 7, offset_dtype=np.int64)
    n = Ap.shape[0] - 1
    got = np.empty((7, n), dtype=np.float32)
    seen = 0
    for rank in range(2):
        r0, r1 = np.load(tmp_path / f"rows_{rank}.npy")
        got[:, r0:r1] = np.load(tmp_path / f"y_{rank}.npy")
        seen += r1 - r0
    assert seen == n
    for i in range(7):
        x = g.gen_x(200 + i, n)
        assert np.all(np.abs(got[i] - cpu.spmv_fp64(Ap, Aj, Ax, x)) <= 1e-5 * cpu.abs_scale(Ap, Aj, Ax, x))


def test_sharded_host_buffer_spmv_single_rank():
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.dist import ShardedHostSpMV, shard_rows
    m = gen.uniform_rows(30000, 30000, 16, 4)
    Ap, Aj, Ax = g.uniform_rows(30000, 30000, 16, 4)
    hs = ShardedHostSpMV(shard_rows(m, 0, 1), m.n_rows, slots=2)
    xs = [torch.from_numpy(g.gen_x(300 + i, m.n_cols)).pin_memory() for i in range(5)]
    ys = [torch.full((m.n_rows,), float("nan")).pin_memory() for _ in range(5)]
    hs.spmv_many(xs, ys)
    for x, y in zip(xs, ys):
        assert np.all(np.abs(y.numpy() - cpu.spmv_fp64(Ap, Aj, Ax, x.numpy()))
                      <= 1e-5 * cpu.abs_scale(Ap, Aj, Ax, x.numpy()))
    with pytest.raises(ValueError):
        hs.submit(0, xs[0][:-1], ys[0])
    hs.close()


@pytest.mark.parametrize("weight", [(1, 1), (0, 1), (1, 4), (3, 1)])
@pytest.mark.parametrize("off", [np.int32, np.int64])
def test_rows_at_cost_matches_the_oracle(weight, off):
    from spmv_samples_b200 import spmv
    rng = np.random.default_rng(5)
    for Ap in (g.rmat(12, 16, 3, offset_dtype=off)[0], g.ragged(3000, 3000, 9.0, 4, offset_dtype=off)[0],
               np.zeros(1, dtype=off), np.zeros(17, dtype=off)):
        n_rows, nnz = Ap.shape[0] - 1, int(Ap[-1])
        total = weight[1] * nnz + weight[0] * n_rows
        targets = sorted(set([0, total, total + 5] + spmv.split_targets(total, 7)
                             + [int(v) for v in rng.integers(0, total + 1, 20)]))
        got = spmv.rows_at_cost(dev(Ap), targets, weight)
        assert got == cpu.rows_at_cost(Ap, targets, weight).tolist()
        if n_rows:
            for parts in (2, 3, 8):
                expect = [0] + cpu.rows_at_cost(Ap, spmv.split_targets(total, parts), weight).tolist() + [n_rows]
                assert spmv.row_split(dev(Ap), parts, weight=weight) == expect
                if weight == (1, 1):
                    assert expect == cpu.row_split(Ap, parts).tolist()   # the merge path itself


@pytest.mark.parametrize("slots", [1, 2, 3, 4])
def test_pipelined_host_buffer_api_matches_serial(slots):
    from spmv_samples_b200 import generate as gen
    from spmv_samples_b200.matrix import CsrMatrix
    m = gen.rmat(15, 16, 3)
    Ap, Aj, Ax = g.rmat(15, 16, 3)
    mat = CsrMatrix.from_device(m)
    xs = [torch.from_numpy(g.gen_x(100 + i, m.n_cols)).pin_memory().numpy() for i in range(7)]
    ys = [torch.empty(m.n_rows, dtype=torch.float32).pin_memory().numpy() for _ in range(7)]
    mat.spmv_many(xs, ys, kind="auto", slots=slots)
    for x, y in zip(xs, ys):
        assert np.array_equal(y, mat.spmv(x, kind="auto"))          # same kernels, same bits
        assert np.all(np.abs(y - cpu.spmv_fp64(Ap, Aj, Ax, x)) <= 1e-5 * cpu.abs_scale(Ap, Aj, Ax, x))
    with pytest.raises(ValueError):
        mat.spmv_many(xs, ys, slots=5)
    mat.close()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_norm_exchange_single_rank(dt):
    """The fused norm kernel with world = 1: sum of squares (deterministic), alpha, slot rotation
    over many steps, the block counter resetting itself, no error raised."""
    from spmv_samples_b200 import _lib
    L = _lib.lib()
    bits = 32 if dt == np.float32 else 64
    mailbox = torch.full((_lib_mailbox_doubles(),), -1.0, dtype=torch.float64, device="cuda")
    ss = torch.zeros(1, dtype=torch.float64, device="cuda")
    alpha = torch.zeros(1, dtype=torch.float32 if dt == np.float32 else torch.float64, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    ptrs = (C.c_void_p * 1)(mailbox.data_ptr())
    s = torch.cuda.current_stream().cuda_stream
    seen = []
    for step in range(7):
        v = g.gen_x(30 + step, 300_007 + 1000 * step, dt)
        d = dev(v)
        _lib.check(L.spmvb200_norm_exchange(bits, d.numel(), d.data_ptr(), 0, 1, step, mailbox.data_ptr(),
                                            ptrs, None, ss.data_ptr(), alpha.data_ptr(), err.data_ptr(), s),
                   "norm_exchange")
        torch.cuda.synchronize()
        expect = float((v.astype(np.float64) ** 2).sum())
        assert abs(float(ss.item()) - expect) <= 1e-12 * expect
        assert abs(float(alpha.item()) - expect ** -0.5) <= 1e-6 * expect ** -0.5
        seen.append(float(ss.item()))
    assert int(err.item()) == 0
    # same input, same bits
    v = g.gen_x(30, 300_007, dt)
    d = dev(v)
    _lib.check(L.spmvb200_norm_exchange(bits, d.numel(), d.data_ptr(), 0, 1, 7, mailbox.data_ptr(), ptrs, None,
                                        ss.data_ptr(), alpha.data_ptr(), err.data_ptr(), s), "norm_exchange")
    torch.cuda.synchronize()
    assert float(ss.item()) == seen[0]


def _lib_mailbox_doubles():
    from spmv_samples_b200.dist import MAILBOX_BYTES
    return MAILBOX_BYTES // 8


# ------------------------------------------------------------------ one process, N GPUs (csrc/multi.cu)
def _native_power(n_gpus, Ap, Aj, Ax, steps, from_device=False, kind=3):
    from spmv_samples_b200 import _lib
    L = _lib.lib()
    n = Ap.shape[0] - 1
    h = C.c_void_p()
    ob, vb = Ap.dtype.itemsize * 8, Ax.dtype.itemsize * 8
    if from_device:
        dAp, dAj, dAx = dev(Ap), dev(Aj), dev(Ax)
        st = L.spmvb200_power_create_from_device(n_gpus, None, ob, vb, n, Aj.size, dAp.data_ptr(), dAj.data_ptr(),
                                                 dAx.data_ptr(), kind, C.byref(h))
    else:
        st = L.spmvb200_power_create(n_gpus, None, ob, vb, n, Aj.size, Ap.ctypes.data, Aj.ctypes.data,
                                     Ax.ctypes.data, kind, C.byref(h))
    _lib.check(st, "spmvb200_power_create")
    try:
        ms = C.c_double()
        _lib.check(L.spmvb200_power_run(h, steps, C.byref(ms)), "spmvb200_power_run")
        x = np.empty(n, dtype=Ax.dtype)
        norm = C.c_double()
        rb = (C.c_int64 * (n_gpus + 1))()
        _lib.check(L.spmvb200_power_get(h, x.ctypes.data, C.byref(norm), rb), "spmvb200_power_get")
        _native_power.exchange = int(L.spmvb200_power_exchange(h))
        # a second run after reset reproduces the first bit for bit
        _lib.check(L.spmvb200_power_reset(h), "spmvb200_power_reset")
        _lib.check(L.spmvb200_power_steps(h, steps), "spmvb200_power_steps")
        x2 = np.empty(n, dtype=Ax.dtype)
        _lib.check(L.spmvb200_power_get(h, x2.ctypes.data, None, None), "spmvb200_power_get")
    finally:
        L.spmvb200_power_destroy(h)
    assert np.array_equal(x, x2)
    return x, norm.value, list(rb), ms.value


def _oracle_power(Ap, Aj, Ax, steps):
    n = Ap.shape[0] - 1
    x = np.full(n, 1.0 / np.sqrt(n), dtype=np.float64)
    alpha = 1.0
    for _ in range(steps):
        y = cpu.spmv_fp64(Ap, Aj, Ax, x.astype(Ax.dtype)) * alpha
        alpha = 1.0 / np.sqrt((y ** 2).sum())
        x = y
    return x


@pytest.mark.parametrize("from_device", [False, True])
@pytest.mark.parametrize("off", [np.int32, np.int64])
def test_native_power_iteration_one_gpu(off, from_device):
    Ap, Aj, Ax = g.rmat(12, 16, 7, offset_dtype=off)
    x, norm, rb, ms = _native_power(1, Ap, Aj, Ax, 8, from_device)
    ref = _oracle_power(Ap, Aj, Ax, 8)
    assert np.linalg.norm(x.astype(np.float64) - ref) <= 1e-4 * np.linalg.norm(ref)
    assert abs(norm - np.sqrt((ref ** 2).sum())) <= 1e-4 * norm
    assert rb == [0, Ap.shape[0] - 1] and ms > 0


@pytest.mark.parametrize("exchange", ["peer_stores", "multicast"])
@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_native_power_iteration_multi_gpu(n_gpus, exchange):
    """Row blocks on n_gpus GPUs of this process, the exchange fused into the SpMV's row stores
    (peer stores, or one multimem.st through an NVLink multicast object: csrc/mcast.cu), mailbox
    norm exchange: same iterates as the oracle recurrence, row split bit-exact against the
    oracle's merge path."""
    from spmv_samples_b200 import spmv
    if torch.cuda.device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    Ap, Aj, Ax = g.rmat(14, 16, 7, offset_dtype=np.int64)
    spmv.set_option("power_exchange", 1 if exchange == "multicast" else 0)
    try:
        x, norm, rb, ms = _native_power(n_gpus, Ap, Aj, Ax, 6)
    except RuntimeError as e:
        if exchange == "multicast" and getattr(e, "status", 0) == 4:   # SPMVB200_ERR_UNSUPPORTED
            pytest.skip("no NVLink multicast on this box")
        raise
    finally:
        spmv.set_option("power_exchange", -1)
    assert _native_power.exchange == (1 if exchange == "multicast" else 0)
    assert rb == cpu.row_split(Ap, n_gpus).tolist()
    ref = _oracle_power(Ap, Aj, Ax, 6)
    assert np.linalg.norm(x.astype(np.float64) - ref) <= 1e-4 * np.linalg.norm(ref)
    assert abs(norm - np.sqrt((ref ** 2).sum())) <= 1e-4 * norm
