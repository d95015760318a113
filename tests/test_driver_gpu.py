"""GPU suite: the C++ side of the boundary -- include/spmv.h's SpMV(kind_str, ...) registry, the
per-kind templates, include/load.hpp and main.cu -- exercised through the built driver
bin/spmv, whose self-check compares every kind with an fp64 host loop and exits non-zero on
any row outside tolerance."""
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "bin", "spmv")


@pytest.fixture(scope="module", autouse=True)
def _driver(built_lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", ROOT, "-s", "bin/spmv"], check=True)


def run(*args):
    return subprocess.run([BIN, *args], capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("cfg", ["synthetic:c1:96", "synthetic:c2:8192", "synthetic:c3:12",
                                 "synthetic:c4:4096", "synthetic:c5:13"])
def test_driver_all_kinds_pass_on_synthetic(cfg):
    kinds = ["merge", "merge_genl", "vector", "light", "auto"] + ([] if "c5" in cfg else ["cusparse"])
    r = run(cfg, *kinds, "--iters", "5", "--x", "random")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Compute delta:" in r.stdout and "Time cost" in r.stdout
    for k in kinds:
        assert re.search(rf"\[{k}\s*\].*PASS", r.stdout), r.stdout
        assert re.search(rf"\[{k}\s*\] total:.*GFLOP/s", r.stdout), r.stdout


@pytest.mark.parametrize("fname", ["general_real.mtx", "symmetric_pattern.mtx", "symmetric_integer.mtx"])
def test_driver_on_matrix_market_files(fname):
    r = run(os.path.join(GOLDEN, fname), "merge", "vector", "light", "auto", "--iters", "3")
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("PASS") == 4 and "FAIL" not in r.stdout
    assert f"Dataset: {fname}" in r.stdout


def test_driver_unknown_kind_exits_like_the_reference():
    r = run("synthetic:c1:16", "no_such_kind", "--iters", "1")
    assert r.returncode != 0
    assert 'SpMV kind "no_such_kind" is NOT SUPPROT' in r.stderr      # spmv.h:46-47


def test_driver_usage_and_missing_file():
    r = run()
    assert r.returncode == 1 and "usage:" in r.stderr
    r = run("/nonexistent/file.mtx", "merge")
    assert r.returncode == 1 and "File could not be opened" in r.stderr  # load.hpp:278-281


def test_driver_accepts_the_reference_command_line():
    """Every label of the reference's SPMV_KINDS (reference/include/spmv.h:18-27) on one command
    line, as a user of the reference would type it, plus the new "stream" kind."""
    kinds = ["cusparse", "cusp", "cusp1", "cusp2", "light_vec", "light_warp", "cub_merge", "merge",
             "merge_genl", "stream"]
    r = run("synthetic:c1:64", *kinds, "--iters", "3", "--x", "random")
    assert r.returncode == 0, r.stdout + r.stderr
    for k in kinds:
        assert re.search(rf"\[{k}\s*\].*PASS", r.stdout), r.stdout


def test_driver_power_iteration_from_one_process():
    """main.cu --power K [--gpus N]: the row-sharded power iteration behind the C ABI
    (csrc/multi.cu), no Python in the loop; on every GPU the box has, up to 8."""
    n = min(torch.cuda.device_count(), 8)
    for gpus in sorted({1, n}):
        r = run("synthetic:c3:14", "merge", "--iters", "2", "--power", "12", "--gpus", str(gpus))
        assert r.returncode == 0, r.stdout + r.stderr
        m = re.search(r"\[power\s*\]\s+([0-9.]+) ms/step.*\|\|A x\|\| = ([0-9.e+-]+)", r.stdout)
        assert m and float(m.group(1)) > 0 and float(m.group(2)) > 0, r.stdout
        rows = re.search(r"rows per GPU:((?: \d+)+)", r.stdout)
        assert rows and sum(int(v) for v in rows.group(1).split()) == 1 << 14
