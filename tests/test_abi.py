"""CPU suite: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/spmv_b200.h declares.  No compute calls are made without a GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    from spmv_samples_b200 import _lib
    names = _lib.exported_symbols()
    # 5 kinds x 4 type pairs + the untyped and auxiliary entry points
    assert len([n for n in names if re.match(r"spmvb200_(merge|vector|light|auto|cusparse)_i32_", n)]) == 20
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(built_lib, n)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                         text=True, check=True).stdout
    exported = set(re.findall(r" T (spmvb200_\w+)", out))
    assert set(names) <= exported
    # nothing but the declared ABI leaks out of the library
    assert exported <= set(names), exported - set(names)


def test_library_is_sm100a_with_tma_bulk_copies(built_lib):
    from spmv_samples_b200 import _lib
    r = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True,
                          text=True).stdout
    assert "merge_tile_reg_kernel" in sass and "merge_tile_tma_kernel" in sass
    assert "UBLKCP" in sass      # cp.async.bulk staging (the TMA-staged ablation variant)
    assert "SYNCS" in sass       # mbarrier completion
    assert re.search(r"LDG\.E\.[A-Z.]*128", sass)   # 128-bit loads of Aj / Ax


def test_status_strings_and_options(built_lib):
    assert built_lib.spmvb200_status_string(0) == b"ok"
    assert b"aligned" in built_lib.spmvb200_status_string(2)
    assert built_lib.spmvb200_version().startswith(b"spmvb200")
    assert built_lib.spmvb200_get_option(b"l2_window") in (0, 1)
    assert built_lib.spmvb200_set_option(b"no_such_option", 1) != 0
    assert built_lib.spmvb200_merge_tile_items(32, 32) > 0


def test_python_mirror_has_reference_surface():
    from spmv_samples_b200 import spmv
    ours = {"merge", "vector", "light", "stream", "auto", "cusparse"}
    # every label of the reference's table (spmv.h:18-27) is accepted; "merge_genl" lives behind
    # spmv_ex(semiring=...) in the Python mirror and is an X line of include/spmv.h
    reference = {"cusparse", "cusp", "cusp1", "cusp2", "light_vec", "light_warp", "cub_merge", "merge"}
    assert set(spmv.SPMV_KINDS) == ours | reference
    assert set(spmv.KIND_IDS) == set(spmv.SPMV_KINDS)
    hdr = open(os.path.join(ROOT, "include", "spmv.h")).read()
    for label in ours | reference | {"merge_genl"}:
        assert f'X("{label}",' in hdr, label
    import inspect
    params = list(inspect.signature(spmv.SpMV).parameters)
    assert params[:9] == ["kind_str", "n_rows", "n_cols", "nnz", "Ap", "Aj", "Ax", "x", "y"]
    for fn in spmv.SPMV_KINDS.values():
        assert list(inspect.signature(fn).parameters)[:8] == [
            "n_rows", "n_cols", "nnz", "Ap", "Aj", "Ax", "x", "y"]


def test_unknown_kind_is_an_error(capsys):
    from spmv_samples_b200 import spmv
    with pytest.raises(spmv.SpMVKindError):
        spmv.SpMV("no_such_kind", 0, 0, 0, None, None, None, None, None)
    assert "NOT SUPPROT" in capsys.readouterr().err


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under spmv_samples_b200/, include/ or main.cu
    may reference it, and there is no CPU fallback in the Python binding."""
    bad = []
    for base in ("spmv_samples_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep):
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "liboracle" in text \
                            or "libspmv_ref" in text:
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_argument_validation_needs_no_device(built_lib):
    """Argument checks come before any CUDA call, so they can be exercised here: the weighted
    split search, the untyped entry's flag word, the SpMM shape checks."""
    import ctypes as C
    from spmv_samples_b200 import _lib
    L = built_lib
    t = (C.c_int64 * 2)(5, 9)
    out = (C.c_int64 * 2)(-1, -1)
    ap = (C.c_int32 * 4)(0, 1, 2, 3)
    f = L.spmvb200_rows_at_cost_o32
    assert f(3, ap, 1, 0, 2, t, out, None) == 1                 # w_den < 1
    assert f(3, ap, -1, 1, 2, t, out, None) == 1                # w_num < 0
    assert f(3, ap, (1 << 20) + 1, 1, 2, t, out, None) == 4     # weight out of the int64-safe range
    assert f(3, ap, 1, 1, 0, None, None, None) == 0             # nothing asked
    assert f(3, ap, 1, 1, 2, None, out, None) == 1              # targets missing
    neg = (C.c_int64 * 2)(5, -1)
    assert f(3, ap, 1, 1, 2, neg, out, None) == 1               # negative cost
    assert f(0, None, 1, 1, 2, t, out, None) == 0 and list(out) == [0, 0]   # no rows: row 0 for every target
    a = _lib.Args()
    a.flags = 4
    assert L.spmvb200_spmv(C.byref(a)) == 1                     # unknown flag bit
    s = _lib.SpmmArgs()
    s.k = 3
    assert L.spmvb200_spmm(C.byref(s)) == 4                     # k must be 2, 4 or 8
    assert L.spmvb200_matrix_wait(None, 0) == 1 and L.spmvb200_matrix_submit_host(None, 0, 9, None, None) == 1
