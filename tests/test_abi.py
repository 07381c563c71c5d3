"""The C-ABI library loads on a CPU-only host, exports every symbol include/tpg_b200.h
declares, agrees with the header on every enum, and its host-only helpers agree with the
oracle.  No compute calls (no GPU here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from tidypopgen_b200 import _lib as L

HEADER = (Path(__file__).resolve().parent.parent / "include" / "tpg_b200.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(tpg_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = declared_functions()
    assert "tpg_pairwise" in names and "tpg_accumulate" in names and len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/tpg_b200.h but not exported"
    assert set(names) == set(L.SIGNATURES), "ctypes signatures out of sync with the header"


def test_enums_match_header():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    vals = {k: int(v) for k, v in re.findall(r"\b(TPG_[A-Z0-9_]+)\s*=\s*(\d+)", body)}
    vals.update({k: int(v) for k, v in re.findall(r"#define\s+(TPG_[A-Z0-9_]+)\s+(\d+)", body)})
    for k, v in vals.items():
        py = k[4:]
        assert hasattr(L, py), f"{k} missing from _lib.py"
        assert getattr(L, py) == v, k


def test_struct_layout_matches_c(tmp_path):
    """Every struct of the header as the C compiler lays it out (a probe compiled from include/tpg_b200.h
    prints sizeof / offsetof) against the ctypes mirrors in _lib.py, field by field."""
    import subprocess
    root = Path(__file__).resolve().parent.parent
    structs = {"tpg_job": L.Job, "tpg_stats": L.Stats, "tpg_result": L.Result, "tpg_config": L.Config}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tpg_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    r = subprocess.run(["gcc", "-std=c99", f"-I{root / 'include'}", str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == C.sizeof(cls), cname
        for f, _ in cls._fields_:
            assert int(out[f"{cname}.{f}"]) == getattr(cls, f).offset, f"{cname}.{f}"
    assert L.Job.bk_path.offset == 16 and L.Result.stats.offset == 56


def test_config_roundtrip_and_validation(tpg_cfg):
    """tpg_config: defaults, set / get round trip, validation errors, reset — host only."""
    base = tpg_cfg()
    assert base["abi_version"] == L.ABI_VERSION and base["gram_route"] == 0 and base["cache_counts"] == 0
    got = tpg_cfg(gram_route=1, io_mode=1, host_threads=3, link_gbps=12.5, device_chunk_loci=4096)
    assert (got["gram_route"], got["io_mode"], got["host_threads"], got["link_gbps"], got["device_chunk_loci"]) == (1, 1, 3, 12.5, 4096)
    assert tpg_cfg() == got
    assert base["cache_genotypes"] == 0 and base["result_shells"] == 0 and base["scan_table"] == 0
    for bad in ({"gram_route": 7}, {"pack_mode": 3}, {"io_mode": -1}, {"host_threads": -2}, {"link_gbps": -1.0},
                {"result_shells": 9}, {"result_shells": -1}, {"scan_table": 4}):
        with pytest.raises(L.TpgError) as ei:
            tpg_cfg(**bad)
        assert ei.value.code == L.ERR_ARG
    assert tpg_cfg() == got, "a rejected configuration must not be applied"
    with pytest.raises(KeyError):
        tpg_cfg(no_such_field=1)
    assert tpg_cfg(reset=True)["gram_route"] == 0


def test_no_env_switches_outside_config():
    """The TPG_* environment variables are read in ONE place (csrc/config.cpp, tpg_config_init)."""
    csrc = Path(__file__).resolve().parent.parent / "tidypopgen_b200" / "csrc"
    for f in list(csrc.glob("*.cu")) + list(csrc.glob("*.cpp")) + list(csrc.glob("*.h")) + list(csrc.glob("*.cuh")):
        if f.name == "config.cpp":
            continue
        assert 'getenv("TPG_' not in f.read_text(), f.name


def test_version_and_device_count_without_gpu():
    lib = L.lib()
    assert b"sm_100a" in lib.tpg_version()
    assert lib.tpg_device_count() >= 0


@pytest.mark.parametrize("n", [1, 7, 12, 929, 2504, 10000, 50000, 10**6, 2 * 10**8])
def test_block_size_matches_oracle(n):
    assert L.lib().tpg_block_size(n) == O.block_size(n)


def test_split_len_matches_oracle():
    lib = L.lib()
    rng = np.random.default_rng(11)
    cases = [(961, 300), (6, 3), (6, 2), (5, 100), (650000, 144475), (1000000, 53601), (500000, 2684),
             (1000000, 13421), (7, 7), (7, 1), (10, 3), (25, 10)]
    cases += [(int(rng.integers(1, 200000)), int(rng.integers(1, 5000))) for _ in range(200)]
    for m, bs in cases:
        ref = O.split_len(m, bs)
        nb = lib.tpg_split_len(m, bs, None, None, None, 0)
        assert nb == len(ref)
        lo, up, sz = (np.zeros(nb, np.int64) for _ in range(3))
        lib.tpg_split_len(m, bs, lo.ctypes.data, up.ctypes.data, sz.ctypes.data, nb)
        assert np.array_equal(np.stack([lo, up, sz], 1), ref), (m, bs)
        assert sz.sum() == m and up[-1] == m and sz.max() - sz.min() <= 1
        assert lib.tpg_as_pad_count(m, bs) == O.n_short(m, bs)
    assert lib.tpg_split_len(0, 5, None, None, None, 0) == -1


def test_counts_layout():
    lib = L.lib()
    assert lib.tpg_counts_ld(929) == 1024 and lib.tpg_counts_ld(1024) == 1024 and lib.tpg_counts_ld(1) == 256
    assert lib.tpg_counts_elems(929) == 4 * 1024 * 1024


def test_synth_host_is_deterministic_and_in_law():
    lib = L.lib()
    a = np.zeros(50 * 4000, np.uint8)
    b = np.zeros(50 * 4000, np.uint8)
    lib.tpg_synth_fill_host(a.ctypes.data, 50, 4000, 0, 42, 0.02, 0)
    lib.tpg_synth_fill_host(b.ctypes.data, 50, 2000, 2000, 42, 0.02, 0)
    assert np.array_equal(a[50 * 2000:], b[:50 * 2000]), "shards must be independently generable"
    assert set(np.unique(a)) <= {0, 1, 2, 3}
    assert 0.01 < (a == 3).mean() < 0.03
    dos = a[a < 3].mean() / 2  # mean allele frequency ~ E U(0.05, 0.5) = 0.275
    assert 0.25 < dos < 0.30
    c = np.zeros(50 * 4000, np.uint8)
    lib.tpg_synth_fill_host(c.ctypes.data, 50, 4000, 0, 42, 0.02, 1)
    assert set(np.unique(c)) <= {0, 1, 2, 4, 5, 6} and np.array_equal(c < 3, a < 3)


def test_compute_fails_loudly_without_gpu():
    if L.lib().tpg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(L.TpgError) as ei:
        L.Context(0)
    assert ei.value.code == L.ERR_CUDA and "no CPU path" in ei.value.msg
    import tidypopgen_b200 as tpg
    X = tpg.FBMCode256.from_matrix(np.array([[0.0, 1, 2], [2, 1, 0]]))
    with pytest.raises(L.TpgError):
        tpg.snp_king(X)


def test_r_glue_type_checks_against_the_header():
    """r/src/gpu_glue.cpp (the Rcpp binding INTEGRATION.md hands to a tidypopgen maintainer) is compiled with
    -fsyntax-only against include/tpg_b200.h and a minimal stand-in for <Rcpp.h> (tests/stubs): argument
    counts and pointer types of every tpg_* call it makes are checked, so the written-out binding cannot
    drift away from the ABI unnoticed (R itself is not in this image)."""
    import subprocess
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", f"-I{root / 'tests' / 'stubs'}",
                        f"-I{root / 'include'}", str(root / "r" / "src" / "gpu_glue.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the header itself is plain C: it must also compile as C
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-Wall", "-Werror", "-x", "c", str(root / "include" / "tpg_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_r_package_build_through_makevars(tmp_path):
    """north_star: "the R wrappers .Call into a thin C-ABI shared library built from new src/*.cu files through
    src/Makevars" (reference: src/Makevars:1-9).  r/install_into_tidypopgen.sh copies the backend into a
    package tree; the nvcc rules of r/src/Makevars then build every object the way `R CMD INSTALL` would
    (R's own variables stubbed, <Rcpp.h> from tests/stubs), and the linked shared object exports the ABI."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    root = Path(__file__).resolve().parent.parent
    pkg = tmp_path / "tidypopgen"
    (pkg / "src").mkdir(parents=True)
    (pkg / "R").mkdir()
    (pkg / "DESCRIPTION").write_text("Package: tidypopgen\nVersion: 0.0.0\n")
    for f in ("snp_king.R", "snp_ibs.R", "snp_allele_sharing.R"):
        (pkg / "R" / f).write_text("# reference driver\n")
    r = subprocess.run(["bash", str(root / "r" / "install_into_tidypopgen.sh"), str(pkg)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert (pkg / "R" / "snp_gpu_drivers.R").exists() and not (pkg / "R" / "snp_king.R").exists()
    objs = ["api.o", "decode.o", "gram.o", "finish.o", "relate.o", "wg17.o", "peak.o", "pack.o", "config.o", "gpu_glue.o"]
    stubs = f"-I. -I{root / 'tests' / 'stubs'}"
    r = subprocess.run(["make", "-j8", "-f", "Makevars", "CXX=g++", f"ALL_CPPFLAGS={stubs}", f"CPPFLAGS={stubs}",
                        "CXXFLAGS=-std=c++17 -fPIC -O1"] + objs, cwd=pkg / "src", capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("arch=compute_100a,code=sm_100a") == 7 and "--fmad=false" in r.stdout
    so = pkg / "src" / "tidypopgen.so"
    r = subprocess.run(["g++", "-shared", "-o", str(so)] + objs + ["-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl",
                        "-lrt", "-lpthread"], cwd=pkg / "src", capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(str(so))
    for name in declared_functions():
        assert hasattr(lib, name), name
