"""ctypes binding of libtpg_b200.so (include/tpg_b200.h).

The shared library is the product; this module only marshals arguments.  There is no
Python or CPU implementation behind it: if the library is missing, or no B200 is present,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libtpg_b200.so"
CSRC = _HERE / "csrc"

ABI_VERSION = 2

# enums (keep in sync with include/tpg_b200.h; tests/test_abi.py checks the header)
OK, ERR_ARG, ERR_IO, ERR_CUDA, ERR_NOMEM, ERR_KERNEL = range(6)
STAT_KING, STAT_IBS_RAW, STAT_IBS_PROPORTION, STAT_IBS_ADJUSTED, STAT_AS, STAT_GRM = range(6)
SRC_FILE, SRC_HOST, SRC_DEVICE = range(3)
SHARD_NONE, SHARD_LOCI, SHARD_TILES = range(3)
FLAG_PLANES_K_MAJOR, FLAG_NO_OVERLAP, FLAG_GRAM_1CTA, FLAG_NO_PACK, FLAG_ASYNC, FLAG_GRAM_I8 = 1, 2, 4, 8, 16, 32
MAT_DD, MAT_NN, MAT_HH, MAT_HN = range(4)
NUM_COUNT_MATS = 4
MXF4_MAX_CHAIN_LOCI = 8388608


class TpgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tpg_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class Job(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("stat", C.c_int32),
        ("src_kind", C.c_int32),
        ("device", C.c_int32),
        ("bk_path", C.c_char_p),
        ("bk_data", C.c_void_p),
        ("nrow_total", C.c_int64),
        ("ncol_total", C.c_int64),
        ("ind_row", C.c_void_p),
        ("n", C.c_int64),
        ("ind_col", C.c_void_p),
        ("m", C.c_int64),
        ("block_size", C.c_int64),
        ("pad_quirk", C.c_int32),
        ("shard_mode", C.c_int32),
        ("shard_rank", C.c_int32),
        ("shard_world", C.c_int32),
        ("gpu_chunk_loci", C.c_int64),
        ("flags", C.c_int32),
        ("reserved", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("ms_total", C.c_double),
        ("ms_read", C.c_double),
        ("ms_h2d", C.c_double),
        ("ms_decode", C.c_double),
        ("ms_gram", C.c_double),
        ("ms_finish", C.c_double),
        ("ms_d2h", C.c_double),
        ("bytes_h2d", C.c_int64),
        ("bytes_d2h", C.c_int64),
        ("n_chunks", C.c_int64),
        ("n_launches", C.c_int64),
        ("mma_macs", C.c_int64),
        ("n_short", C.c_int64),
        ("n_packed_chunks", C.c_int64),
        ("n_mxf4_chunks", C.c_int64),
        ("n_i8_chunks", C.c_int64),
        ("max_chain_loci", C.c_int64),
        ("bytes_file", C.c_int64),
        ("n_direct_io", C.c_int64),
        ("n_cached_jobs", C.c_int64),
        ("ms_reduce", C.c_double),
        ("n_geno_cached", C.c_int64),
        ("n_result_shells", C.c_int64),
        ("n_wide_launches", C.c_int64),
    ]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


class Result(C.Structure):
    _fields_ = [
        ("out0", C.c_void_p),
        ("out1", C.c_void_p),
        ("aux0", C.c_void_p),
        ("aux1", C.c_void_p),
        ("out_on_device", C.c_int32),
        ("reserved", C.c_int32),
        ("out0_path", C.c_char_p),
        ("out1_path", C.c_char_p),
        ("stats", Stats),
    ]


class Config(C.Structure):
    """tpg_config (include/tpg_b200.h): the library's tuning switches, one documented struct."""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("gram_route", C.c_int32),
        ("gram_ctas", C.c_int32),
        ("gram_ksplit", C.c_int32),
        ("gram_max_kblocks", C.c_int32),
        ("gram_clusters", C.c_int32),
        ("gram_no_trim", C.c_int32),
        ("gram_wide", C.c_int32),
        ("gram_l2_rows", C.c_int32),
        ("gram_l2_cols", C.c_int32),
        ("gram_order", C.c_int32),
        ("pack_mode", C.c_int32),
        ("pk2_i8", C.c_int32),
        ("no_compact", C.c_int32),
        ("gather_smem_kb", C.c_int32),
        ("gather_layout", C.c_int32),
        ("io_mode", C.c_int32),
        ("no_thp", C.c_int32),
        ("decode_k4_v1", C.c_int32),
        ("host_threads", C.c_int32),
        ("scan_loci_per_cta", C.c_int32),
        ("scan_no_indiv", C.c_int32),
        ("scan_table", C.c_int32),
        ("cache_counts", C.c_int32),
        ("cache_genotypes", C.c_int32),
        ("result_shells", C.c_int32),
        ("device_chunk_loci", C.c_int64),
        ("host_chunk_mb", C.c_int64),
        ("link_gbps", C.c_double),
        ("hybrid_factor", C.c_double),
    ]


# name -> (restype, argtypes); every symbol include/tpg_b200.h declares.
_ERR = [C.c_char_p, C.c_size_t]
SIGNATURES = {
    "tpg_version": (C.c_char_p, []),
    "tpg_device_count": (C.c_int, []),
    "tpg_block_size": (C.c_int64, [C.c_int64]),
    "tpg_split_len": (C.c_int64, [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "tpg_as_pad_count": (C.c_int64, [C.c_int64, C.c_int64]),
    "tpg_pk2_stride": (C.c_int64, [C.c_int64]),
    "tpg_pk2_pack_host": (None, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "tpg_config_init": (None, [C.POINTER(Config)]),
    "tpg_config_set": (C.c_int, [C.POINTER(Config)] + _ERR),
    "tpg_config_get": (None, [C.POINTER(Config)]),
    "tpg_pairwise": (C.c_int, [C.POINTER(Job), C.POINTER(Result)] + _ERR),
    "tpg_pairwise_many": (C.c_int, [C.POINTER(Job), C.c_void_p, C.c_int32, C.POINTER(Result), C.c_void_p, C.c_int32,
                                    C.c_int32] + _ERR),
    "tpg_pairwise_multi": (C.c_int, [C.POINTER(Job), C.POINTER(Result), C.c_void_p, C.c_int32, C.c_int32] + _ERR),
    "tpg_pairwise_ctx": (C.c_int, [C.c_void_p, C.POINTER(Job), C.POINTER(Result)] + _ERR),
    "tpg_grm_from_as": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_double)] + _ERR),
    "tpg_filter_high_relatedness": (C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_void_p,
                                              C.c_void_p, C.POINTER(C.c_int64)] + _ERR),
    "tpg_king_filter": (C.c_int, [C.POINTER(Job), C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p,
                                  C.POINTER(Stats)] + _ERR),
    "tpg_king_filter_groups": (C.c_int, [C.POINTER(Job), C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(Stats)] + _ERR),
    "tpg_as_wg17": (C.c_int, [C.POINTER(Job), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.POINTER(Stats)] + _ERR),
    "tpg_wg17": (C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p] + _ERR),
    "tpg_scan": (C.c_int, [C.POINTER(Job), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p, C.POINTER(Stats)] + _ERR),
    "tpg_ctx_create": (C.c_void_p, [C.c_int] + _ERR),
    "tpg_ctx_destroy": (None, [C.c_void_p]),
    "tpg_ctx_trim": (None, [C.c_void_p]),
    "tpg_release": (None, [C.c_int]),
    "tpg_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "tpg_tile_owner": (C.c_int32, [C.c_int64, C.c_int64, C.c_int64, C.c_int32]),
    "tpg_result_shell_bounds": (C.c_int32, [C.c_int64, C.c_int32, C.c_void_p]),
    "tpg_counts_ld": (C.c_int64, [C.c_int64]),
    "tpg_counts_elems": (C.c_int64, [C.c_int64]),
    "tpg_accumulate": (C.c_int, [C.c_void_p, C.POINTER(Job), C.c_void_p, C.POINTER(Stats)] + _ERR),
    "tpg_finish": (C.c_int, [C.c_void_p, C.POINTER(Job), C.c_void_p, C.POINTER(Result)] + _ERR),
    "tpg_decode_planes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                    C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_int32] + _ERR),
    "tpg_gram": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                           C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_float)] + _ERR),
    "tpg_mma_peak": (C.c_int, [C.c_int, C.c_int32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)] + _ERR),
    "tpg_synth_fill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64,
                                 C.c_double, C.c_int32] + _ERR),
    "tpg_synth_fill_host": (None, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_double,
                                   C.c_int32]),
}

_lib = None


def build(verbose: bool = False) -> Path:
    """Compile libtpg_b200.so for sm_100a with nvcc (in-tree).  Raises on failure."""
    cmd = ["make", "-C", str(CSRC), "-j4"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0 or not LIB_PATH.exists():
        raise RuntimeError("building libtpg_b200.so failed (nvcc -gencode arch=compute_100a,code=sm_100a)")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("TPG_B200_AUTOBUILD", "1") == "1":
            build()
        else:
            raise ImportError(f"{LIB_PATH} is missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError = the .so does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def config(**fields) -> dict:
    """Reads (no arguments) or updates the process-wide tpg_config — the counterpart of the R glue's
    options(tidypopgen.gpu.<field>).  config(reset=True) restores the built-in defaults (+ TPG_* env).
    Returns the configuration in force as a dict."""
    L = lib()
    c = Config()
    if fields.pop("reset", False):
        L.tpg_config_init(C.byref(c))
    else:
        L.tpg_config_get(C.byref(c))
    names = {k for k, _ in Config._fields_}
    for k, v in fields.items():
        if k not in names or k == "abi_version":
            raise KeyError(f"tpg_config has no field '{k}'")
        setattr(c, k, v)
    eb = errbuf()
    check(L.tpg_config_set(C.byref(c), eb, len(eb)), eb)
    return {k: getattr(c, k) for k, _ in Config._fields_}


def check(rc: int, errbuf) -> None:
    if rc != OK:
        raise TpgError(rc, errbuf.value.decode(errors="replace"))


def errbuf():
    return C.create_string_buffer(1024)


class Context:
    """Per-device handle for the staged / kernel-level interface."""

    def __init__(self, device: int = 0):
        eb = errbuf()
        self.device = device
        self.handle = lib().tpg_ctx_create(device, eb, len(eb))
        if not self.handle:
            raise TpgError(ERR_CUDA, eb.value.decode(errors="replace"))

    def close(self):
        if getattr(self, "handle", None):
            lib().tpg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
