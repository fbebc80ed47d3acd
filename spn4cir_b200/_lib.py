"""Build + ctypes binding of ``libspn_b200.so`` (the C ABI declared in ``include/spn_b200.h``).

The library is built IN-TREE with nvcc for sm_100a only and loaded with ctypes; no torch types cross
the boundary (device pointers, sizes and a raw ``cudaStream_t``).  There is no fallback: if the
library is missing or a call fails, :class:`SpnError` is raised.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC_DIR = PKG_DIR / "csrc"
REPO_ROOT = PKG_DIR.parent
LIB_PATH = PKG_DIR / "libspn_b200.so"
HEADER_PATH = REPO_ROOT / "include" / "spn_b200.h"
DEBUG_HEADER_PATH = REPO_ROOT / "include" / "spn_b200_debug.h"  # bring-up / test-only entry points

SOURCES = ["common.cu", "bank_ce_tc.cu", "shard.cu", "rank_window.cu", "part_ce.cu", "umma_selftest.cu", "prologue.cu", "sim_rank.cu", "sim_tc.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177",
]

# status codes (include/spn_b200.h)
SPN_OK = 0
STATUS_NAMES = {
    0: "SPN_OK", -1: "SPN_ERR_BAD_SHAPE", -2: "SPN_ERR_BAD_ALIGN", -3: "SPN_ERR_BAD_DTYPE",
    -4: "SPN_ERR_WORKSPACE", -5: "SPN_ERR_CUDA", -6: "SPN_ERR_NCCL", -7: "SPN_ERR_BAD_ARG",
}
OP_BANK_CE_FWD, OP_BANK_CE_BWD, OP_SIM_TOPK, OP_SIM_RANK, OP_BANK_CE_STEP, OP_BANK_CE_FUSED = 1, 2, 3, 4, 5, 6
OP_SIM_TC = 7
DT_F32, DT_F16, DT_BF16 = 0, 1, 2


class SpnError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        self.code = code
        super().__init__(f"{fn} -> {STATUS_NAMES.get(code, code)}: {msg}")


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise SpnError("build", -5, "nvcc not found; libspn_b200.so must be built with CUDA >= 12.9")
    return cand


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC_DIR.glob("*.cu")) + list(CSRC_DIR.glob("*.cuh")) + [HEADER_PATH, DEBUG_HEADER_PATH]
    return any(p.stat().st_mtime > t for p in deps)


class build_lock:
    """Exclusive inter-process lock around a build (the ranks of a torchrun job may all find the library stale at once;
    they must not write the same object files concurrently: the first one builds, the others wait and find it fresh)."""

    def __enter__(self):
        import fcntl
        (REPO_ROOT / "build").mkdir(parents=True, exist_ok=True)
        self.f = open(REPO_ROOT / "build" / ".lock", "w")
        fcntl.flock(self.f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self.f, fcntl.LOCK_UN)
        self.f.close()
        return False


def build_library(force: bool = False, verbose: bool = False, extra_flags=(), out: Path = None) -> Path:
    """Compile csrc/*.cu into spn4cir_b200/libspn_b200.so (sm_100a). Cross-compiles without a GPU.
    One nvcc -c per source (in parallel, objects cached under build/ by mtime), then one link."""
    out = Path(out) if out is not None else LIB_PATH
    if out == LIB_PATH and not force and not _stale():
        return LIB_PATH
    with build_lock():
        if out == LIB_PATH and not force and not _stale():  # another process built it while this one waited
            return LIB_PATH
        return _build_library_locked(force, verbose, extra_flags, out)


def _build_library_locked(force: bool, verbose: bool, extra_flags, out: Path) -> Path:
    from concurrent.futures import ThreadPoolExecutor
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra_flags)
    tag = "obj" if not extra_flags else "obj_" + "".join(c if c.isalnum() else "_" for c in "".join(extra_flags))
    objdir = REPO_ROOT / "build" / tag
    objdir.mkdir(parents=True, exist_ok=True)
    hdr_t = max(p.stat().st_mtime for p in list(CSRC_DIR.glob("*.cuh")) + [HEADER_PATH, DEBUG_HEADER_PATH])

    def compile_one(src: str):
        s, o = CSRC_DIR / src, objdir / (Path(src).stem + ".o")
        if not force and o.exists() and o.stat().st_mtime > max(s.stat().st_mtime, hdr_t):
            return o, ""
        cmd = [nvcc, *flags, *(["-Xptxas", "-v"] if verbose else []), "-c", str(s), "-o", str(o)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            o.unlink(missing_ok=True)
            raise SpnError("build", -5, f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
        return o, res.stdout + res.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        done = list(ex.map(compile_one, SOURCES))
    tmp = out.with_suffix(".so.tmp%d" % os.getpid())
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(tmp),
                          *[str(o) for o, _ in done]], capture_output=True, text=True)
    if res.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise SpnError("build", -5, "link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, out)
    if verbose:
        print("".join(log for _, log in done))
    return out


_f32p = POINTER(c_float)

# name -> (argtypes); every function returns int except spn_last_error
_SIGNATURES = {
    "spn_version": [],
    "spn_device_supported": [c_int],
    "spn_workspace_bytes": [c_int, c_int, c_int64, c_int, c_int, c_int, POINTER(c_size_t)],
    "spn_cast_to_bf16": [c_void_p, c_int, c_void_p, c_int64, c_void_p],
    "spn_query_assemble": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p],
    "spn_query_assemble_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p],
    "spn_bank_ce_fwd": [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int64, c_int, c_int, c_int64,
                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_merge_stats": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_void_p],
    "spn_bank_ce_loss": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_void_p, c_void_p,
                         c_void_p, c_void_p],
    "spn_bank_ce_bwd": [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_float, c_void_p, c_float, c_int64,
                        c_int, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_bank_stats": [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p],
    "spn_bank_ce_prep": [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int64, c_int, c_int, c_int64,
                         c_void_p, c_void_p, c_void_p],
    "spn_bank_ce_fused": [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int64, c_int, c_int, c_int64, c_int,
                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_bank_ce_finalize": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                             c_float, c_float, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_void_p],
    "spn_shard_pack": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "spn_shard_fold": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p],
    "spn_shard_home_prologue": [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p],
    "spn_peer_barrier": [c_void_p, c_int, c_int, c_uint32, c_void_p],
    "spn_shard_pull": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "spn_bank_ce_fused_peers": [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int64, c_int, c_int, c_int64, c_int,
                                c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_bank_ce_step": [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_float, c_int, c_int64, c_int,
                         c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_part_ce": [c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                    c_void_p, c_void_p],
    "spn_sim_topk": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int32, c_void_p,
                     c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_sim_topk_after": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int32, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_merge_topk": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "spn_sim_target_score": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int32, c_void_p, c_void_p],
    "spn_sim_rank": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                     c_int, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_sim_topk_tc": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int32, c_void_p,
                        c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_sim_rank_tc": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                        c_int, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p],
    "spn_debug_sim_tc_layout": [c_int, c_int, c_int, c_int, POINTER(c_size_t), POINTER(c_int), POINTER(c_int)],
    "spn_rank_window": [c_void_p, c_int, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "spn_sim_matrix": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "spn_debug_read_timeline": [c_void_p, c_int],
    "spn_debug_umma_bench": [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "spn_debug_dsmem_bench": [c_int, c_int, c_void_p, c_int, c_void_p],
    "spn_debug_umma_gemm": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["spn_last_error"])

_lib = None
_lock = threading.Lock()


def load_library(build_if_missing: bool = True) -> ctypes.CDLL:
    """Return the loaded CDLL with argtypes set; build it first when the .so is absent/stale."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        override = os.environ.get("SPN_LIB_PATH")  # debug builds (e.g. -DSPN_TIMELINE) used by tools/
        if override:
            lib = ctypes.CDLL(override)
            for name, args in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = args
                fn.restype = c_int
            lib.spn_last_error.argtypes = []
            lib.spn_last_error.restype = c_char_p
            _lib = lib
            return lib
        if build_if_missing and _stale():
            if shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists():
                build_library()
        if not LIB_PATH.exists():
            raise SpnError("load", -5, f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, args in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = c_int
        lib.spn_last_error.argtypes = []
        lib.spn_last_error.restype = c_char_p
        _lib = lib
        return lib


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point; raise SpnError with spn_last_error() on failure."""
    lib = load_library()
    rc = getattr(lib, name)(*args)
    if rc != SPN_OK:
        raise SpnError(name, rc, lib.spn_last_error().decode("utf-8", "replace"))


def workspace_bytes(op: int, B: int, N: int, D: int, T: int = 1, K: int = 0) -> int:
    out = c_size_t(0)
    call("spn_workspace_bytes", op, B, N, D, T, K, ctypes.byref(out))
    return int(out.value)
