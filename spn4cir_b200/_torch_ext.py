"""Build + load of ``_spn_torch.so``: the PyTorch extension (``csrc_torch/spn_torch.cpp``) that registers the
``torch.ops.spn4cir.*`` operators and their C++ autograd nodes over the C ABI of ``libspn_b200.so``.

Plain C++ (no device code): compiled IN-TREE with g++ against the torch headers and linked to ``libspn_b200.so``
(``$ORIGIN`` rpath), so both shared objects travel with the source tree.  There is no fallback: if the extension cannot
be built or loaded, the operators raise.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import threading
from pathlib import Path

from ._lib import CSRC_DIR, HEADER_PATH, LIB_PATH, PKG_DIR, REPO_ROOT, SpnError, build_library, build_lock

EXT_PATH = PKG_DIR / "_spn_torch.so"
EXT_SRC = PKG_DIR / "csrc_torch" / "spn_torch.cpp"

_loaded = False
_lock = threading.Lock()


def _stale() -> bool:
    if not EXT_PATH.exists():
        return True
    t = EXT_PATH.stat().st_mtime
    return EXT_SRC.stat().st_mtime > t or HEADER_PATH.stat().st_mtime > t


def build_torch_extension(force: bool = False, verbose: bool = False) -> Path:
    """g++ -shared csrc_torch/spn_torch.cpp -> spn4cir_b200/_spn_torch.so (works without a GPU)."""
    if not force and not _stale():
        return EXT_PATH
    build_library()  # the extension links against libspn_b200.so
    with build_lock():
        if not force and not _stale():  # another process built it while this one waited for the lock
            return EXT_PATH
        return _build_locked(verbose)


def _compilers():
    """g++ candidates, the system one first.  $CXX may point at a toolchain whose only libstdc++ is the static archive
    (this image: /opt/gcc): an extension linked against it carries a second, never-initialised copy of the iostream
    machinery next to the libstdc++.so torch uses -- formatting a number into an error message then crashes the process
    instead of raising."""
    seen, out = set(), []
    for c in (shutil.which("g++"), "/usr/bin/g++", os.environ.get("CXX"), shutil.which("c++")):
        if c and os.path.exists(c) and os.path.realpath(c) not in seen:
            seen.add(os.path.realpath(c))
            out.append(c)
    return out


def _embeds_libstdcxx(so: Path) -> bool:
    """True when the shared object DEFINES ostream symbols itself (libstdc++ linked statically into it)."""
    nm = shutil.which("nm")
    if not nm:
        return False
    res = subprocess.run([nm, "-D", "--defined-only", str(so)], capture_output=True, text=True)
    return res.returncode == 0 and "_ZNSo9_M_insertIlEERSoT_" in res.stdout


def _build_locked(verbose: bool) -> Path:
    import torch
    from torch.utils import cpp_extension as ce
    compilers = _compilers()
    if not compilers:
        raise SpnError("build", -5, "g++ not found: the torch extension cannot be built")
    cuda_home = os.environ.get("CUDA_HOME") or "/usr/local/cuda"
    incs = list(dict.fromkeys(ce.include_paths() + [str(Path(cuda_home) / "include")]))
    torch_lib = str(Path(torch.__file__).resolve().parent / "lib")
    tmp = EXT_PATH.with_suffix(".so.tmp%d" % os.getpid())
    errors = []
    for cxx in compilers:
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-deprecated-declarations",
               f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_EXTENSION_NAME=_spn_torch",
               *[f"-I{i}" for i in incs], str(EXT_SRC), "-o", str(tmp),
               f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
               f"-L{PKG_DIR}", f"-l:{LIB_PATH.name}", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            tmp.unlink(missing_ok=True)
            errors.append(f"{cxx} failed on spn_torch.cpp:\n" + (res.stdout + res.stderr)[-3000:])
            continue
        if _embeds_libstdcxx(tmp):
            tmp.unlink(missing_ok=True)
            errors.append(f"{cxx} links libstdc++ statically into the extension (a second copy next to torch's libstdc++.so)")
            continue
        os.replace(tmp, EXT_PATH)
        if verbose:
            print(f"[{cxx}]\n" + res.stdout + res.stderr)
        return EXT_PATH
    raise SpnError("build", -5, "no usable g++ for the torch extension:\n" + "\n".join(errors))


def load_torch_extension() -> None:
    """Register torch.ops.spn4cir.* (idempotent); builds the extension first when it is absent or stale."""
    global _loaded
    with _lock:
        if _loaded:
            return
        import torch
        if _stale():
            build_torch_extension()
        if not EXT_PATH.exists():
            raise SpnError("load", -5, f"{EXT_PATH} is missing: run __graft_entry__.build()")
        torch.ops.load_library(str(EXT_PATH))
        _loaded = True


def ops():
    """``torch.ops.spn4cir`` with the extension loaded."""
    import torch
    load_torch_extension()
    return torch.ops.spn4cir
