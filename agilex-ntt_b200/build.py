"""Build recipe for the in-tree native code (sm_100a only).

`build_all()` is what __graft_entry__.build() calls: nvcc cross-compiles here without a GPU; the resulting
lib/*.so and bin/* travel to the GPU box with the repo snapshot (they are git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIBDIR, "libagxntt.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _sources(d: str, exts=(".cu", ".cuh", ".cpp", ".h", ".hpp")) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout[-4000:], r.stderr[-4000:]))


def source_hash() -> str:
    """SHA-256 over the library's sources (names and contents): what `lib/libagxntt.so.srchash` records at build time."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(_sources(CSRC) + [os.path.join(ROOT, "include", "agxntt.h")]):
        h.update(os.path.relpath(f, ROOT).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def recorded_hash() -> str | None:
    try:
        with open(LIB + ".srchash") as fh:
            return fh.read().strip()
    except OSError:
        return None


def build_lib(force: bool = False) -> str:
    deps = _sources(CSRC) + [os.path.join(ROOT, "include", "agxntt.h")]
    want = source_hash()
    if force or _stale(LIB, deps) or recorded_hash() != want:      # mtimes lie after a checkout; the hash does not
        os.makedirs(LIBDIR, exist_ok=True)
        _run([nvcc()] + NVCC_FLAGS + ["-shared", "-o", LIB,
                                      os.path.join(CSRC, "agx_api.cu"), os.path.join(CSRC, "agx_tables.cpp")])
        with open(LIB + ".srchash", "w") as fh:
            fh.write(want + "\n")
    return LIB


def build_tools(force: bool = False) -> dict[str, str]:
    """Stand-alone binaries: the main.cpp-shaped compat driver, the reference's own main.cpp, the C99 example."""
    os.makedirs(BINDIR, exist_ok=True)
    out = {}
    drv_src = os.path.join(HOST, "main_compat.cpp")
    drv = os.path.join(BINDIR, "agx_main_compat")
    if os.path.exists(drv_src):
        lib = build_lib()
        if force or _stale(drv, _sources(HOST) + [lib]):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            _run([cxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", HOST,
                  "-o", drv, drv_src, os.path.join(HOST, "ntt_shim.cpp"),
                  "-L", LIBDIR, "-lagxntt", "-Wl,-rpath,$ORIGIN/../lib"])
        out["main_compat"] = drv
        # The reference's own, UNMODIFIED src/main.cpp built against host/ (only where the reference tree exists;
        # the binary -- not the source -- travels to the GPU box).  Proves the drop-in claim literally.
        ref_main = "/root/reference/src/main.cpp"
        ref_bin = os.path.join(BINDIR, "agx_ref_main")
        if os.path.exists(ref_main) and (force or _stale(ref_bin, _sources(HOST) + [lib, ref_main])):
            cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            _run([cxx, "-O2", "-std=c++17", "-w", "-I", os.path.join(ROOT, "include"), "-I", HOST,
                  "-o", ref_bin, ref_main, os.path.join(HOST, "ntt_shim.cpp"),
                  "-L", LIBDIR, "-lagxntt", "-Wl,-rpath,$ORIGIN/../lib"])
        if os.path.exists(ref_bin):
            out["ref_main"] = ref_bin
        # the header and library from plain C99
        c_src, c_bin = os.path.join(HOST, "c_example.c"), os.path.join(BINDIR, "agx_c_example")
        if os.path.exists(c_src) and (force or _stale(c_bin, [c_src, lib, os.path.join(ROOT, "include", "agxntt.h")])):
            cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
            _run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-O2", "-I", os.path.join(ROOT, "include"),
                  "-o", c_bin, c_src, "-L", LIBDIR, "-lagxntt", "-Wl,-rpath,$ORIGIN/../lib"])
        if os.path.exists(c_bin):
            out["c_example"] = c_bin
    return out


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_tools(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)
    print("built", LIB)
