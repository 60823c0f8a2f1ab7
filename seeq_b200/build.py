"""Builds the in-tree native artefacts (nvcc for sm_100a + gcc).

  seeq_b200/libseeq_b200.so   the product: CUDA kernels + C-ABI (sqb*), the
                              libseeq API (seeq*), the file driver and seeq()
  seeq_b200/_relink/seeq      the reference CLI (seeq-main.c, compiled from the
                              read-only mount, not copied) linked against it
  seeq_b200/_relink/seeq*.so  the reference CPython module (seeqmodule.c)
                              linked against it

The last two exist only where /root/reference is mounted; they travel to the
GPU box as prebuilt files.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libseeq_b200.so")
RELINK = os.path.join(HERE, "_relink")
REFERENCE_DIR = os.environ.get("SEEQ_REFERENCE_DIR", "/root/reference")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

# (source, extra defines, object name): the matcher instances are spread over translation units that compile side by side
CU_UNITS = [("sqb_engine.cu", [], "sqb_engine.cu.o"),
            ("sqb_engine_wm.cu", ["-DSQB_WM_FUSED=0"], "sqb_engine_wm.cu.o"),
            ("sqb_engine_wm.cu", ["-DSQB_WM_FUSED=1"], "sqb_engine_wmf.cu.o"),
            ("sqb_engine_bsf.cu", [], "sqb_engine_bsf.cu.o"),
            ("sqb_bgzf.cu", [], "sqb_bgzf.cu.o")]
C_SOURCES = ["seeq_api.c", "seeq_file.c"]
HEADERS = sorted(os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))) + \
          [os.path.join(INC, h) for h in ("libseeq.h", "seeq.h", "seeq_b200.h")]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd) -> None:
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout)
        raise RuntimeError("build step failed: " + cmd[0])


def build_library(force: bool = False, verbose: bool = False) -> str:
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    jobs = []
    for src, defs, oname in CU_UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, oname)
        if force or _newer(o, [s] + HEADERS):
            jobs.append([NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp", *defs,
                         "-I" + INC, "-I" + CSRC, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else []))
        objs.append(o)
    if jobs:                                  # the translation units compile side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            list(pool.map(_run, jobs))
    for src in C_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src + ".o")
        if force or _newer(o, [s] + HEADERS):
            _run(["gcc", "-std=c99", "-O2", "-fPIC", "-Wall", "-Wextra", "-I" + INC, "-I" + CSRC,
                  "-c", s, "-o", o])
        objs.append(o)
    if force or _newer(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fopenmp", "-lgomp"])
    return LIB


def build_variant(tag: str, defines, verbose: bool = False) -> str:
    """An A/B build of the library with extra -D flags -> seeq_b200/libseeq_b200_<tag>.so (git-ignored; travels
    to the GPU box; selected with SEEQ_B200_LIB, see tools/gpu_session.sh `ab`)."""
    objdir = os.path.join(HERE, "build_" + tag)
    os.makedirs(objdir, exist_ok=True)
    lib = os.path.join(HERE, "libseeq_b200_%s.so" % tag)
    objs, jobs = [], []
    for src, defs, oname in CU_UNITS:
        o = os.path.join(objdir, oname)
        jobs.append([NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-fopenmp", *defs, *defines,
                     "-I" + INC, "-I" + CSRC, "-c", os.path.join(CSRC, src), "-o", o] + (["-Xptxas", "-v"] if verbose else []))
        objs.append(o)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        list(pool.map(_run, jobs))
    for src in C_SOURCES:
        o = os.path.join(objdir, src + ".o")
        _run(["gcc", "-std=c99", "-O2", "-fPIC", "-Wall", "-Wextra", "-I" + INC, "-I" + CSRC, "-c",
              os.path.join(CSRC, src), "-o", o])
        objs.append(o)
    _run([NVCC, *ARCH, "-shared", "-o", lib, *objs, "-Xcompiler", "-fopenmp", "-lgomp"])
    return lib


def build_relinks(force: bool = False) -> dict:
    """Re-link the reference front-ends against our library (no source copied)."""
    out = {}
    src = os.path.join(REFERENCE_DIR, "src")
    if not os.path.isdir(src):
        return out
    os.makedirs(RELINK, exist_ok=True)
    rpath = "-Wl,-rpath,$ORIGIN/.."
    cli = os.path.join(RELINK, "seeq")
    main_c = os.path.join(src, "seeq-main.c")
    if force or _newer(cli, [main_c, LIB]):
        # -ftrivial-auto-var-init=zero: seeq-main.c never assigns args.split
        _run(["gcc", "-std=gnu99", "-O2", "-w", "-ftrivial-auto-var-init=zero", "-I" + INC,
              main_c, "-o", cli, "-L" + HERE, "-lseeq_b200", rpath])
    out["cli"] = cli
    # the CPython module: seeqmodule.c is compiled where it lies (included by path from our wrapper
    # translation unit, which adds the batched method SeeqObject.matchBatch)
    mod_c = os.path.join(src, "seeqmodule.c")
    wrap_c = os.path.join(CSRC, "seeqmodule_b200.c")
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    mod = os.path.join(RELINK, "seeq" + ext)
    if force or _newer(mod, [mod_c, wrap_c, LIB]):
        _run(["gcc", "-std=gnu99", "-O2", "-w", "-fPIC", "-shared", "-DMAJOR_VERSION=1", "-DMINOR_VERSION=2",
              "-DSQB_REFERENCE_MODULE=\"" + mod_c + "\"",
              "-I" + INC, "-I" + sysconfig.get_paths()["include"], wrap_c, "-o", mod,
              "-L" + HERE, "-lseeq_b200", rpath])
    out["module"] = mod
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_library(force, verbose)
    build_relinks(force)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "variant":      # python seeq_b200/build.py variant TAG -DX=1 ...
        print(build_variant(sys.argv[2], sys.argv[3:]))
        sys.exit(0)
    build_all(force="-B" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
