"""Builds hifimeth_b200/libhm_engine.so (the C-ABI library) in-tree with nvcc for sm_100a only."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libhm_engine.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["engine.cu", "cnn_tensor.cu", "onnx_weights.cpp", "host_record.cpp", "bgzf_bam.cpp", "fast_deflate.cpp", "call_main.cpp"]
EXE = PKG / "bin" / "hifimeth-b200"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "hm_engine.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    objs = []
    build_dir = PKG / "build"
    build_dir.mkdir(exist_ok=True)
    log = []
    for src in SOURCES:
        obj = build_dir / (src + ".o")
        cmd = [NVCC, *FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(str(obj))
    # Link with the host compiler and the static CUDA runtime: nvcc's own link step would embed a default
    # sm_52 device-link stub, and a shared libcudart could collide with the one torch bundles.
    cuda_lib = str(Path(NVCC).resolve().parent.parent / "lib64")
    r = subprocess.run(["g++", "-shared", "-o", str(LIB), *objs, "-L" + cuda_lib, "-lcudart_static", "-ldl", "-lrt", "-lpthread", "-lz"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    # the CLI: main.cpp against the shared library (rpath = the package directory)
    EXE.parent.mkdir(exist_ok=True)
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(EXE), str(CSRC / "main.cpp"), "-L" + str(PKG), "-lhm_engine",
                        "-Wl,-rpath,$ORIGIN/.."], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of the CLI failed")
    (build_dir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
