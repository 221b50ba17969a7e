"""Builds the C-ABI library of another git revision (or of the working tree with extra -D flags) into ab/<name>.so, for A/B runs of
two builds on ONE GPU box (boxes of the pool differ by +-3 % under the power cap): HM_ENGINE_LIB=ab/<name>.so python bench.py ...

    python tools/build_variant.py <name> [--rev REV] [-DFLAG ...]
"""
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hifimeth_b200 import build as b  # noqa: E402


def main():
    name = sys.argv[1]
    rev = None
    defs = []
    args = sys.argv[2:]
    while args:
        a = args.pop(0)
        if a == "--rev":
            rev = args.pop(0)
        else:
            defs.append(a)
    out = ROOT / "ab"
    out.mkdir(exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        if rev:
            subprocess.run(f"git -C {ROOT} archive {rev} hifimeth_b200/csrc include | tar -x -C {tmp}", shell=True, check=True)
            csrc = tmp / "hifimeth_b200" / "csrc"
        else:
            csrc = b.CSRC
        objs = []
        for src in b.SOURCES:
            obj = tmp / (src + ".o")
            subprocess.run([b.NVCC, *[f for f in b.FLAGS if f not in ("-Xptxas", "-v")], *defs, "-c", str(csrc / src), "-o", str(obj)], check=True)
            objs.append(str(obj))
        cuda_lib = str(Path(b.NVCC).resolve().parent.parent / "lib64")
        subprocess.run(["g++", "-shared", "-o", str(out / f"{name}.so"), *objs, "-L" + cuda_lib, "-lcudart_static", "-ldl", "-lrt", "-lpthread", "-lz"], check=True)
    print(out / f"{name}.so")


if __name__ == "__main__":
    main()
