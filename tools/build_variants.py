"""Diagnostics: build experimental variants of libsicp_b200.so that differ in -D switches of one
translation unit, for A/B timing in a single GPU session:

    python tools/build_variants.py nn.cu tagA:-DSICP_MATCH_C4=1 tagB:-DSICP_MATCH_RECPOS=0 ...
    SICP_B200_LIB=simpleicp_b200/_variants/libsicp_tagA.so python tools/sort_probe.py

The other objects are reused from simpleicp_b200/_build (run the normal build first)."""
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from simpleicp_b200 import _build  # noqa: E402


def main():
    src = sys.argv[1]
    _build.build()
    out = _build.PKG / "_variants"
    out.mkdir(exist_ok=True)
    nvcc = _build._nvcc()
    for spec in sys.argv[2:]:
        tag, _, defs = spec.partition(":")
        obj = out / f"{Path(src).stem}_{tag}.o"
        cmd = [nvcc, *_build.ARCH, *_build.FLAGS, *[d for d in defs.split(",") if d], "-c", str(_build.CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise SystemExit(1)
        objs = [str(obj)] + [str(_build.OBJ / (s.rsplit(".", 1)[0] + ".o")) for s in _build.SOURCES if s != src]
        lib = out / f"libsicp_{tag}.so"
        r = subprocess.run([nvcc, *_build.ARCH, "-shared", "-o", str(lib), *objs], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise SystemExit(1)
        print(lib)


if __name__ == "__main__":
    main()
