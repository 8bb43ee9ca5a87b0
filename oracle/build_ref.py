"""Materialise ``oracle/_ref/`` -- the UNMODIFIED reference modules of the hot path, importable on the GPU box.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container and not on the GPU box; the reference is pure
Python (nothing to compile), so "building" it means packing the packages the path needs, byte for byte, into one
build artefact that a ``gpurun`` snapshot carries (the Python analogue of the ``.so`` a compiled reference would yield):

    oracle/_ref/reference_modules.zip   members models/** and transport/** of /root/reference (model.py, math.py, sampling.py,
                                        modules/*, transport/*), imported straight from the archive (zipimport)
    oracle/_ref/MANIFEST.json           sha256 of every member (``verify()`` re-hashes them: they must stay unmodified)

plus two third-party stand-ins that are absent from the image (written by THIS script, not reference code):

    oracle/_ref/imwatermark.py   empty ``WatermarkEncoder`` (models/util.py:7 imports it; never used on the path)
    oracle/_ref/torchdiffeq.py   fixed-grid Euler ``odeint`` with upstream semantics (see the docstring in the stub)

``oracle/_ref/`` is git-ignored (no reference source enters the history) and NOT gpurun-ignored (it travels to the GPU box, like
the built .so).  ``__graft_entry__.build()`` calls ``build()`` when /root/reference is present.  Consumers: ``oracle/ref_runner.py``
(-m gpu parity tests, bench.py --impl reference).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(DST, "reference_modules.zip")
PACKAGES = ("models", "transport")

_IMWATERMARK = '''"""Stand-in for the absent third-party `imwatermark` (written by oracle/build_ref.py; models/util.py:7 imports it)."""


class WatermarkEncoder:
    def set_watermark(self, *a, **k):
        return None

    def encode(self, img, *a, **k):
        return img
'''

_TORCHDIFFEQ = '''"""Stand-in for the absent third-party `torchdiffeq` (written by oracle/build_ref.py; transport/integrators.py:2 imports it).

Restates the upstream fixed-grid solver for method="euler" (torchdiffeq/_impl/{odeint,solvers,fixed_grid,misc}.py, unpinned in
the reference's requirements.txt):
  * the grid is `t` itself; `solution[0] = y0`; `y1 = y0 + dt * f(t0, y0)` with `dt = t1 - t0` taken from the un-rounded grid;
  * the ODE function is wrapped in `_PerturbFunc`, whose forward casts the time to the state dtype: `t = t.to(y.abs().dtype)`;
  * `t` is moved to `y0.device`; the stacked trajectory has y0's dtype.
"""
import torch


def odeint(func, y0, t, *, rtol=1e-7, atol=1e-9, method=None, options=None, event_fn=None):
    if method != "euler":
        raise NotImplementedError("the stand-in implements the fixed-grid euler solver only (the pipeline default)")
    if isinstance(y0, tuple):
        raise NotImplementedError("tuple states are not used on the inference path")
    t = t.to(y0.device)
    solution = torch.empty(len(t), *y0.shape, dtype=y0.dtype, device=y0.device)
    solution[0] = y0
    y = y0
    for j in range(1, len(t)):
        t0, t1 = t[j - 1], t[j]
        dt = t1 - t0
        f0 = func(t0.to(y.abs().dtype), y)
        y = y + dt * f0
        solution[j] = y
    return solution
'''


def _sha_bytes(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is in place (freshly packed or already present), False when there is no reference here."""
    if not os.path.isdir(REF_SRC):
        return os.path.exists(os.path.join(DST, "MANIFEST.json")) and os.path.exists(ARCHIVE)
    if os.path.isdir(DST) and not force:
        try:
            if verify(against_source=True):
                return True
        except Exception:  # noqa: BLE001  (stale or partial artefact: rebuild)
            pass
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    manifest = {}
    with zipfile.ZipFile(ARCHIVE, "w", compression=zipfile.ZIP_DEFLATED) as z:
        for pkg in PACKAGES:
            for root, dirs, files in os.walk(os.path.join(REF_SRC, pkg)):
                dirs[:] = sorted(d for d in dirs if d != "__pycache__")
                # explicit directory members: models/modules has no __init__.py and zipimport finds namespace packages by them
                z.writestr(zipfile.ZipInfo(os.path.relpath(root, REF_SRC) + "/", date_time=(2020, 1, 1, 0, 0, 0)), b"")
                for fn in sorted(files):
                    if not fn.endswith(".py"):
                        continue
                    src = os.path.join(root, fn)
                    rel = os.path.relpath(src, REF_SRC)
                    data = open(src, "rb").read()
                    z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), data)
                    manifest[rel] = _sha_bytes(data)
    with open(os.path.join(DST, "imwatermark.py"), "w") as f:
        f.write(_IMWATERMARK)
    with open(os.path.join(DST, "torchdiffeq.py"), "w") as f:
        f.write(_TORCHDIFFEQ)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "archive": os.path.basename(ARCHIVE), "files": manifest, "stand_ins": ["imwatermark.py", "torchdiffeq.py"]},
                  f, indent=1, sort_keys=True)
    return True


def verify(against_source: bool = False) -> bool:
    """Every archive member still has the hash recorded at packing time (and, in the build container, equals the source)."""
    man = json.load(open(os.path.join(DST, "MANIFEST.json")))
    with zipfile.ZipFile(ARCHIVE) as z:
        names = set(z.namelist())
        for rel, sha in man["files"].items():
            if rel not in names or _sha_bytes(z.read(rel)) != sha:
                raise RuntimeError(f"oracle/_ref: member {rel} of the archive does not match the manifest")
            if against_source and os.path.isdir(REF_SRC) and _sha_bytes(open(os.path.join(REF_SRC, rel), "rb").read()) != sha:
                raise RuntimeError(f"oracle/_ref: member {rel} differs from {REF_SRC}/{rel}")
    return True


if __name__ == "__main__":
    ok = build(force=True)
    print("oracle/_ref ready" if ok else "no /root/reference here and no previous copy")
