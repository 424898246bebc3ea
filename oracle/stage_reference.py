"""Install the UNMODIFIED reference into baseline/_ref/ (TEST / BASELINE INFRASTRUCTURE ONLY -- the product package
never imports it) and put it on sys.path for the reference arm of bench.py.

/root/reference exists only in the build container; the GPU box sees what travels with the repo snapshot.
baseline/_ref/ is git-ignored (no reference source ever enters the history) but NOT gpurun-ignored, so the installed
package travels like the built .so.  The recipe is the one offline install the task statement allows:

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference under /tmp>

(--no-deps: omegaconf / kornia / h5py ... are not in the wheelhouse; the matcher's import closure -- 13 pure-Python
files -- needs only torch, numpy and omegaconf, and omegaconf is covered by the ~100-line stand-in oracle/_shim/omegaconf;
the copy under /tmp is needed because the wheel build writes into the source tree and /root/reference is read-only).

`bench.py --impl reference` and its `gpu_eager_baseline` leg then time the reference's own modules
(`cpu_baseline.kind = "reference"`); when baseline/_ref is absent they fall back to the oracle port (`"port"`).
Run by __graft_entry__.build() whenever /root/reference is present.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GLUEFACTORY_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def installed():
    return os.path.isfile(os.path.join(DST, "gluefactory", "models", "matchers", "lightglue.py"))


def stage(verbose=True, force=False):
    """Returns True when baseline/_ref holds the installed reference."""
    if installed() and not force:
        return True
    if not os.path.isdir(REF):
        if verbose:
            print(f"[stage_reference] {REF} not present and baseline/_ref empty: reference arm falls back to the port")
        return False
    tmp = tempfile.mkdtemp(prefix="gf_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, symlinks=True, ignore=shutil.ignore_patterns(".git"))
        os.makedirs(DST, exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--upgrade", "--target", DST, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or r.returncode:
            print(r.stdout[-800:])
        if r.returncode:
            raise RuntimeError("pip install of the reference into baseline/_ref failed")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return installed()


def import_reference():
    """Put the installed reference (and the omegaconf stand-in it needs) on sys.path; returns get_model or None."""
    if not installed():
        return None
    for p in (os.path.join(HERE, "_shim"), DST):
        if p not in sys.path:
            sys.path.insert(0, p)
    from gluefactory.models import get_model  # noqa: E402

    return get_model


if __name__ == "__main__":
    print("installed:", stage(force="-f" in sys.argv))
