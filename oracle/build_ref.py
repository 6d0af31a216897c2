"""Recipe that installs the UNMODIFIED reference package into oracle/_ref/ -- TEST / BASELINE INFRASTRUCTURE ONLY.

daisyRec is pure Python (no compiled sources), so "building" the reference is installing its package: this script
pip-installs `/root/reference` (from a scratch copy under /tmp: the tree is read-only and setuptools writes build/
and egg-info next to setup.py) with `--no-deps --target oracle/_ref`.  colorlog / colorama / optuna, which only
`daisy.utils.config` and `tune.py` need, are absent from the image; nothing on the timed path imports them.  If pip
cannot run, the package directory is copied as-is (same files).  Nothing is edited; the three library-compat shims
live in oracle/ref_harness.py and are applied in process.

oracle/_ref/ is git-ignored (no reference sources enter the history) but travels to the GPU box with gpurun, where
`bench.py --impl reference` and the `cpu_baseline` leg time `daisy.model.MFRecommender.MF.fit` over a real
`torch.utils.data.DataLoader` on the box's host cores.  The product (daisyrec_b200/) never imports it.

    python oracle/build_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("DAISY_REF_SRC", "/root/reference")
OUT = os.path.join(HERE, "_ref")
STAMP = os.path.join(OUT, ".installed_from")


def installed():
    return os.path.isfile(os.path.join(OUT, "daisy", "model", "MFRecommender.py"))


def build(force=False):
    """-> path of oracle/_ref when the reference is installed there (now or earlier), else None."""
    if installed() and not force:
        return OUT
    if not os.path.isdir(os.path.join(REF_SRC, "daisy")):
        return OUT if installed() else None          # GPU box: only the prebuilt copy exists
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    how = "pip"
    with tempfile.TemporaryDirectory(prefix="daisy_ref_src_") as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns("data", "images", ".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", OUT, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0 or not installed():
            how = "copytree (pip failed: %s)" % (r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.returncode)
    if not installed():
        shutil.copytree(os.path.join(REF_SRC, "daisy"), os.path.join(OUT, "daisy"), dirs_exist_ok=True)
    assets = os.path.join(OUT, "daisy", "assets")
    if not os.path.isfile(os.path.join(assets, "mf.yaml")):  # setup.py may not declare the YAML assets as package data
        shutil.copytree(os.path.join(REF_SRC, "daisy", "assets"), assets, dirs_exist_ok=True)
    with open(STAMP, "w") as f:
        f.write(f"{REF_SRC} via {how}\n")
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference tree not present; oracle/_ref not built")
