"""Stages the UNMODIFIED reference package (plus the statsmodels stand-in) under baseline/_ref/ so that
`bench.py --impl reference` and the `cpu_baseline` leg can time the reference ITSELF on the GPU box's
host cores (baseline/_ref is git-ignored but travels with the gpurun snapshot).

    python baseline/stage_reference.py        # build container only: needs /root/reference

Install route.  The base contract's `pip install --no-index --no-build-isolation --find-links
/opt/wheelhouse --target baseline/_ref /root/reference` fails at dependency resolution (statsmodels is an
unpinned dependency, requirements.txt:4 / setup.py:36, and is not in the offline wheelhouse), so the
fallback it names is used: `--no-deps` from a copy under /tmp (the source tree is read-only).  Two
accommodations are then applied, both recorded in baseline/_ref/STAGING.txt and in DESIGN.md:
  1. statsmodels: the ~50-line stand-in oracle/_shim/statsmodels (OLS by pinv, add_constant; SURVEY.md
     Appendix A) is copied next to the package.  It reproduces the reference's golden CSVs.
  2. pandas 3: inner_model.py:75 `path.loc[dv,]` raises under pandas 3.0 (tuple indexing with a trailing
     comma); that ONE expression is rewritten as `path.loc[dv]`.  Nothing else is touched.
Nothing under baseline/_ref is imported by the product or by the tests' CUDA path.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PLSPM_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(force: bool = False) -> str:
    """Returns the staging directory, or '' when the reference source is not available here (GPU box)."""
    marker = os.path.join(DST, "STAGING.txt")
    if os.path.exists(marker) and not force:
        return DST
    if not os.path.isdir(os.path.join(REF, "plspm")):
        return DST if os.path.exists(marker) else ""
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    how = "pip --no-deps --target"
    tmp = tempfile.mkdtemp(prefix="plspm_ref_src_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns(".git", "docs", "tests"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DST, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(os.path.join(DST, "plspm", "weights.py")):
            how = "copy of the package directory (pip --no-deps failed: %s)" % (r.stderr.strip().splitlines() or ["?"])[-1]
            shutil.rmtree(os.path.join(DST, "plspm"), ignore_errors=True)
            shutil.copytree(os.path.join(REF, "plspm"), os.path.join(DST, "plspm"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    p = os.path.join(DST, "plspm", "inner_model.py")
    text = open(p).read()
    old, new = "path.loc[dv,][path.loc[dv,] == 1]", "path.loc[dv][path.loc[dv] == 1]"
    assert old in text, "reference changed: inner_model.py:75 not found"
    open(p, "w").write(text.replace(old, new))
    shutil.copytree(os.path.join(ROOT, "oracle", "_shim", "statsmodels"), os.path.join(DST, "statsmodels"))
    with open(marker, "w") as f:
        f.write("reference: %s (plspm 0.5.7)\ninstall: %s\naccommodations: statsmodels stand-in (oracle/_shim); "
                "inner_model.py:75 `path.loc[dv,]` -> `path.loc[dv]` (pandas 3)\n" % (REF, how))
    return DST


if __name__ == "__main__":
    d = stage(force="--force" in sys.argv)
    print(d or "reference source not found; nothing staged")
    if d:
        print(open(os.path.join(d, "STAGING.txt")).read())
