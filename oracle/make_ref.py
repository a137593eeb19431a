"""Recipe: put the UNMODIFIED reference (BetterBelle/eco-dqn) next to the oracle so that the GPU box, which has no
/root/reference, can time the reference's own CPU implementation of the hot path (bench.py --impl reference, and the
`cpu_baseline` block of the product arm).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nothing under eco-dqn_b200/ imports it.

    python oracle/make_ref.py            # run in the build container; __graft_entry__.build() calls it

What it does (outputs only under oracle/_ref/, which is git-ignored -- no reference source enters the history -- but NOT
gpurun-ignored, so it travels to the GPU box like the built .so):

  * copies /root/reference/src/  (envs, networks, agents: the reference's hot path, byte for byte)
  * copies /root/reference/experiments/utils.py  (test_network / __test_network_batched, the driver the metric times)
  * writes a three-line `docplex` stub: src/agents/solver.py imports docplex.mp.model at module level for its
    CplexSolver (never used on this path); docplex/cplex are proprietary and absent from this image
  * writes MANIFEST.json with the sha256 of every copied file, so a reader can check the copy is unmodified

The reference is pure Python (+ numba JIT for one function): there is nothing to compile.  It has no setup.py /
pyproject.toml, so `pip install --target baseline/_ref /root/reference` (the base contract's install) does not apply.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ECO_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make(verbose=True):
    if not os.path.isdir(os.path.join(REF, "src")):
        if verbose:
            print("make_ref: %s not present (GPU box?) -- keeping whatever oracle/_ref holds" % REF)
        return os.path.isdir(os.path.join(OUT, "src"))
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(os.path.join(OUT, "experiments"))
    shutil.copytree(os.path.join(REF, "src"), os.path.join(OUT, "src"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copy2(os.path.join(REF, "experiments", "utils.py"), os.path.join(OUT, "experiments", "utils.py"))
    open(os.path.join(OUT, "experiments", "__init__.py"), "w").close()
    os.makedirs(os.path.join(OUT, "docplex", "mp"))
    open(os.path.join(OUT, "docplex", "__init__.py"), "w").close()
    open(os.path.join(OUT, "docplex", "mp", "__init__.py"), "w").close()
    with open(os.path.join(OUT, "docplex", "mp", "model.py"), "w") as f:
        f.write("# stub written by oracle/make_ref.py: the reference imports this name for its CplexSolver only\nModel = object\n")
    manifest = {}
    for root, _, files in os.walk(OUT):
        for fn in sorted(files):
            p = os.path.join(root, fn)
            rel = os.path.relpath(p, OUT)
            src = os.path.join(REF, rel)
            if os.path.exists(src):
                assert sha256(p) == sha256(src), rel
                manifest[rel] = sha256(p)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"reference": "BetterBelle/eco-dqn", "copied_from": REF, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("make_ref: %d reference files copied unmodified into %s" % (len(manifest), OUT))
    return True


def import_reference():
    """Make `src.*` and `experiments.utils` of oracle/_ref importable; returns False when the copy is absent."""
    if not os.path.isdir(os.path.join(OUT, "src")):
        return False
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    return True


if __name__ == "__main__":
    make()
