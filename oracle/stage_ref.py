"""Recipe that stages the UNMODIFIED reference sources of the hot path into ``oracle/_ref/`` (TEST / BENCH INFRASTRUCTURE).

The reference is pure Python, so there is nothing to compile: "building" the reference arm means copying, byte for byte,
the files that implement the path from the read-only checkout into the git-ignored directory ``oracle/_ref/`` -- which,
like the built ``.so``, travels to the GPU box with the working tree (it is not listed in ``.gpurunignore``) while never
entering the repository's history.  ``__graft_entry__.build()`` runs this whenever ``/root/reference`` is present;
``bench.py --impl reference`` and the ``cpu_baseline`` leg then time these files on the box's host cores
(``oracle/ref_runner.py``), under the ``oracle/_gym_stub`` stand-in for gymnasium.

    python oracle/stage_ref.py [--reference /root/reference]

Files: src/env/hedging_env_v2.py (A1-A5), src/env/hedging_env.py (v1),
src/sim/option_price_assignment.py (A6-A8), src/tools/bs_delta.py (A12).  A MANIFEST.json with the sha256 of each file is
written next to them; ``ref_runner.staged()`` re-checks it before anything is timed.
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("src/env/hedging_env_v2.py", "src/env/hedging_env.py", "src/sim/option_price_assignment.py",
         "src/tools/bs_delta.py")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(reference="/root/reference", dest=REF_DIR):
    """Copy FILES from ``reference`` to ``dest`` (flat), write MANIFEST.json; returns the manifest dict."""
    if not os.path.isdir(reference):
        raise FileNotFoundError(f"{reference}: reference checkout not present")
    os.makedirs(dest, exist_ok=True)
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference, rel)
        dst = os.path.join(dest, os.path.basename(rel))
        shutil.copyfile(src, dst)
        manifest[os.path.basename(rel)] = dict(source=rel, sha256=sha256(dst), bytes=os.path.getsize(dst))
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    for name, m in stage(a.reference).items():
        print(f"staged {m['source']} -> oracle/_ref/{name}  sha256 {m['sha256'][:16]}  {m['bytes']} B")
