"""Recipe for ``oracle/_ref``: a git-ignored copy of the reference's Python packages, made where ``/root/reference`` exists
(the build container, by ``__graft_entry__.build()``) so that it travels to the GPU box with the repo snapshot.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is pure Python (no build step): "building" it is copying the ``*.py`` files of
its two packages, unmodified, to where ``oracle/ref_shim.py`` can import them when ``/root/reference`` itself is absent.
``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg then time the REAL reference (``kind: "reference"``)
instead of the oracle port.  Nothing under ``shapeformer_b200/`` imports it; ``oracle/_ref/`` is never committed.

    python -m oracle.make_ref            # (re)create oracle/_ref from /root/reference
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("SFB200_REFERENCE_SOURCE", "/root/reference")
PACKAGES = ("shapeformer", "xgutils")


def make(source=SOURCE, dest=DEST):
    """Copy <source>/{shapeformer,xgutils}/**/*.py -> <dest>/ (same relative paths).  Returns the number of files, or 0
    when the source tree is absent (GPU box: the prebuilt copy is used as is)."""
    if not os.path.isdir(os.path.join(source, "shapeformer")):
        return 0
    n = 0
    for pkg in PACKAGES:
        out_root = os.path.join(dest, pkg)
        if os.path.isdir(out_root):
            shutil.rmtree(out_root)
        for root, dirs, files in os.walk(os.path.join(source, pkg)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            for f in files:
                if f.endswith(".py"):
                    rel = os.path.relpath(os.path.join(root, f), source)
                    os.makedirs(os.path.dirname(os.path.join(dest, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(root, f), os.path.join(dest, rel))
                    n += 1
    with open(os.path.join(dest, "README"), "w") as fh:
        fh.write("Unmodified copy of the reference's Python packages made by oracle/make_ref.py; git-ignored, never committed.\n")
    return n


if __name__ == "__main__":
    k = make()
    print(f"oracle/_ref: {k} files copied from {SOURCE}" if k else f"{SOURCE} not present: oracle/_ref left as is")
    sys.exit(0)
