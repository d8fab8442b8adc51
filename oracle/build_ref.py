"""
Recipe for oracle/_ref/: the reference's own Python package, copied UNMODIFIED from /root/reference/accbpg so that
bench.py's CPU arm (`--impl reference`, `cpu_baseline.kind == "reference"`) can time the real reference on the GPU box,
where /root/reference does not exist.  TEST / MEASUREMENT INFRASTRUCTURE.

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

oracle/_ref/ is listed in .gitignore (nothing of the reference enters the history) and not in .gpurunignore (it travels
to the GPU box with the snapshot, like a built .so).  The reference is pure Python: "building" it is a file copy.  Only
the package directory is taken (12 .py files); data files, notebooks and archives stay where they are.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(os.environ.get("ACCBPG_REFERENCE", "/root/reference"), "accbpg")
DST = os.path.join(HERE, "_ref", "accbpg")


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: oracle/_ref left as it is")
        return None
    os.makedirs(DST, exist_ok=True)
    digest = hashlib.sha256()
    names = sorted(n for n in os.listdir(SRC) if n.endswith(".py"))
    for name in names:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        digest.update(open(os.path.join(DST, name), "rb").read())
    with open(os.path.join(HERE, "_ref", "MANIFEST.txt"), "w") as fh:
        fh.write("unmodified copy of /root/reference/accbpg/*.py (oracle/build_ref.py)\n")
        fh.write("files: " + " ".join(names) + "\n")
        fh.write("sha256 of the concatenation: " + digest.hexdigest() + "\n")
    if verbose:
        print(f"oracle/_ref/accbpg: {len(names)} files, sha256 {digest.hexdigest()[:16]}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if build() or True else 1)
