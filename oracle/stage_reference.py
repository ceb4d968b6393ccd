"""Pack the UNMODIFIED reference files that the end-to-end acceptance test drives (RAFT model, shipped
raft-small.pth, two demo frames) into oracle/_ref/reference_raft.tar so they can travel to the GPU box, where
/root/reference does not exist.  Test infrastructure only: the archive is git-ignored, nothing in it is
imported by the product, and no reference source is copied into the tracked tree.
Exits 0 with a note when the reference checkout is absent."""
import os
import sys
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RAFT_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "reference_raft.tar")
FILES = ["core/__init__.py", "core/raft.py", "core/update.py", "core/extractor.py", "core/corr.py",
         "core/utils/__init__.py", "core/utils/utils.py", "raft-small.pth",
         "demo-frames/frame_0016.png", "demo-frames/frame_0017.png"]


def main():
    if not os.path.isdir(os.path.join(REF, "core")):
        print(f"stage_reference: {REF} not present; keeping whatever is in {os.path.dirname(OUT)}")
        return 0
    srcs = [os.path.join(REF, f) for f in FILES]
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        print(f"stage_reference: {OUT} up to date")
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with tarfile.open(OUT, "w") as tar:
        for f, s in zip(FILES, srcs):
            tar.add(s, arcname=f)
    print(f"stage_reference: wrote {OUT} ({os.path.getsize(OUT) / 1e6:.1f} MB)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
