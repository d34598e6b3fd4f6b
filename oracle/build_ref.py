"""Recipe that populates oracle/_ref/ with the UNMODIFIED reference (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference is a pure-Python application (no build system, no setup.py): "building" it for the CPU baseline arm means
copying the Python packages the hot path imports — ltx_video/, wan/, utils/ — byte for byte from /root/reference into
oracle/_ref/, which is git-ignored (no reference source enters the history) but NOT gpurun-ignored, so it travels to
the GPU box where /root/reference does not exist.  `bench.py --impl reference` and bench.py's cpu_baseline leg import
it from there through oracle/refshim (stand-ins for the third-party `diffusers` / `mmgp`, which are not installed).
Run by __graft_entry__.build() whenever /root/reference is present:   python oracle/build_ref.py
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("LTX_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("ltx_video", "wan", "utils")
KEEP = (".py", ".json", ".yaml", ".yml", ".txt")


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "ltx_video")):
        print(f"build_ref: {SRC} not present; oracle/_ref left as it is ({'present' if os.path.isdir(DST) else 'absent'})")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n, digest = 0, hashlib.sha256()
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = sorted(d for d in dirs if d != "__pycache__")
            for f in sorted(files):
                if not f.endswith(KEEP):
                    continue
                s = os.path.join(root, f)
                d = os.path.join(DST, os.path.relpath(s, SRC))
                os.makedirs(os.path.dirname(d), exist_ok=True)
                shutil.copyfile(s, d)
                with open(s, "rb") as fh:
                    digest.update(fh.read())
                n += 1
    with open(os.path.join(DST, "MANIFEST"), "w") as fh:
        fh.write(f"source {SRC}\nfiles {n}\nsha256 {digest.hexdigest()}\n")
    print(f"build_ref: copied {n} files of {PACKAGES} from {SRC} to {DST} (sha256 {digest.hexdigest()[:16]})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
