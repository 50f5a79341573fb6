"""TEST INFRASTRUCTURE ONLY -- recipe that vendors the UNMODIFIED reference package into oracle/_ref/.

    python oracle/make_ref.py            # needs /root/reference (the build container)

The reference (SensorsINI/v2e) is pure Python: "building" it is copying its package directory as it is.
oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but NOT
gpurun-ignored, so the copy travels to the GPU box next to the built .so files. There it is the CPU arm of
bench.py (`--impl reference`, cpu_baseline kind "_ref") and the checker of the full-size parity tests;
nothing under v2e_b200/ imports it. __graft_entry__.build() runs this recipe when /root/reference exists.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("V2E_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
# the hot path and its callers: the package, minus notebooks / images / GUI launchers
SKIP_DIRS = {"desktop", "ddd20_interfaces", "ddd20_utils", "__pycache__"}
SKIP_EXT = {".ipynb", ".png", ".md"}


def make_ref(force=False):
    src_pkg = os.path.join(SRC, "v2ecore")
    if not os.path.isfile(os.path.join(src_pkg, "emulator.py")):
        return None                       # not in the build container: use what travelled (or nothing)
    dst_pkg = os.path.join(DST, "v2ecore")
    stamp = os.path.join(DST, ".stamp")
    newest = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(src_pkg) for f in fs)
    if not force and os.path.exists(stamp) and os.path.getmtime(stamp) >= newest:
        return dst_pkg
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    for root, dirs, files in os.walk(src_pkg):
        dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
        rel = os.path.relpath(root, src_pkg)
        out = os.path.join(dst_pkg, rel) if rel != "." else dst_pkg
        os.makedirs(out, exist_ok=True)
        for f in files:
            if os.path.splitext(f)[1] in SKIP_EXT:
                continue
            shutil.copy2(os.path.join(root, f), os.path.join(out, f))
    with open(stamp, "w") as fh:
        fh.write("copied from %s\n" % src_pkg)
    return dst_pkg


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
