"""TEST / MEASUREMENT INFRASTRUCTURE — recipe that stages the UNMODIFIED reference for the GPU box.

The reference (saiakarsh193/GenVox) is pure Python: "building" it is making its package tree importable.  /root/reference
exists only in the build container, so this recipe copies the reference's own Python packages (models/, configs/, core/,
utils/ - `*.py` only, byte for byte, nothing generated or edited) into oracle/_ref/, which is git-ignored (never part of
the repo's history) but travels to the GPU box with the snapshot, exactly like the built .so files.

    python -m oracle.build_ref          # also run by __graft_entry__.build() when /root/reference is present

Users: oracle/ref_import.py (falls back to oracle/_ref when /root/reference is absent), and through it
bench.py --impl reference / the cpu_baseline and gpu_torch_baseline legs ("kind": "reference").  Never the product.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("GENVOX_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("models", "configs", "core", "utils")


def staged() -> bool:
    return os.path.isfile(os.path.join(DST, "models", "tts", "tacotron2.py"))


def build(verbose: bool = True) -> bool:
    """Copy the reference packages into oracle/_ref.  Returns False (and leaves oracle/_ref alone) when the reference
    tree is not present - the GPU box uses what the build container staged."""
    if not os.path.isfile(os.path.join(SRC, "models", "tts", "tacotron2.py")):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for pkg in PACKAGES:
        for root, _, files in os.walk(os.path.join(SRC, pkg)):
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                src = os.path.join(root, f)
                rel = os.path.relpath(src, SRC)
                dst = os.path.join(DST, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                with open(src, "rb") as fh:
                    manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"staged {len(manifest)} reference files into {DST}")
    return True


if __name__ == "__main__":
    raise SystemExit(0 if build() else 1)
