#!/usr/bin/env python
"""Recipe that makes the UNMODIFIED reference travel: copies the reference's own Python sources for the hot path's callers
from /root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored -- it ships to the GPU box like a built .so).

    python oracle/build_ref.py            # no-op (keeps what is there) when /root/reference is absent, e.g. on the GPU box

Why this directory exists: the round-1 review ("What's missing" #2 and "Next round" #1) asked for exactly this --
"the claim 'the reference cannot travel to the GPU box' is wrong: gitignored artefacts ship with gpurun, so a committed recipe
that populates oracle/_ref/ (already in .gitignore) from /root/reference would ship too ... Commit a recipe (oracle/build_ref.py)
that copies the needed NP/*.py into gitignored oracle/_ref/" -- so that the UNMODIFIED reference is the timed baseline, the
PSNR twin's other arm and the script the drop-in is exercised under, on the GPU box.  The files are the reference itself, kept
OUT of the repository's history on purpose (they are not this project's code and are never imported by the product).

Test infrastructure, not product: only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline / quality legs
import anything from oracle/_ref.  Nothing is edited: every file is byte-identical to its source and the manifest
(oracle/_ref/MANIFEST.json) records source path + sha256 so that this can be checked.  No reference source enters git.

What is copied (NP = /root/reference/nerf-pytorch-master):
  run_nerf.py, run_nerf_helpers.py                      the vanilla script + model/sampling helpers
  run_nerf_view.py, run_nerf_view_cal_correspondance.py the ConsistentNeRF scripts (train(), hard-mask loop, losses)
  load_blender.py, load_dtu.py, load_llff.py, load_deepvoxels.py, load_LINEMOD.py     their dataset loaders
  alky/vis_utils.py                                     img2psnr_mask / img2ssim / depth visualisation used by train()
NOT copied: configs/pairs.th (a pickle of numpy arrays: torch.load refuses it under weights_only, and unpickling untrusted
content is avoided altogether) -- oracle/twin.py writes an equivalent file of plain lists for its synthetic scenes.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
SRC = "/root/reference/nerf-pytorch-master"
FILES = ["run_nerf.py", "run_nerf_helpers.py", "run_nerf_view.py", "run_nerf_view_cal_correspondance.py", "load_blender.py",
         "load_dtu.py", "load_llff.py", "load_deepvoxels.py", "load_LINEMOD.py", "alky/vis_utils.py"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available() -> bool:
    """oracle/_ref holds every file of the manifest (built here, shipped to the GPU box)."""
    man = os.path.join(REF, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    try:
        files = json.load(open(man))["files"]
    except Exception:
        return False
    return all(os.path.exists(os.path.join(REF, f)) and _sha(os.path.join(REF, f)) == meta["sha256"] for f, meta in files.items())


def build(verbose: bool = True) -> bool:
    """Populate oracle/_ref from /root/reference; returns available()."""
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[build_ref] {SRC} not present: keeping oracle/_ref as shipped (available={available()})", file=sys.stderr)
        return available()
    manifest = {"source": SRC, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(REF, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest["files"][rel] = {"source": src, "sha256": _sha(dst), "bytes": os.path.getsize(dst)}
        assert _sha(src) == manifest["files"][rel]["sha256"]
    with open(os.path.join(REF, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    with open(os.path.join(REF, "README.txt"), "w") as f:
        f.write("UNMODIFIED reference files (skhu101/ConsistentNeRF, nerf-pytorch-master), copied byte for byte by oracle/build_ref.py\n"
                "(sha256 in MANIFEST.json).  Git-ignored on purpose: not this project's code, never imported by consistentnerf_b200/.\n"
                "They are here so that the reference itself can run on the GPU box (bench.py --impl reference, cpu_baseline, gpu_eager,\n"
                "quality; tests/test_gpu_dropin_train.py) -- requested by the round-1 review (VERDICT.md, missing #2 / next #1).\n")
    if verbose:
        print(f"[build_ref] copied {len(FILES)} unmodified reference files into {REF}", file=sys.stderr)
    return available()


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
