#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).

    python tools/stage_reference.py [--reference /root/reference]

baseline/_ref/           byte-for-byte copies of the reference's flat modules (hourglass, loss, transform, evaluate, train,
                         optim, utils, config, data, main) - what `bench.py --impl reference`, the bench's `library_bar` /
                         `cpu_baseline` legs and tests/test_reference_drivers_gpu.py import when present.
baseline/_ref/patched/   train.py and evaluate.py with the ONE-token fix torch >= 2 needs to run them at all
                         (`output.squeeze_(1)` on a split() view -> `output = output.squeeze(1)`);
                         the unified diff is written next to them (squeeze_patch.diff) and checked to touch exactly the
                         expected lines.
baseline/_ref/MANIFEST.json   sha256 of every staged file + the reference commit.

Nothing here is product source: the package never imports baseline/ (tests/test_host_logic.py checks that).
"""
from __future__ import annotations

import argparse
import difflib
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ["hourglass.py", "loss.py", "transform.py", "evaluate.py", "train.py", "optim.py", "utils.py", "config.py",
         "data.py", "main.py"]
PATCHES = {
    # file: (old line content (stripped), new line content (stripped), expected occurrences)
    "train.py": ("output.squeeze_(1)", "output = output.squeeze(1)", 1),
    "evaluate.py": ("output.squeeze_(1)", "output = output.squeeze(1)", 1),
}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("HD_REFERENCE", "/root/reference"))
    args = ap.parse_args()
    ref = args.reference
    if not os.path.isdir(ref):
        print(f"stage_reference: {ref} not found (only the build container has it)", file=sys.stderr)
        return 1
    dst = os.path.join(ROOT, "baseline", "_ref")
    os.makedirs(os.path.join(dst, "patched"), exist_ok=True)
    manifest = {"files": {}, "patched": {}}
    sub = os.path.join(ref, ".SUBMODULES.json")
    if os.path.exists(sub):
        manifest["reference_meta"] = json.load(open(sub))
    for f in FILES:
        shutil.copyfile(os.path.join(ref, f), os.path.join(dst, f))
        manifest["files"][f] = sha(os.path.join(dst, f))
    diffs = []
    for f, (old, new, expect) in PATCHES.items():
        src = open(os.path.join(ref, f)).read().splitlines(keepends=True)
        out, hits = [], 0
        for line in src:
            if line.strip() == old:
                indent = line[:len(line) - len(line.lstrip())]
                out.append(indent + new + "\n")
                hits += 1
            else:
                out.append(line)
        if hits != expect:
            raise SystemExit(f"stage_reference: expected {expect} `{old}` line(s) in {f}, found {hits}")
        with open(os.path.join(dst, "patched", f), "w") as fh:
            fh.writelines(out)
        d = list(difflib.unified_diff(src, out, f"a/{f}", f"b/{f}", n=0))
        changed = [l for l in d if l[:1] in "+-" and l[:3] not in ("+++", "---")]
        assert len(changed) == 2 * expect, changed
        diffs += d
        manifest["patched"][f] = sha(os.path.join(dst, "patched", f))
    with open(os.path.join(dst, "patched", "squeeze_patch.diff"), "w") as fh:
        fh.writelines(diffs)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print(f"staged {len(FILES)} reference modules under {dst}; patch:\n" + "".join(diffs))
    return 0


if __name__ == "__main__":
    sys.exit(main())
