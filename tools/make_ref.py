#!/usr/bin/env python
"""Build `baseline/_ref/` -- a RUNNABLE copy of the reference's own `multimae` code -- from /root/reference.

    python tools/make_ref.py            (authoring container only: it reads /root/reference)

`baseline/_ref/` is git-ignored (the reference's sources never enter this repository's history) but not
gpurun-ignored, so it travels with the tree to the GPU box, where /root/reference does not exist.  It is what
`bench.py --impl reference`, `bench.py`'s `cpu_baseline` / `gpu_eager_reference` legs and
`tests/test_reference_parity_gpu.py` run: the reference's stock code path, none of this repository's kernels.

Layout written:
  baseline/_ref/refmm/      <- pretraining/multimae/*.py                     (imported as `refmm`)
  baseline/_ref/refdown/    <- downstream/instance_segmentation/modeling/multimae/{multimae_utils, zorro_utils,
                               input_adapters, multimae_big_imcomplete}.py   (imported as `refdown`; ViTBaseline)
  baseline/_ref/MANIFEST.json   source file -> sha256 of the source, and the edits applied

The ONLY edits (without them the package does not import or run at all):
  1. pretraining/multimae/zorro_utils.py:255 holds U+FF1A (a full-width colon) where ':' is meant: SyntaxError for the
     whole package.  Replaced by ':'.
  2. the same file's `Block_Fusion` (:243-258) is an unfinished draft (4-D input sliced as 3-D); the class the authors
     actually run is the one in downstream/instance_segmentation/modeling/multimae/zorro_utils.py:243-258 (the two files
     differ in exactly these lines).  The downstream class body replaces the draft.
  3. refdown/__init__.py is empty (the original __init__ imports the detectron2 heads, which are not on this path).
Nothing else is touched (CRLF line ends are normalised to LF).  The per-step `print(rand_per_sample_choice.shape)` of
multimae.py:177 is left in; callers that need a clean stdout redirect it.
"""
import hashlib
import io
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MMF_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PRE = os.path.join("pretraining", "multimae")
DOWN = os.path.join("downstream", "instance_segmentation", "modeling", "multimae")
DOWN_FILES = ("multimae_utils.py", "zorro_utils.py", "input_adapters.py", "multimae_big_imcomplete.py")


def _read(path):
    with io.open(path, "r", encoding="utf-8") as f:
        return f.read().replace("\r\n", "\n")


def _class_src(src, name):
    return re.search(r"^class %s\(.*?(?=^class |\Z)" % name, src, flags=re.S | re.M).group(0)


def available():
    return os.path.isdir(os.path.join(REF, PRE))


def build(force=False, quiet=False):
    """returns the path of baseline/_ref (built now if the reference tree is present, else whatever is already there)"""
    if not available():
        return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None
    manifest = {"source": REF, "files": {}, "edits": []}
    out = {}
    down_zorro = _read(os.path.join(REF, DOWN, "zorro_utils.py"))
    for fn in sorted(os.listdir(os.path.join(REF, PRE))):
        if not fn.endswith(".py"):
            continue
        sp = os.path.join(REF, PRE, fn)
        src = _read(sp)
        manifest["files"][os.path.join("refmm", fn)] = {"from": os.path.join(PRE, fn), "sha256": hashlib.sha256(open(sp, "rb").read()).hexdigest()}
        if fn == "zorro_utils.py":
            n = src.count("：")
            src = src.replace("：", ":")
            manifest["edits"].append("refmm/zorro_utils.py: %d x U+FF1A -> ':' (line 255)" % n)
            src = src.replace(_class_src(src, "Block_Fusion"), _class_src(down_zorro, "Block_Fusion"))
            manifest["edits"].append("refmm/zorro_utils.py: class Block_Fusion taken from %s:243-258" % os.path.join(DOWN, "zorro_utils.py"))
        out[os.path.join("refmm", fn)] = src
    for fn in DOWN_FILES:
        sp = os.path.join(REF, DOWN, fn)
        manifest["files"][os.path.join("refdown", fn)] = {"from": os.path.join(DOWN, fn), "sha256": hashlib.sha256(open(sp, "rb").read()).hexdigest()}
        out[os.path.join("refdown", fn)] = _read(sp)
    out[os.path.join("refdown", "__init__.py")] = ""
    manifest["edits"].append("refdown/__init__.py: empty (the original imports the detectron2 heads)")
    stamp = hashlib.sha256(json.dumps({k: hashlib.sha256(v.encode()).hexdigest() for k, v in sorted(out.items())}).encode()).hexdigest()
    manifest["stamp"] = stamp
    mpath = os.path.join(DST, "MANIFEST.json")
    if not force and os.path.exists(mpath):
        try:
            if json.load(open(mpath)).get("stamp") == stamp:
                return DST
        except ValueError:
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for rel, src in out.items():
        p = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with io.open(p, "w", encoding="utf-8") as f:
            f.write(src)
    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1)
    if not quiet:
        print("wrote %s (%d files)" % (DST, len(out)))
    return DST


def load(which="refmm"):
    """import the reference package (`refmm` = pretraining multimae, `refdown` = the downstream ViTBaseline files) from
    baseline/_ref; returns None when it was never built (no /root/reference and no earlier build)"""
    if which in sys.modules:
        return sys.modules[which]
    dst = build(quiet=True)
    if dst is None:
        return None
    if dst not in sys.path:
        sys.path.insert(0, dst)
    import importlib
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if which == "refmm":
            pkg = importlib.import_module("refmm")
            for mod in ("multimae_utils", "zorro_utils", "input_adapters", "output_adapters", "output_adapters_simple", "criterion",
                        "multimae", "multimae_crossattn", "multimae_lstm_s2dsm"):
                importlib.import_module("refmm." + mod)
            return pkg
        pkg = importlib.import_module("refdown")
        for mod in ("multimae_utils", "zorro_utils", "input_adapters", "multimae_big_imcomplete"):
            importlib.import_module("refdown." + mod)
        return pkg


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference tree not found at %s and no earlier build" % REF)
