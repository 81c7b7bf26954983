"""CPU tests of the host-side logic: C-ABI exports, argument / flag validation, loud failure without a GPU."""
import ctypes
import os
import re
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from real_time_helmet_detection_b200 import _lib, hourglass  # noqa: F401  (hourglass registers the hd_net_* symbols)
    header = open(os.path.join(ROOT, "include", "hd_b200.h")).read()
    declared = set(re.findall(r"\b(hd_[a-z0-9_]+)\s*\(", header))
    declared -= {"hd_stream_t"}
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.exported_symbols()) <= declared | {"hd_version", "hd_last_error"}
    assert _lib.lib().hd_version() == 1


def test_unsupported_flags_raise_like_the_reference():
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    with pytest.raises(NotImplementedError, match="Not expected activation"):
        StackedHourglass(1, 128, 6, activation="Swish")
    with pytest.raises(NotImplementedError, match="Not expected pool"):
        StackedHourglass(1, 128, 6, pool="Median")
    for kw in (dict(activation="Mish"), dict(pool="SPP"), dict(neck_pool="Max"), dict(increase_ch=32)):
        with pytest.raises(NotImplementedError):
            StackedHourglass(1, 128, 6, **kw)


def test_no_cpu_fallback():
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    from real_time_helmet_detection_b200.loss import LossCalculator
    from real_time_helmet_detection_b200.transform import hm2box
    net = StackedHourglass(1, 128, 6)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 100, 100))          # H, W must be multiples of 64, as in the reference
    with pytest.raises(RuntimeError, match="CUDA"):
        hm2box(torch.zeros(2, 8, 8), torch.zeros(2, 8, 8), torch.zeros(2, 8, 8))
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        crit(torch.rand(1, 2, 8, 8), torch.rand(1, 2, 8, 8), torch.rand(1, 2, 8, 8), torch.rand(1, 2, 8, 8),
             torch.rand(1, 2, 8, 8), torch.rand(1, 2, 8, 8), torch.rand(1, 1, 8, 8))


def test_native_planner_and_argument_validation():
    from real_time_helmet_detection_b200 import _lib
    from real_time_helmet_detection_b200.hourglass import StackedHourglass
    L = _lib.lib()
    net = StackedHourglass(2, 128, 6)
    h = net._native()
    assert L.hd_net_num_units(h) == len(net.units()) == 70
    small = L.hd_net_workspace_bytes(h, 2, 128, 128, 1)
    big = L.hd_net_workspace_bytes(h, 4, 128, 128, 1)
    assert 0 < small < big and L.hd_net_workspace_bytes(h, 2, 100, 128, 1) == 0
    assert L.hd_net_workspace_bytes(h, 2, 128, 128, 0) < small
    # shape validation happens before any CUDA call
    assert L.hd_conv2d_igemm(None, None, None, None, None, None, None, None, 1, 8, 8, 100, 128, 128, 3, 0, 128, 0, 0, 1,
                             None) == -22
    assert b"multiple of 64" in L.hd_last_error()
    assert L.hd_decode_nms(None, 0, 0, None, 0, 0, None, 0, 0, 1, 1, 2, 4, 4, 1000, 4.0, 0.2, 0.2, 0, 0, 1, None, None,
                           None, None, None, None) == -22
    assert b"out of range" in L.hd_last_error()


def test_loss_log_contract():
    from real_time_helmet_detection_b200.loss import LossCalculator
    crit = LossCalculator(1.0, 1.0, 0.1, 2.0, 4.0)
    assert crit.log == {"hm": [], "offset": [], "size": [], "total": []}
    crit.log = {k: [1.0, 3.0] for k in ("hm", "offset", "size", "total")}
    assert crit.get_log() == "hm:  2.00, offset:  2.00, size:  2.00, total:  2.00"


def test_product_never_imports_oracle_or_baseline():
    """The oracle (oracle/) and the staged reference / library-bar helpers (baseline/) are test and measurement
    infrastructure: no module of the package, the runner or the C sources may import, include or execute them."""
    pkg = os.path.join(ROOT, "real_time_helmet_detection_b200")
    offenders = []
    for base, _dirs, files in os.walk(pkg):
        if os.path.basename(base) in ("build", "__pycache__"):
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                continue
            text = open(os.path.join(base, f), errors="ignore").read()
            if (re.search(r"^\s*(from|import)\s+(oracle|baseline)\b", text, re.M) or
                    re.search(r'^\s*#\s*include\s+["<][^">]*(oracle|baseline)', text, re.M)):
                offenders.append(os.path.relpath(os.path.join(base, f), ROOT))
    runner = open(os.path.join(ROOT, "runner", "hd_infer.cpp")).read()
    assert not re.search(r'#\s*include\s+["<][^">]*(oracle|baseline)', runner)
    assert not offenders, offenders


def test_reference_staging_and_shims_layout():
    """tools/stage_reference.py keeps the reference out of the history (baseline/_ref is git-ignored), and the three shim
    modules of INTEGRATION.md exist with the reference's flat names."""
    gi = open(os.path.join(ROOT, ".gitignore")).read()
    assert "baseline/_ref/" in gi
    shims = os.path.join(ROOT, "real_time_helmet_detection_b200", "shims")
    assert sorted(f for f in os.listdir(shims) if f.endswith(".py")) == ["hourglass.py", "loss.py", "transform.py"]
    from baseline import refload
    if refload.available():                      # build container: the staged copies are the unmodified files
        import hashlib
        import json
        man = json.load(open(os.path.join(refload.REF, "MANIFEST.json")))
        for name, digest in man["files"].items():
            assert hashlib.sha256(open(os.path.join(refload.REF, name), "rb").read()).hexdigest() == digest, name
            ref = os.path.join("/root/reference", name)
            if os.path.exists(ref):
                assert open(ref, "rb").read() == open(os.path.join(refload.REF, name), "rb").read(), name
        diff = open(os.path.join(refload.REF, "patched", "squeeze_patch.diff")).read()
        changed = [l for l in diff.splitlines() if l[:1] in "+-" and l[:3] not in ("+++", "---")]
        assert len(changed) == 4 and all("squeeze" in l for l in changed)       # the one-token fix, twice


def test_documented_knobs_exist_in_the_sources():
    """Every HD_* environment knob / macro the documents name (profiles/README.md, README.md)
    is spelled the way the sources read it - a renamed knob would otherwise turn an A/B instruction into a silent no-op."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    named = set()
    for doc in ("INTEGRATION.md", "DESIGN.md", os.path.join("profiles", "README.md"), "README.md"):
        with open(os.path.join(root, doc)) as f:
            named |= set(re.findall(r"\bHD_[A-Z0-9_]{3,}\b", f.read()))
    text = []
    pats = ["real_time_helmet_detection_b200/**/*", "tools/*.py", "tests/*.py", "include/*.h", "runner/*.cpp", "bench.py", "__graft_entry__.py"]
    for pat in pats:
        for path in glob.glob(os.path.join(root, pat), recursive=True):
            if os.path.isfile(path) and path.endswith((".cu", ".cuh", ".h", ".py", ".cpp")):
                with open(path, errors="ignore") as f:
                    text.append(f.read())
    text = "\n".join(text)
    missing = sorted(k for k in named if k not in text)
    assert not missing, missing
